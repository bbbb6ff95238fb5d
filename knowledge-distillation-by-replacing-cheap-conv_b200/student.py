"""DepthwiseStudent -- frozen teacher + student copy whose named blocks are swapped for cheap-conv blocks.

Host-side mirror of models/students/depthwise_student.py:16-284 of the reference: same constructor,
attributes (`teacher`, `student`, `student_hidden_outputs`, `teacher_hidden_outputs`, `save_hidden`,
`replaced_block_names`) and methods (`replace`, `register_hint_layers`, `unfreeze`, `get_block`,
`_set_block`, `forward`, `inference`, `reset`, `train`, `dump_*`), so trainers, configs and checkpoints
(`teacher.*` / `student.<block>.separable_conv.weight` keys) written for the reference keep working.  The
only functional difference is the block class that `replace` instantiates: the kdcc
`DepthwiseSeparableBlock` whose convolutions run in libkdcc.so.

`inference_test` (sliding-window TTA) stitches on the GPU, see kdcc/tta.py.  Out of scope here: the BeautifulTable
rendering of the block report.
"""
import copy
import gc
from functools import reduce

import torch
from torch import nn

from .blocks import DepthwiseSeparableBlock

BLOCKS_LEVEL_SPLIT_CHAR = '.'


class DepthwiseStudent(nn.Module):
    def __init__(self, teacher_model, config):
        super().__init__()
        self.config = config
        # the teacher is cloned twice: one frozen reference copy, one copy that gets operated on
        self.teacher = copy.deepcopy(teacher_model)
        self.teacher.eval()
        for prm in self.teacher.parameters():
            prm.requires_grad = False
        self.student = copy.deepcopy(self.teacher)  # note: inherits eval mode and frozen parameters

        self.replaced_block_names = []
        self.student_hidden_outputs = []
        self.teacher_hidden_outputs = []
        self._student_hook_handlers = []
        self._teacher_hook_handlers = []
        self.aux_block_names = []
        self.save_hidden = True

    # ---- block addressing: dotted attribute paths, digits index Sequential / ModuleList ----------------
    def get_block(self, block_name, model):
        def step(obj, part):
            return obj[int(part)] if part.isdigit() else getattr(obj, part)
        return reduce(step, block_name.split(BLOCKS_LEVEL_SPLIT_CHAR), model)

    def _set_block(self, block_name, block, model):
        parts = block_name.split(BLOCKS_LEVEL_SPLIT_CHAR)
        owner = model if len(parts) == 1 else self.get_block(BLOCKS_LEVEL_SPLIT_CHAR.join(parts[:-1]), model)
        setattr(owner, parts[-1], block)

    # ---- surgery (epoch boundaries) --------------------------------------------------------------------
    def replace(self, blocks, **kwargs):
        """blocks: [{"name": ..., "epoch": ..., "args"(optional): {"kernel_size", "padding", "dilation"}}, ...];
        kwargs: the default geometry (config['pruning']['args'])."""
        for spec in blocks:
            name = spec['name']
            self.replaced_block_names.append(name)
            old = self.get_block(name, self.teacher)
            geom = spec['args'] if 'args' in spec else kwargs
            new = DepthwiseSeparableBlock(in_channels=old.in_channels, out_channels=old.out_channels,
                                          kernel_size=geom['kernel_size'], padding=geom['padding'],
                                          dilation=geom['dilation'], groups=old.in_channels, bias=old.bias)
            # the reference calls .cuda(); following the replaced block's device is the same thing on a GPU box
            new = new.to(next(old.parameters()).device)
            self._set_block(name, new, self.student)
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def _remove_hooks(self):
        for handlers in (self._student_hook_handlers, self._teacher_hook_handlers):
            while handlers:
                handlers.pop().remove()

    def register_hint_layers(self, block_names):
        """Forward hooks that collect the named blocks' outputs of teacher and student (by reference: later
        in-place edits of the tensor, e.g. the residual add_, are visible in the hint -- SURVEY.md F10)."""
        if len(block_names) > 0:
            self._remove_hooks()
        for name in block_names:
            self.aux_block_names.append(name)

            def keep_teacher(module, inputs, output):
                if self.save_hidden:
                    self.teacher_hidden_outputs.append(output)

            def keep_student(module, inputs, output):
                if self.save_hidden:
                    self.student_hidden_outputs.append(output)

            self._teacher_hook_handlers.append(self.get_block(name, self.teacher).register_forward_hook(keep_teacher))
            self._student_hook_handlers.append(self.get_block(name, self.student).register_forward_hook(keep_student))
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def unfreeze(self, block_names):
        for name in block_names:
            for prm in self.get_block(name, self.student).parameters():
                prm.requires_grad = True

    def reset(self):
        self._remove_hooks()
        while self.replaced_block_names:
            name = self.replaced_block_names.pop()
            self._set_block(name, copy.deepcopy(self.get_block(name, self.teacher)), self.student)

    # ---- forward -----------------------------------------------------------------------------------------
    def forward(self, x):
        self.student_hidden_outputs = []
        self.teacher_hidden_outputs = []
        with torch.no_grad():
            teacher_pred = self.teacher(x)
        student_pred = self.student(x)
        return student_pred, teacher_pred

    def inference(self, x):
        self.student_hidden_outputs = []
        self.teacher_hidden_outputs = []
        return self.student(x)

    def inference_test(self, data, args):
        """Sliding-window multi-scale + mirrored inference of the student (depthwise_student.py:187-206); the windows
        are stitched on the device with the reference's normalisation (kdcc.tta.reverse_mapping)."""
        from . import tta
        self.student_hidden_outputs = []
        self.teacher_hidden_outputs = []
        return tta.inference_test(self.student, data, args)

    def train(self, mode=True):
        self.save_hidden = bool(mode)
        super().train(mode)
        self.teacher.eval()  # the teacher never leaves eval mode
        return self

    # ---- reports -----------------------------------------------------------------------------------------
    def trainable_parameters(self):
        return [p for p in self.student.parameters() if p.requires_grad]

    def dump_trainable_params(self):
        return '\nTrainable parameters: {}'.format(sum(p.numel() for p in self.parameters() if p.requires_grad))

    def dump_student_teacher_blocks_info(self):
        lines = ["block | teacher params | student params"]
        for name in self.replaced_block_names:
            t = sum(p.numel() for p in self.get_block(name, self.teacher).parameters())
            s = sum(p.numel() for p in self.get_block(name, self.student).parameters())
            lines.append("{} | {} | {}".format(name, t, s))
        return "\n".join(lines)

    def __str__(self):
        return super().__str__() + '\n' + self.dump_student_teacher_blocks_info()
