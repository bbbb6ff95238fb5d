"""Host-side mirror of the reference's student wrapper / layerwise loop: surgery API, checkpoint keys, hooks,
flat gradient bucket, device-side confusion matrix (CPU), the 2-rank gloo gradient all-reduce, and -- on a
GPU -- one full LayerwiseStep against the same step computed with the oracle's torch-CPU port."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

import kdcc
from oracle import torch_port as tp


class TinyTeacher(nn.Module):
    """Shaped like the real teachers where it matters: named/indexed sub-blocks with bias-free 3x3 convs."""

    def __init__(self, classes=19):
        super().__init__()
        self.stem = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        self.body = nn.Sequential(nn.Conv2d(16, 32, 3, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU(),
                                  nn.Conv2d(32, 32, 3, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU())
        self.head = nn.Conv2d(32, classes, 1, bias=False)

    def forward(self, x):
        return self.head(self.body(self.stem(x)))


PLAN = [{"name": "body.0", "epoch": 1}, {"name": "body.3", "epoch": 1, "args": {"kernel_size": 3, "padding": 2, "dilation": 2}}]
GEOM = {"kernel_size": 3, "padding": 1, "dilation": 1}


def make_student(device="cpu"):
    torch.manual_seed(0)
    teacher = TinyTeacher().to(device)
    st = kdcc.DepthwiseStudent(teacher, {"trainer": {"verbosity": 2}})
    st.replace(PLAN, **GEOM)
    st.register_hint_layers([b["name"] for b in PLAN])
    st.unfreeze([b["name"] for b in PLAN])
    return st


def test_surgery_keeps_reference_checkpoint_layout():
    st = make_student()
    assert st.replaced_block_names == ["body.0", "body.3"]
    assert isinstance(st.get_block("body.0", st.student), kdcc.DepthwiseSeparableBlock)
    assert isinstance(st.get_block("body.0", st.teacher), nn.Conv2d)
    keys = set(st.state_dict())
    for k in ("teacher.body.0.weight", "student.body.0.separable_conv.weight", "student.body.0.pointwise_conv.weight",
              "student.body.3.separable_conv.weight", "student.body.3.pointwise_conv.weight", "student.stem.weight"):
        assert k in keys
    assert tuple(st.state_dict()["student.body.0.separable_conv.weight"].shape) == (16, 1, 3, 3)
    assert tuple(st.state_dict()["student.body.3.pointwise_conv.weight"].shape) == (32, 32, 1, 1)
    # per-block geometry override ("args") is honoured
    blk = st.get_block("body.3", st.student)
    assert blk.separable_conv.dilation == (2, 2) and blk.separable_conv.padding == (2, 2)
    # only the new blocks train; teacher stays frozen and in eval mode even after .train()
    assert sum(p.numel() for p in st.trainable_parameters()) == 16 * 9 + 32 * 16 + 32 * 9 + 32 * 32
    st.train()
    assert not st.teacher.training and st.save_hidden
    st.train(False)
    assert not st.save_hidden
    # a state dict written by the reference layout loads back
    other = make_student()
    other.load_state_dict(copy.deepcopy(st.state_dict()))
    st.reset()
    assert st.replaced_block_names == [] and isinstance(st.get_block("body.0", st.student), nn.Conv2d)


def test_grad_bucket_views_and_confusion_matrix():
    st = make_student()
    bucket = kdcc.GradBucket(st.trainable_parameters())
    assert bucket.flat.numel() == sum(-(-p.numel() // 32) * 32 for p in st.trainable_parameters())
    assert all((p.grad.data_ptr() - bucket.flat.data_ptr()) % 128 == 0 for p in bucket.params)
    for p in bucket.params:
        assert p.grad.data_ptr() >= bucket.flat.data_ptr()
        p.grad.fill_(1.0)
    assert float(bucket.flat.sum()) == sum(p.numel() for p in bucket.params) == bucket.dense().numel()   # padding stays zero
    bucket.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in bucket.params)


def _ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1)
    model = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
    x_all, y_all = torch.randn(8, 6), torch.randn(8, 3)
    bucket = kdcc.GradBucket(model.parameters())
    xs, ys = x_all.chunk(world)[rank], y_all.chunk(world)[rank]
    ((model(xs) - ys) ** 2).mean().backward()      # local mean over the shard
    bucket.all_reduce_mean()
    if rank == 0:
        ref = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        ((ref(x_all) - y_all) ** 2).mean().backward()   # mean over the global batch
        ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        out.put(float((bucket.dense() - ref_flat).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_global_batch_gradient():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-6


class RefBlock(nn.Module):
    """The reference block restated with the oracle's torch port (CPU checker for the GPU test)."""

    def __init__(self, blk):
        super().__init__()
        self.w_dw = nn.Parameter(blk.separable_conv.weight.detach().cpu().clone())
        self.w_pw = nn.Parameter(blk.pointwise_conv.weight.detach().cpu().clone())
        self.p, self.d = blk.separable_conv.padding[0], blk.separable_conv.dilation[0]

    def forward(self, x):
        return tp.block_forward(x, self.w_dw, self.w_pw, self.p, self.d)


@pytest.mark.gpu
def test_layerwise_step_matches_torch_port_on_cpu():
    # the frozen convs around the blocks are stock torch: keep them in true fp32 so the comparison is about kdcc
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    st = make_student("cuda")
    names = [b["name"] for b in PLAN]
    # CPU checker: same teacher/student weights, blocks restated with the torch port
    ref = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    ref.teacher.load_state_dict({k: v.cpu() for k, v in st.teacher.state_dict().items()})
    ref.student.load_state_dict({k: v.cpu() for k, v in st.teacher.state_dict().items()})
    for n in names:
        ref._set_block(n, RefBlock(st.get_block(n, st.student)), ref.student)
    ref.register_hint_layers(names)

    torch.manual_seed(3)
    data = torch.randn(2, 3, 24, 32)
    target = torch.randint(0, 19, (2, 24, 32))
    crit = [nn.CrossEntropyLoss(ignore_index=255), kdcc.KLDivergenceLoss(temperature=2), kdcc.MSELoss(num_classes=1000)]
    opt = torch.optim.SGD(st.trainable_parameters(), lr=0.0)
    step = kdcc.LayerwiseStep(st, crit, opt, accumulation_steps=2)
    out = step(data.cuda(), target.cuda(), batch_idx=1)   # idx 1 with 2 accumulation steps: no optimizer step yet
    assert out["loss"].dim() == 0 and out["loss"].is_cuda

    s_ref, t_ref = ref(data)
    hint = sum(tp.mse_loss(a, b, 1000) for a, b in zip(ref.student_hidden_outputs, ref.teacher_hidden_outputs)) / 2
    hint.backward()
    kd_ref = tp.kl_div_loss(s_ref, t_ref, 2.0) / 2
    assert abs(float(out["hint_loss"]) - float(hint)) <= 1e-4 * abs(float(hint)), (float(out["hint_loss"]), float(hint))
    assert abs(float(out["kd_loss"]) - float(kd_ref)) <= 1e-4 * abs(float(kd_ref)), (float(out["kd_loss"]), float(kd_ref))
    for n in names:
        mine, theirs = st.get_block(n, st.student), ref.get_block(n, ref.student)
        for g_mine, g_ref in ((mine.separable_conv.weight.grad, theirs.w_dw.grad), (mine.pointwise_conv.weight.grad, theirs.w_pw.grad)):
            err = float((g_mine.cpu() - g_ref).abs().max() / g_ref.abs().max())
            assert err < 1e-4, (n, err)
    # gradients live in the flat bucket (what the NCCL all-reduce would touch)
    assert float(step.bucket.flat.abs().sum()) > 0
    cm = kdcc.ConfusionMatrix(19, 255, device="cuda")
    cm.update(out["output_st"], target.cuda())
    assert int(cm.mat.sum()) == target.numel()


class _PairModel(nn.Module):
    """Stand-in for a DepthwiseStudent on the CPU: one trainable linear 'student', one frozen 'teacher', one hint pair."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.student, self.teacher = nn.Linear(6, 4), nn.Linear(6, 4)
        for prm in self.teacher.parameters():
            prm.requires_grad = False
        self.student_hidden_outputs, self.teacher_hidden_outputs = [], []

    def trainable_parameters(self):
        return [prm for prm in self.student.parameters() if prm.requires_grad]

    def forward(self, x):
        s, t = self.student(x), self.teacher(x)
        self.student_hidden_outputs, self.teacher_hidden_outputs = [s], [t]
        return s, t


def test_classification_step_backpropagates_kd_only_and_steps_on_the_reference_cadence():
    """trainer/classification_trainer.py:24-43: loss = kd_loss; step when (batch_idx + 1) % accumulation_steps == 0."""
    model = _PairModel()
    crit = [lambda out, tgt: nn.functional.cross_entropy(out, tgt),
            lambda s, t: ((s - t) ** 2).mean(),          # stands for the KD criterion
            lambda s, t: (s - t).abs().mean() * 7.0]     # hint criterion: must NOT reach the gradient
    opt = torch.optim.SGD(model.trainable_parameters(), lr=0.1)
    step = kdcc.ClassificationStep(model, crit, opt, accumulation_steps=2)
    torch.manual_seed(4)
    xs, ys = [torch.randn(5, 6) for _ in range(4)], [torch.randint(0, 4, (5,)) for _ in range(4)]
    # the same four batches by hand
    ref = _PairModel()
    ropt = torch.optim.SGD(ref.trainable_parameters(), lr=0.1)
    w0 = model.student.weight.detach().clone()
    for i, (x, y) in enumerate(zip(xs, ys)):
        out = step(x, y, i)
        s, t = ref(x)
        (((s - t) ** 2).mean() / 2).backward()
        if (i + 1) % 2 == 0:
            ropt.step()
            ropt.zero_grad()
        assert torch.allclose(out["kd_loss"], ((s - t) ** 2).mean() / 2)
        assert torch.allclose(out["hint_loss"], (s - t).abs().mean() * 7.0 / 2) and not out["hint_loss"].requires_grad
        assert out["loss"] is out["kd_loss"]
        if i == 0:
            assert torch.equal(model.student.weight.detach(), w0)   # no step on index 0 (the layerwise loop would)
        assert torch.allclose(model.student.weight, ref.student.weight, atol=1e-7)
    assert not torch.equal(model.student.weight.detach(), w0)


def test_ensemble_step_composes_the_kd_term_like_the_reference():
    """trainer/ensemble_trainer.py:73-90: kd = (sum_k W * kd(s, m_k(x)) + kd(s, teacher)) / (W * K + 1); loss = kd + sup."""
    model = _PairModel()
    torch.manual_seed(9)
    others = [nn.Linear(6, 4) for _ in range(3)]
    sup = lambda out, tgt: nn.functional.cross_entropy(out, tgt)
    kd = lambda s, t: ((s - t) ** 2).mean()
    opt = torch.optim.SGD(model.trainable_parameters(), lr=0.05)
    step = kdcc.EnsembleStep(model, others, [sup, kd], opt, accumulation_steps=1, weight=2)
    ref = _PairModel()
    ropt = torch.optim.SGD(ref.trainable_parameters(), lr=0.05)
    for i in range(3):
        x, y = torch.randn(5, 6), torch.randint(0, 4, (5,))
        out = step(x, y, i)
        s, t = ref(x)
        with torch.no_grad():
            outs = [m(x) for m in others]
        want_kd = (sum(2 * kd(s, o) for o in outs) + kd(s, t)) / (2 * 3 + 1)
        (want_kd + sup(s, y)).backward()
        ropt.step()
        ropt.zero_grad()
        assert torch.allclose(out["kd_loss"], want_kd) and torch.allclose(out["loss"], want_kd + sup(s, y))
        assert torch.allclose(model.student.weight, ref.student.weight, atol=1e-7)
    # a fused multi-teacher criterion plugs in with the same call shape
    calls = []
    fused = lambda s, outs, t: (calls.append(len(outs)) or sum(kd(s, o) for o in outs) + kd(s, t))
    kdcc.EnsembleStep(model, others, [sup, kd], opt, kd_multi=fused)(torch.randn(5, 6), torch.randint(0, 4, (5,)), 0)
    assert calls == [3]


def test_prepare_train_epoch_follows_the_reference_schedule():
    """trainer/layerwise_trainer.py:78-186: surgery only at the epochs the config names, a NEW optimizer at epoch 1,
    param groups (with the sticky layerwise lr) afterwards, the flat gradient bucket rebuilt with the trainable set."""
    torch.manual_seed(0)
    st = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    pruning = {"args": dict(GEOM),
               "pruning_plan": [{"name": "body.0", "epoch": 1}, {"name": "body.3", "epoch": 3, "args": {"kernel_size": 3, "padding": 2, "dilation": 2}}],
               "hint": [{"name": "body.0", "epoch": 1}, {"name": "body.3", "epoch": 3}],
               "unfreeze": [{"name": "body.0", "epoch": 1}, {"name": "body.3", "epoch": 3, "lr": 0.5}, {"name": "head", "epoch": 3}]}
    made = []
    make = lambda params: (made.append(len(params)) or torch.optim.SGD(params, lr=0.1))
    opt_args = {"lr": 0.1}
    opt = kdcc.prepare_train_epoch(st, pruning, 1, None, make, opt_args)
    assert made == [2] and st.replaced_block_names == ["body.0"]           # dw + pw weight of the one replaced block
    assert isinstance(st.get_block("body.0", st.student), kdcc.DepthwiseSeparableBlock)
    assert isinstance(st.get_block("body.3", st.student), nn.Conv2d)       # not yet
    step = kdcc.LayerwiseStep(st, [None, None, None], opt)
    n1 = step.bucket.flat.numel()
    assert kdcc.prepare_train_epoch(st, pruning, 2, opt, make, opt_args, step=step) is opt and made == [2]   # nothing named at epoch 2
    assert step.bucket.flat.numel() == n1
    opt3 = kdcc.prepare_train_epoch(st, pruning, 3, opt, make, opt_args, step=step)
    assert opt3 is opt and made == [2] and len(opt.param_groups) == 3      # same optimizer, one group per unfrozen layer
    assert st.replaced_block_names == ["body.0", "body.3"]
    blk = st.get_block("body.3", st.student)
    assert blk.separable_conv.dilation == (2, 2)
    assert opt.param_groups[1]["lr"] == 0.5 and opt.param_groups[2]["lr"] == 0.5   # the override sticks (reference :171-172)
    assert opt_args["lr"] == 0.5
    assert step.bucket.flat.numel() == n1 + sum(p.numel() for p in blk.parameters()) + st.get_block("head", st.student).weight.numel()
    assert len(st.student_hidden_outputs) == 0
    # a plan with nothing in it: the whole student trains
    st2 = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    kdcc.prepare_train_epoch(st2, {"args": dict(GEOM), "pruning_plan": [], "hint": [], "unfreeze": []}, 1, None, make)
    assert all(p.requires_grad for p in st2.student.parameters()) and made[-1] == len(list(st2.student.parameters()))
