"""RAdam as the reference implements it (utils/optim/radam.py), one fused kernel per parameter tensor.

Same constructor, `param_groups` / `state` layout (`step`, `exp_avg`, `exp_avg_sq`) and step-size schedule as the
reference class -- so `optimizer.state_dict()` checkpoints (base/base_trainer.py:162-185) load either way -- but the
elementwise body (radam.py:41-97: ~10 torch kernels per tensor) is `kdcc_radam_step`: p, g, m, v are read once and
p, m, v written once.  `attach_lp_copy(param, tensor)` makes the same pass also write the bf16 copy of the new
weights (what the pointwise GEMMs of the next step read).
"""
import math

import torch
from torch.optim.optimizer import Optimizer

from . import _abi
from .functional import lp_copy_of, lp_refreshed


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        # same argument checks (ValueError) as radam.py:9-16
        for ok, what, val in ((lr >= 0.0, "learning rate", lr), (eps >= 0.0, "epsilon value", eps),
                              (0.0 <= betas[0] < 1.0, "beta parameter at index 0", betas[0]),
                              (0.0 <= betas[1] < 1.0, "beta parameter at index 1", betas[1])):
            if not ok:
                raise ValueError("Invalid %s: %s" % (what, val))
        self.degenerated_to_sgd = degenerated_to_sgd
        # `buffer` only exists so that state_dict()s are interchangeable with the reference's (radam.py:19-24 caches
        # the step scalars there; they are recomputed here, which is a few host flops per tensor)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, buffer=[[None, None, None] for _ in range(10)])
        super().__init__(params, defaults)
        self._lp = {}
        self._sources = {}

    def attach_lp_copy(self, param, lp_tensor):
        """`lp_tensor` (bf16, same numel, contiguous) is rewritten with the new value of `param` by every step."""
        if lp_tensor.dtype != torch.bfloat16 or lp_tensor.numel() != param.numel() or not lp_tensor.is_contiguous():
            raise _abi.KdccError("the low-precision copy must be a contiguous bf16 tensor of the parameter's size")
        self._lp[param] = lp_tensor

    def attach_grad_sources(self, param, sources):
        """Data-parallel form: `sources()` -> (view of source 0's copy of this parameter's gradient, floats between sources,
        number of sources); the step then uses the mean of the copies (kdcc_radam_step_multi), i.e. the gradient all-reduce
        is fused into the optimizer pass.  See kdcc.PeerGradBucket, which fills the copies over NVLink."""
        self._sources[param] = sources

    def step_scalars(self, step, beta1, beta2):
        """(N_sma, step_size) of radam.py:65-84 (host float64, as there)."""
        decay2 = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * decay2 / (1 - decay2)
        bias1 = 1 - beta1 ** step
        if n_sma >= 5:   # variance rectification term, then Adam's first-moment bias correction
            rect = (1 - decay2) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)
            size = math.sqrt(rect) / bias1
        else:
            size = 1.0 / bias1 if self.degenerated_to_sgd else -1
        return n_sma, size

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _abi.lib()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                if not (p.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()):
                    raise _abi.KdccError("kdcc.optim.RAdam steps contiguous fp32 CUDA parameters (master weights); got %s %s"
                                         % (p.dtype, p.device))
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state['step'] += 1
                n_sma, size = self.step_scalars(state['step'], beta1, beta2)
                mode = 0 if n_sma >= 5 else (1 if size > 0 else 2)
                lp = self._lp.get(p)
                cached = None
                if lp is None:   # a bf16 copy some forward made of this master weight: refresh it in the same pass
                    lp = cached = lp_copy_of(p)
                tail = (lp.data_ptr() if lp is not None else None, p.numel(), beta1, beta2, 1 - beta1, 1 - beta2, group['eps'],
                        -group['weight_decay'] * group['lr'], -size * group['lr'], mode, torch.cuda.current_stream(p.device).cuda_stream)
                src = self._sources.get(p)
                if src is None:
                    _abi.check(L.kdcc_radam_step(p.data_ptr(), g.data_ptr(), state['exp_avg'].data_ptr(), state['exp_avg_sq'].data_ptr(),
                                                 *tail), "kdcc_radam_step")
                else:
                    g0, stride, n_src = src()
                    _abi.check(L.kdcc_radam_step_multi(p.data_ptr(), g0.data_ptr(), int(stride), int(n_src), state['exp_avg'].data_ptr(),
                                                       state['exp_avg_sq'].data_ptr(), *tail), "kdcc_radam_step_multi")
                # the kernel wrote p through a raw pointer: move its version counter like an in-place op would, so autograd's
                # saved-tensor checks and the weight-copy cache see the update
                torch.autograd.graph.increment_version(p)
                if cached is not None:
                    lp_refreshed(p)
        return loss
