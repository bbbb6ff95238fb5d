"""The distillation hot path as one fused, autograd-free pass over pre-allocated buffers.

This is the body of the reference's layerwise training loop restricted to the rows SURVEY.md §8 puts on
the path (trainer/layerwise_trainer.py:223-239): for every replaced site run the cheap-conv block
forward, the hint loss against the teacher feature (which emits d loss / d y in the same pass),
the block backward (pointwise dW + dX, depthwise dW + dX), then the logits KD loss, the student-gradient
all-reduce and the optimizer step.  Every arithmetic step is a libkdcc.so call on raw pointers; torch
only owns the memory, the stream, the NCCL all-reduce and the optimizer.

`HotPathStep` is what bench.py times and what `LayerwiseStep` (trainer.py) uses for the loss side.
"""
import math

import torch

from . import _abi
from .functional import _ptr, _stream

# (C_in, C_out) of the sites replaced by the shipped Cityscapes plans, all on 128x128 maps for a 1024^2 crop
PLAN_51M_DEEPLAB = [(512, 512)] * 5 + [(1024, 2048)] + [(4096, 256)] * 3      # cfg/cityscapes/51M_deeplab_all.json:123-160
PLAN_58M_DEEPLAB = [(512, 512)] * 6 + [(512, 1024), (1024, 2048)] + [(4096, 256)] * 3  # cfg/cityscapes/58M_deeplab_all.json
PLAN_CIFAR_RESNET44 = [(64, 64)] * 8                                           # cfg/cifar10/resnet44/config1.json


class EventLog:
    """CUDA-event timeline on the launching stream: one event after every kernel call.

    `only`: record just these families (an event before and one after each of their calls) and nothing else.  An event
    record between two kernels ends the programmatic-dependent-launch overlap of their tail and prologue, so a full
    timeline costs the step ~4 %; bench.py therefore brackets only the dominant family inside its timed region and takes
    the complete per-kernel table from a separate instrumented pass."""

    def __init__(self, external=False, only=None):
        # external=True: events that may be recorded inside a CUDA-graph capture (they become event-record nodes and
        # are re-recorded by every replay, so the log then holds the timeline of the LAST replay)
        self.names, self.events, self.external = [], [], external
        self.only = set(only) if only else None

    def pre(self, name):
        """Start-of-call mark; needed only when the log is selective (otherwise the previous call's mark is the start)."""
        if self.only is not None and name in self.only:
            self._record("begin")

    def mark(self, name):
        if self.only is not None and name not in self.only:
            return
        self._record(name)

    def _record(self, name):
        ev = torch.cuda.Event(enable_timing=True, external=True) if self.external else torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        self.names.append(name)
        self.events.append(ev)

    def durations_ms(self):
        """{name: [ms, ...]} -- each call's time is the gap to the previous mark ('begin' marks reset it)."""
        out = {}
        for i in range(1, len(self.events)):
            if self.names[i] == "begin":
                continue
            out.setdefault(self.names[i], []).append(self.events[i - 1].elapsed_time(self.events[i]))
        return out


class HotPathStep:
    def __init__(self, plan, batch, height, width, kernel_size=9, dilation=5, padding=20, dtype=torch.bfloat16,
                 device="cuda", logits_shape=None, kd_temperature=1.0, hint_num_classes=1000.0,
                 accumulation_steps=1, kd_grad=False, need_dx=True, seed=0, layout="nchw", hint_weighted=False,
                 order="reference"):
        self.plan = list(plan)
        self.N, self.H, self.W = batch, height, width
        self.k, self.d, self.p = kernel_size, dilation, padding
        self.Ho = height + 2 * padding - dilation * (kernel_size - 1)
        self.Wo = width + 2 * padding - dilation * (kernel_size - 1)
        self.dtype, self.device = dtype, torch.device(device)
        self.code = _abi.F32 if dtype == torch.float32 else _abi.BF16
        self.T, self.nc, self.acc_steps = float(kd_temperature), float(hint_num_classes), int(accumulation_steps)
        self.kd_grad = kd_grad
        # per site: does the block's input need a gradient?  (a bool applies to every site; the first trainable conv of a
        # network, fed by frozen layers only, does not -- SURVEY.md 3.4)
        self.need_dx = [bool(need_dx)] * len(self.plan) if isinstance(need_dx, bool) else [bool(v) for v in need_dx]
        # "nchw" = the reference's layout -> tensor-core depthwise + W.X pointwise; "nhwc" = channels-last kernels
        #   "nhwc" with a kernel larger than 3x3: channels_last INPUTS and OUTPUTS (what a cuDNN channels_last trunk hands over
        #   and expects back), re-laid to channel planes at the block boundary so the depthwise stays on the tensor cores;
        #   "nhwc_native": the NHWC CUDA-core kernels throughout
        self.io_layout = _abi.NCHW if layout == "nchw" else _abi.NHWC
        self.relayout = layout == "nhwc" and kernel_size > 3 and dtype == torch.bfloat16 and (height * width) % 8 == 0 and \
            all(ci % 8 == 0 and co % 8 == 0 for ci, co in self.plan)
        self.layout = _abi.NCHW if (layout == "nchw" or self.relayout) else _abi.NHWC
        self.L = _abi.lib()
        kk = kernel_size * kernel_size

        # flat fp32 parameter / gradient buckets: [dw weights of site 0 | pw weights of site 0 | site 1 ...]
        sizes = []
        for ci, co in self.plan:
            sizes += [ci * kk, co * ci]
        total = sum(sizes)
        gen = torch.Generator(device="cpu").manual_seed(seed)
        flat = torch.empty(total, dtype=torch.float32)
        off = 0
        self._views = []
        for (ci, co) in self.plan:
            # nn.Conv2d default init bounds: U(+-1/sqrt(fan_in)), fan_in = k*k (depthwise) or C_in (pointwise)
            flat[off:off + ci * kk].uniform_(-1.0 / kernel_size, 1.0 / kernel_size, generator=gen)
            flat[off + ci * kk:off + ci * kk + co * ci].uniform_(-1.0 / math.sqrt(ci), 1.0 / math.sqrt(ci), generator=gen)
            self._views.append((off, off + ci * kk, off + ci * kk + co * ci))
            off += ci * kk + co * ci
        self.flat_params = flat.to(self.device)
        self.flat_grads = torch.zeros_like(self.flat_params)
        self.num_trainable = total

        cmax = max(ci for ci, _ in self.plan)
        omax = max(co for _, co in self.plan)
        n = batch
        if self.layout == _abi.NCHW:
            e = lambda c, h, w: torch.empty((n, c, h, w), dtype=dtype, device=self.device)
        else:
            e = lambda c, h, w: torch.empty((n, h, w, c), dtype=dtype, device=self.device)
        # scratch shared by all sites (each site's forward+backward completes before the next starts)
        self.mid = e(cmax, self.Ho, self.Wo)
        self.dmid = e(cmax, self.Ho, self.Wo)
        self.dx = e(cmax, height, width) if any(self.need_dx) else None
        self.y = e(omax, self.Ho, self.Wo)
        self.dy = e(omax, self.Ho, self.Wo)
        if self.relayout:  # channel-plane copies of what arrives / leaves channels_last
            self.x_planes = e(cmax, height, width)
            self.dx_cl = e(cmax, height, width) if any(self.need_dx) else None
        # order "reference": forward of every site, the hint losses, then the backward in reverse site order -- what
        # `model(data)` ... `loss.backward()` does -- so every site keeps its depthwise output, block output and gradient
        # until its backward ("interleaved": forward + loss + backward site by site over shared scratch; 5 % slower)
        self.order = order
        if order == "reference":
            es = (lambda c, h, w: torch.empty((n, c, h, w), dtype=dtype, device=self.device))
            self.mid_s = [es(ci, self.Ho, self.Wo) for ci, _ in self.plan]
            self.y_s = [es(co, self.Ho, self.Wo) for _, co in self.plan]
            self.dy_s = [es(co, self.Ho, self.Wo) for _, co in self.plan]
            self.xp_s = [es(ci, height, width) for ci, _ in self.plan] if self.relayout else None
        # bf16 copy of the WHOLE flat parameter bucket, refreshed by one cast launch per step (the pointwise GEMMs read
        # their weights from it; one launch instead of one per site)
        self.flat_lp = torch.empty(total, dtype=dtype, device=self.device)
        ws_bytes = self.L.kdcc_loss_workspace_bytes()
        for ci, co in self.plan:
            M = n * self.Ho * self.Wo
            ws_bytes = max(ws_bytes, self.L.kdcc_dw_bwd_workspace_bytes(n, height, width, ci, self.k, self.d, self.p, self.layout, self.code),
                           self.L.kdcc_pw_bwd_workspace_bytes(1, M, ci, co, self.code))
        self.ws = torch.empty(ws_bytes + 64, dtype=torch.uint8, device=self.device)
        self.hint_losses = torch.zeros(len(self.plan), dtype=torch.float32, device=self.device)
        # hint_weighted: losses/WeightedHintMSELoss.py (per-channel filter weights, e.g. normalised Taylor importances --
        # utils/util.py:191-216) instead of losses/MSELoss.py * num_classes
        self.hint_w = None
        if hint_weighted:
            gw = torch.Generator(device="cpu").manual_seed(seed + 17)
            self.hint_w = [torch.rand(co, generator=gw).to(self.device) for _, co in self.plan]
            ws_bytes += 4 * batch * omax
            self.ws = torch.empty(ws_bytes + 64, dtype=torch.uint8, device=self.device)
        self.kd_loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.logits_shape = logits_shape
        self.dlogits = torch.empty(logits_shape, dtype=torch.float32, device=self.device) if (logits_shape and kd_grad) else None
        self.launches_per_step = 0
        # True once something else keeps flat_lp equal to bf16(flat_params) -- kdcc.optim.RAdam.attach_lp_copy writes it
        # in the optimizer pass -- so the step does not need its own cast launch (refresh_lp() once before the first step)
        self.lp_maintained = False

    # ---- synthetic inputs of the right shapes (the frozen trunk that would produce them is out of scope) ----
    def make_inputs(self, seed=1, pinned_host=False):
        g = torch.Generator(device="cpu").manual_seed(seed)
        dev = "cpu" if pinned_host else self.device

        def rnd(shape, dtype, scale=1.0):
            t = (torch.randn(shape, generator=g) * scale).to(dtype)
            return t.pin_memory() if pinned_host else t.to(self.device)

        if self.io_layout == _abi.NCHW:
            xs = [rnd((self.N, ci, self.H, self.W), self.dtype) for ci, _ in self.plan]
            ts = [rnd((self.N, co, self.Ho, self.Wo), self.dtype) for _, co in self.plan]
        else:
            xs = [rnd((self.N, self.H, self.W, ci), self.dtype) for ci, _ in self.plan]
            ts = [rnd((self.N, self.Ho, self.Wo, co), self.dtype) for _, co in self.plan]
        ls = lt = None
        if self.logits_shape:
            ls, lt = rnd(self.logits_shape, torch.float32, 3.0), rnd(self.logits_shape, torch.float32, 3.0)
        return xs, ts, ls, lt

    def refresh_lp(self):
        """bf16 copy of the whole flat parameter bucket, now (one cast launch)."""
        if self.code == _abi.BF16:
            _abi.check(self.L.kdcc_cast_f32_to_bf16(_ptr(self.flat_params), _ptr(self.flat_lp), self.flat_params.numel(), _stream()), "cast")

    def _site_weights(self, i):
        a, b, c = self._views[i]
        return self.flat_params[a:b], self.flat_params[b:c], self.flat_grads[a:b], self.flat_grads[b:c]

    def step(self, xs, teacher_feats, logits_s=None, logits_t=None, log=None, site_done=None):
        """One pass; returns (hint_loss_sum, kd_loss) as 0-dim device tensors (no host sync).  `site_done(lo, hi)` is called
        when the gradients flat_grads[lo:hi] of a site are final (its weight-gradient kernels are enqueued): the hook of the
        data-parallel gradient exchange, which overlaps the transfer with the remaining sites."""
        L, st, code, n, lay = self.L, _stream(), self.code, self.N, self.layout
        H, W, Ho, Wo, k, d, p = self.H, self.W, self.Ho, self.Wo, self.k, self.d, self.p
        chk = _abi.check
        ws, wsn = _ptr(self.ws), self.ws.numel()
        mark = log.mark if log is not None else (lambda name: None)
        pre = log.pre if log is not None else (lambda name: None)
        launches = 0
        mark("begin")
        if code == _abi.BF16 and not self.lp_maintained:
            chk(L.kdcc_cast_f32_to_bf16(_ptr(self.flat_params), _ptr(self.flat_lp), self.flat_params.numel(), st), "cast")
            mark("cast_w")
            launches += 1
        rl = self.relayout
        play = _abi.PLANES_TO_NHWC if rl else lay      # channels_last callers: the GEMMs read / write y and dy channels_last
        M = n * Ho * Wo
        per_site = self.order == "reference"

        def buffers(i):
            return (self.mid_s[i], self.y_s[i], self.dy_s[i], self.xp_s[i] if rl else None) if per_site else \
                   (self.mid, self.y, self.dy, self.x_planes if rl else None)

        def weights(i):
            w_dw, w_pw, g_dw, g_pw = self._site_weights(i)
            w_lp = self.flat_lp[self._views[i][1]:self._views[i][2]] if code == _abi.BF16 else w_pw
            return w_dw, w_lp, g_dw, g_pw

        def forward(i):
            nonlocal launches
            ci, co = self.plan[i]
            w_dw, w_lp, _, _ = weights(i)
            mid, y, dy, xp = buffers(i)
            x = xs[i]
            if rl:   # channels_last input -> channel planes
                chk(L.kdcc_layout_convert(_ptr(x), _ptr(xp), n, ci, H * W, 1, code, st), "relayout")
                x = xp
                mark("relayout")
                launches += 1
            pre("dw_fwd")
            chk(L.kdcc_dw_fwd(_ptr(x), _ptr(w_dw), None, _ptr(mid), n, H, W, ci, k, d, p, lay, code, st), "dw_fwd")
            mark("dw_fwd")
            chk(L.kdcc_pw_fwd(_ptr(mid), _ptr(w_lp), None, None, 0, _ptr(y), None, M, ci, co, n, play, code, st), "pw_fwd")
            mark("pw_fwd")
            launches += 2

        def hint(i):
            nonlocal launches
            ci, co = self.plan[i]
            mid, y, dy, xp = buffers(i)
            chk(L.kdcc_hint_loss(_ptr(y), _ptr(teacher_feats[i]), _ptr(self.hint_w[i]) if self.hint_w else None, 0, _ptr(dy),
                                 _ptr(self.hint_losses[i:]), ws, wsn, n, co, Ho * Wo, self.io_layout,
                                 1.0 if self.hint_w else self.nc, code, 1.0 / self.acc_steps, st), "hint_loss")
            mark("hint_loss")
            launches += 2   # pass + finalize

        def backward(i):
            nonlocal launches
            ci, co = self.plan[i]
            w_dw, w_lp, g_dw, g_pw = weights(i)
            mid, y, dy, xp = buffers(i)
            x = xp if rl else xs[i]
            chk(L.kdcc_pw_bwd_dw(_ptr(dy), _ptr(mid), _ptr(g_pw), ws, wsn, M, ci, co, n, play, code, st), "pw_bwd_dw")
            mark("pw_bwd_dw")
            chk(L.kdcc_pw_bwd_dx(_ptr(dy), _ptr(w_lp), _ptr(self.dmid), None, 0, M, ci, co, n, play, code, st), "pw_bwd_dx")
            mark("pw_bwd_dx")
            pre("dw_bwd")
            chk(L.kdcc_dw_bwd(_ptr(x), _ptr(w_dw), _ptr(self.dmid), _ptr(self.dx) if self.need_dx[i] else None, _ptr(g_dw),
                              None, ws, wsn, n, H, W, ci, k, d, p, lay, code, st), "dw_bwd")
            mark("dw_bwd")
            if site_done is not None:
                site_done(self._views[i][0], self._views[i][2])
            if rl and self.need_dx[i]:   # the input gradient goes back channels_last
                chk(L.kdcc_layout_convert(_ptr(self.dx), _ptr(self.dx_cl), n, ci, H * W, 0, code, st), "relayout")
                mark("relayout")
                launches += 1
            # pw_dw 2 (gemm + split reduce), pw_dx 1, dw_bwd: wgrad 1 + reduce 1 (+ dx conv 1)
            launches += 2 + 1 + (3 if self.need_dx[i] else 2)

        sites = range(len(self.plan))
        if per_site:
            # the reference loop's order (trainer/layerwise_trainer.py:223-235): the student's forward pass visits every
            # block, the hint criterion is evaluated over the hooked pairs, then loss.backward() walks the blocks in reverse
            for i in sites:
                forward(i)
            for i in sites:
                hint(i)
            for i in reversed(sites):
                backward(i)
        else:
            for i in sites:
                forward(i)
                hint(i)
                backward(i)
        if logits_s is not None:
            N_, C_ = logits_s.shape[0], logits_s.shape[1]
            HW = logits_s.numel() // (N_ * C_)
            chk(L.kdcc_kd_loss(_ptr(logits_s), _ptr(logits_t), _ptr(self.dlogits) if self.kd_grad else None,
                               _ptr(self.kd_loss), ws, wsn, N_, C_, HW, C_ * HW, HW, 1, self.T, 0, _abi.F32,
                               1.0 / self.acc_steps, st), "kd_loss")
            mark("kd_loss")
            launches += 2
        self.launches_per_step = launches
        return self.hint_losses.sum(), self.kd_loss

    def capture(self, xs, teacher_feats, logits_s=None, logits_t=None, log=None):
        """Capture one pass (every libkdcc launch of `step`) into a CUDA graph.  Returns (graph, (hint_sum, kd)); the
        result tensors are rewritten by every `graph.replay()`.  `log` must be an EventLog(external=True)."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # one un-captured pass: function attributes / lazy module loads happen here
            self.step(xs, teacher_feats, logits_s, logits_t)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.step(xs, teacher_feats, logits_s, logits_t, log=log)
        return graph, out

    # ---- algorithmic work per step, used for the roofline figures -----------------------------
    def algorithmic(self):
        """{kernel: (bytes_or_flops_per_step, 'B'|'FLOP')} summed over the sites of one step."""
        es = 4 if self.dtype == torch.float32 else 2
        P_in, P_out, n, kk = self.H * self.W, self.Ho * self.Wo, self.N, self.k * self.k
        out = {"dw_fwd": 0, "dw_bwd": 0, "pw_fwd": 0, "pw_bwd_dx": 0, "pw_bwd_dw": 0, "hint_loss": 0, "kd_loss": 0, "relayout": 0}
        for i, (ci, co) in enumerate(self.plan):
            out["dw_fwd"] += n * (P_in + P_out) * ci * es + ci * kk * 4
            out["dw_bwd"] += n * (P_in + P_out + (P_in if self.need_dx[i] else 0)) * ci * es + ci * kk * 4
            flops = 2 * n * P_out * ci * co
            out["pw_fwd"] += flops
            out["pw_bwd_dx"] += flops
            out["pw_bwd_dw"] += flops
            out["hint_loss"] += 3 * n * P_out * co * es
            if self.relayout:   # x in (read + write), dx out where it exists; y and dy are handled inside the GEMMs
                out["relayout"] += 2 * es * n * P_in * ci * (2 if self.need_dx[i] else 1)
        if self.logits_shape:
            numel = 1
            for s_ in self.logits_shape:
                numel *= s_
            out["kd_loss"] = (3 if self.kd_grad else 2) * numel * 4
        units = {"dw_fwd": "B", "dw_bwd": "B", "hint_loss": "B", "kd_loss": "B", "relayout": "B",
                 "pw_fwd": "FLOP", "pw_bwd_dx": "FLOP", "pw_bwd_dw": "FLOP"}
        return {k_: (v, units[k_]) for k_, v in out.items()}
