#!/usr/bin/env python
"""bench.py -- student-KD hot-path throughput (img/s at 1024x1024 crops) on N B200s of one node.

A "step" is one pass of the distillation hot path over one synthetic batch: for each of
the replaced sites of cfg/cityscapes/51M_deeplab_all.json the cheap-conv block forward (depthwise 9x9
dil 5 + pointwise GEMM), the hint MSE loss against the teacher feature with its gradient, the block
backward (pointwise dW/dX, depthwise dW/dX), the KD loss on the 19-class logits, the NCCL all-reduce of
the student gradients (N > 1) and the RAdam step.  The frozen DeepLabV3+ trunk that produces the site
inputs is outside the hot path and is not run; inputs have the trunk's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (libkdcc.so, CUDA)
  python bench.py --impl reference ...                           # the reference's torch-CPU path (oracle port)
  torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "student KD train img/s @1024^2 crop (cheap-conv hot path)"
PLAN_NAME = "cityscapes/51M_deeplab_all"
# (C_in, C_out) of the nine replaced sites (cfg/cityscapes/51M_deeplab_all.json:123-160), 128x128 maps at a 1024^2 crop
PLAN_51M = [(512, 512)] * 5 + [(1024, 2048)] + [(4096, 256)] * 3


def plan_51m():
    return list(PLAN_51M)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kdcc", choices=["kdcc", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="images (1024^2 crops) per GPU per step")
    ap.add_argument("--crop", type=int, default=1024)
    ap.add_argument("--dw", default="k9d5p20", help="depthwise geometry kKdDpP (Cityscapes cfgs: k9d5p20; CIFAR: k3d1p1)")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"],
                    help="activation layout: nchw = the reference's (tensor-core depthwise), nhwc = channels_last kernels")
    ap.add_argument("--kd-grad", action="store_true", help="also emit d KD / d logits (ClassificationTrainer path)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch.optim.RAdam + a separate weight cast instead of the fused kdcc RAdam step")
    ap.add_argument("--graph", action="store_true",
                    help="replay a CUDA graph of the pass instead of stream launches (measured: 475.8 vs 474.5 img/s, i.e. the "
                         "step is not launch-bound once the CPU runs ahead; stream launches stay the default)")
    return ap.parse_args()


def parse_geom(s):
    import re
    m = re.fullmatch(r"k(\d+)d(\d+)p(\d+)", s)
    if not m:
        raise SystemExit("bad --dw %r" % s)
    return tuple(int(v) for v in m.groups())


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); smax = float(f[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


# -------------------------------------------------------------------------------------------------------
# reference / CPU arm: the reference's own calls (torch CPU backend) restated in oracle/torch_port.py
# -------------------------------------------------------------------------------------------------------
def tensor_issue_floor(plan, batch, geom, layout, maps, measured_ms, sm_mhz, sms=148):
    """What the measured MMA cost model (tools/mma_probe.cu) says the whole-plane tensor-core depthwise
    backward (dX conv + dW) cannot go below: per 128 x 128 plane 2*d*k MMAs of 43 + 32/2 = 59 clk for the conv and
    k * 8 MMAs of 10 + 128/2 = 74 clk for the weight gradient, one plane per SM at a time, at the sampled SM clock.
    Only defined for the geometry those kernels cover; returns None otherwise."""
    k, d, p = geom
    if layout != "nchw" or maps > 128 or p % d != 0 or not sm_mhz or not measured_ms:
        return None
    planes = batch * sum(ci for ci, _ in plan)
    clk = planes * (2 * d * k * 59 + k * 8 * 74) / float(sms)
    floor_ms = clk / (sm_mhz * 1e3)
    return {"ms": round(floor_ms, 3), "frac": round(floor_ms / measured_ms, 4), "sm_mhz": sm_mhz,
            "model": "planes x (2*d*k x 59 clk conv + 8*k x 74 clk dW) / %d SMs" % sms}


def cpu_reference_image_seconds(plan, maps, geom, budget_s, crop):
    """Seconds the torch-CPU reference path needs for ONE image of the workload, from a bounded sample:
    each distinct site shape is run once (forward, hint MSE, backward) and weighted by its multiplicity,
    plus the KD loss on one image's logits.  If even that exceeds the budget the maps are halved and the
    time scaled by area (both convolutions and losses are linear in the pixel count)."""
    import torch
    from oracle import torch_port as tp
    k, d, p = geom
    torch.set_num_threads(os.cpu_count() or 1)
    shapes = {}
    for s in plan:
        shapes[s] = shapes.get(s, 0) + 1

    def run_site(ci, co, hw):
        x = torch.randn(1, ci, hw, hw)
        site = {"x": x, "w_dw": (torch.rand(ci, 1, k, k) - 0.5).requires_grad_(True),
                "w_pw": (torch.rand(co, ci, 1, 1) - 0.5).requires_grad_(True),
                "teacher": torch.randn(1, co, hw + 2 * p - d * (k - 1), hw + 2 * p - d * (k - 1)), "padding": p, "dilation": d}
        t0 = time.perf_counter()
        y = tp.block_forward(site["x"], site["w_dw"], site["w_pw"], p, d)
        loss = tp.mse_loss(y, site["teacher"], 1000)
        loss.backward()
        return time.perf_counter() - t0

    run_site(16, 16, 16)  # thread-pool / oneDNN warm-up
    scale, hw = 1.0, maps
    probe = run_site(*min(shapes, key=lambda s: s[0] * s[1]), hw)
    flops = lambda s: s[0] * (k * k + s[1])
    projected = probe / flops(min(shapes, key=lambda s: s[0] * s[1])) * sum(flops(s) * 1 for s in shapes)
    while projected > budget_s and hw > 16:
        hw //= 2
        scale *= 4.0
        projected /= 4.0
    total = 0.0
    for (ci, co), count in shapes.items():
        total += count * run_site(ci, co, hw) * scale
    lc = int(crop * (hw / maps))
    s_log, t_log = 3 * torch.randn(1, 19, lc, lc), 3 * torch.randn(1, 19, lc, lc)
    t0 = time.perf_counter()
    with torch.no_grad():
        tp.kl_div_loss(s_log, t_log, 1.0)
    total += (time.perf_counter() - t0) * scale
    sample = ("1 image: each distinct site shape %s once at %dx%d maps%s (fwd + MSE hint + bwd, weighted by multiplicity) "
              "+ KL on (1,19,%d,%d); torch CPU fp32 NCHW" % (sorted(shapes), hw, hw,
                                                           "" if scale == 1.0 else " scaled x%g by area" % scale, lc, lc))
    return total, sample


def run_reference(args, rank):
    if rank != 0:
        return
    geom = parse_geom(args.dw)
    maps = args.crop // 8
    per_step_budget = max(4.0, 150.0 / max(1, args.steps + args.warmup))
    times, sample = [], ""
    for i in range(args.warmup + args.steps):
        t, sample = cpu_reference_image_seconds(plan_51m(), maps, geom, per_step_budget, args.crop)
        if i >= args.warmup:
            times.append(t)
    sec = sum(times) / len(times)
    cores = os.cpu_count() or 1
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": PLAN_NAME + " hot path, 1024x1024 crop, dw " + args.dw, "global_batch": 1,
                       "note": "reference = its own torch calls on the host CPU (oracle/torch_port.py); one image per step"},
            "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------------
def run_kdcc(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import kdcc
    from kdcc.hotpath import EventLog, HotPathStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the kdcc path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kdcc._abi.lib()
    k, d, p = parse_geom(args.dw)
    maps = args.crop // 8
    plan = plan_51m()
    N = args.batch
    logits_shape = (N, 19, args.crop, args.crop)
    hp = HotPathStep(plan, N, maps, maps, k, d, p, dtype=torch.bfloat16, device=dev, logits_shape=logits_shape,
                     kd_temperature=1.0, hint_num_classes=1000.0, accumulation_steps=1, kd_grad=args.kd_grad, seed=rank, layout=args.layout)
    xs, ts, ls, lt = hp.make_inputs(seed=100 + rank)
    param = torch.nn.Parameter(hp.flat_params)
    param.grad = hp.flat_grads
    # cfg/cityscapes/51M_deeplab_all.json:64-69: the reference's own RAdam (utils/optim/radam.py), here as one fused
    # kdcc kernel that also emits the bf16 weights of the next step's GEMMs (--torch-optimizer: torch.optim.RAdam + cast)
    if args.torch_optimizer:
        opt = torch.optim.RAdam([param], lr=5e-3)
    else:
        opt = kdcc.optim.RAdam([param], lr=5e-3)
        if hp.flat_lp.dtype == torch.bfloat16:
            opt.attach_lp_copy(param, hp.flat_lp)
            hp.refresh_lp()
            hp.lp_maintained = True

    # --graph: the 97 libkdcc launches of one pass are captured once into a CUDA graph and replayed; the all-reduce and
    # the optimizer stay ordinary stream work; CUDA events recorded inside the capture give the per-kernel timeline of
    # the last replay.  Default: plain stream launches (the CPU enqueues a step in ~5.5 ms and runs ahead of the GPU).
    use_graph = args.graph
    glog = EventLog(external=True) if use_graph else None
    graph, gout = hp.capture(xs, ts, ls, lt, log=glog) if use_graph else (None, None)

    def one_step(log=None):
        if use_graph:
            graph.replay()
            hint, kd = gout
            if log is not None:
                log.mark("begin")
        else:
            hint, kd = hp.step(xs, ts, ls, lt, log=log)
        if world > 1:
            dist.all_reduce(hp.flat_grads, op=dist.ReduceOp.AVG)
            if log is not None:
                log.mark("grad_allreduce")
        opt.step()
        if log is not None:
            log.mark("optimizer")
        return hint, kd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    log = EventLog()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        hint, kd = one_step(log)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * N * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public call: pinned host inputs -> device, step, loss back to the host ----
    e2e = None
    if args.e2e_steps > 0:
        hx, ht, hls, hlt = hp.make_inputs(seed=200 + rank, pinned_host=True)
        h2d = sum(t_.numel() * t_.element_size() for t_ in hx + ht + [hls, hlt])
        out_host = torch.empty(2, dtype=torch.float32).pin_memory()

        # Two device input sets: while step i computes on one, the copy stream lands step i+1's host buffers in
        # the other.  Every step's H2D copy and the D2H of its losses are inside the timed region.
        sets = [(xs, ts, ls, lt), hp.make_inputs(seed=300 + rank)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_copy(i):
            dx, dt, dls, dlt = sets[i & 1]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i & 1])  # the step that last read this set has finished
                for dst, src in zip(dx + dt + [dls, dlt], hx + ht + [hls, hlt]):
                    dst.copy_(src, non_blocking=True)
                copied[i & 1].record(copy_stream)

        def e2e_run(steps):
            main = torch.cuda.current_stream()
            for c in consumed:
                c.record(main)
            issue_copy(0)
            for i in range(steps):
                if i + 1 < steps:
                    issue_copy(i + 1)
                main.wait_event(copied[i & 1])
                dx, dt, dls, dlt = sets[i & 1]
                h_, k_ = hp.step(dx, dt, dls, dlt)
                if world > 1:
                    dist.all_reduce(hp.flat_grads, op=dist.ReduceOp.AVG)
                opt.step()
                consumed[i & 1].record(main)
                out_host.copy_(torch.stack([h_, k_]), non_blocking=True)

        e2e_run(2)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(args.e2e_steps)
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * N * args.e2e_steps / (ems * 1e-3), "unit": "img/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": 8, "steps": args.e2e_steps,
               "pipeline": "H2D of step i+1 on a copy stream overlaps the kernels of step i (two device input sets)", "last_losses": [float(out_host[0]), float(out_host[1])]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline from the CUDA events recorded inside the timed region --------------------------
    pk = peaks()
    durs = log.durations_ms()            # stream work of every timed step (all of it when --no-graph)
    samples = {name: args.steps for name in durs}
    if use_graph:
        for name, lst in glog.durations_ms().items():   # events inside the graph: the last timed step
            durs[name] = lst
            samples[name] = 1
    alg = hp.algorithmic()
    kernels, dominant, dom_ms = {}, None, -1.0
    for name, lst in durs.items():
        per_step_ms = sum(lst) / samples[name]
        entry = {"ms_per_step": round(per_step_ms, 4), "share": round(per_step_ms / ms_per_step, 4), "launches_per_step": len(lst) // samples[name]}
        if name in alg:
            work, unit = alg[name]
            if unit == "B":
                entry.update(bound="hbm", achieved=round(work / (per_step_ms * 1e-3) / 1e9, 1), unit="GB/s",
                             frac=round(work / (per_step_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4))
            else:
                entry.update(bound="tensor", achieved=round(work / (per_step_ms * 1e-3) / 1e12, 1), unit="TFLOP/s",
                             frac=round(work / (per_step_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], 4))
            if per_step_ms > dom_ms:
                dominant, dom_ms = name, per_step_ms
        kernels[name] = entry
    dk = kernels[dominant]
    # measured DRAM traffic of the dominant family over one step: ncu dram__bytes_read.sum + dram__bytes_write.sum of
    # every launch of one step of this same default workload (tools/gpu_profile.sh -> profiles/r01_traffic.json)
    traffic = None
    try:
        if (N, args.dw, args.layout, args.crop) == (4, "k9d5p20", "nchw", 1024):
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                traffic = json.load(f)["families"][dominant]["dram_bytes_per_step"]
    except Exception:
        traffic = None
    note = ""
    if dominant.startswith("dw") and k >= 7:
        note = ("achieved = algorithmic bytes of all %d launches of one step / their CUDA-event time; traffic = ncu DRAM bytes of the "
                "same launches.  k=9 depthwise is 81 MAC per output element (20 MAC per HBM byte): the kernels run on tcgen05 and are "
                "bound by the issue cost of their small-N MMAs (90 x 59 clk conv, 72 x 74 clk dW per plane, DESIGN.md 4.0/4.1), "
                "not by HBM; frac is still reported against the HBM copy peak" % dk["launches_per_step"])
    roofline = {"kernel": dominant, "bound": dk["bound"], "achieved": dk["achieved"],
                "peak": pk["hbm_gbs"] if dk["bound"] == "hbm" else pk["bf16_tflops_sustained"], "unit": dk["unit"],
                "frac": dk["frac"], "traffic": traffic, "algorithmic_bytes": alg[dominant][0] if alg[dominant][1] == "B" else None,
                "peak_source": pk["source"] + (" copy bandwidth" if dk["bound"] == "hbm" else " sustained cuBLAS bf16"),
                "note": note}
    try:  # the resource that actually binds the dominant family (an annotation: it must never break the line)
        if dominant == "dw_bwd" and k >= 7:
            roofline["tensor_issue_floor"] = tensor_issue_floor(plan, N, (k, d, p), args.layout, maps, dk["ms_per_step"],
                                                                (clocks or {}).get("sm_mhz"))
    except Exception:
        pass

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        sec, sample = cpu_reference_image_seconds(plan, maps, (k, d, p), 25.0, args.crop)
        cpu_baseline = {"value": 1.0 / sec, "unit": "img/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "%s hot path: %d cheap-conv sites (dw %s + pw GEMM) fwd+bwd, hint MSE x%d, KD loss on (N,19,%d,%d), "
                                   "grad all-reduce, RAdam; frozen trunk not run" % (PLAN_NAME, len(plan), args.dw, len(plan), args.crop, args.crop),
                       "global_batch": world * N, "per_gpu_batch": N, "crop": args.crop, "feature_maps": "%dx%d" % (maps, maps),
                       "parallelism": "dp%d" % world, "trainable_params": hp.num_trainable, "layout": args.layout,
                       "launch": ("CUDA graph of the pass replayed per step; per-kernel times = CUDA events inside the graph, last timed step"
                                  if use_graph else "stream launches; per-kernel times = CUDA events between launches, all timed steps"),
                       "cache": "inputs larger than L2 (per-step working set of several GB >> 126 MB), no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": (hp.launches_per_step + (0 if args.torch_optimizer else 1)) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
            "losses": {"hint": float(hint), "kd": float(kd)}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_kdcc(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
