"""Times the NCHW depthwise kernels one by one through the C ABI (CUDA events, L2-exceeding working set).

    python tools/time_dw.py [N C H W k d p] [reps]

Prints ms and clocks-per-plane-per-SM for fwd, dX and dW.  Env KDCC_TC_DEBUG etc. are passed through to the
kernels, so `KDCC_TC_DEBUG=1 python tools/time_dw.py` gives the skip-this-stage timings used in DESIGN.md.
"""
import sys
import torch
sys.path.insert(0, ".")
import kdcc
from kdcc import _abi

args = [int(v) for v in sys.argv[1:8]] if len(sys.argv) >= 8 else [4, 4096, 128, 128, 9, 5, 20]
N, C, H, W, k, d, p = args
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 10
L = _abi.lib()
dev = "cuda"
x = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16)
dy = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16)
w = (torch.rand(C, k * k, device=dev) * 2 - 1) / k
y = torch.empty_like(x)
dx = torch.empty_like(x)
dw = torch.empty(C, k * k, device=dev)
ws = torch.empty(max(16, L.kdcc_dw_bwd_workspace_bytes(N, H, W, C, k, d, p, _abi.NCHW, _abi.BF16)), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
ptr = lambda t: t.data_ptr() if t is not None else None


def fwd():
    _abi.check(L.kdcc_dw_fwd(ptr(x), ptr(w), None, ptr(y), N, H, W, C, k, d, p, _abi.NCHW, _abi.BF16, st), "fwd")


def bwd_dx():
    _abi.check(L.kdcc_dw_bwd(ptr(x), ptr(w), ptr(dy), ptr(dx), None, None, ptr(ws), ws.numel(), N, H, W, C, k, d, p,
                             _abi.NCHW, _abi.BF16, st), "dx")


def bwd_dw():
    _abi.check(L.kdcc_dw_bwd(ptr(x), ptr(w), ptr(dy), None, ptr(dw), None, ptr(ws), ws.numel(), N, H, W, C, k, d, p,
                             _abi.NCHW, _abi.BF16, st), "dw")


import threading, time
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())


def clock_under_load(fn, seconds=0.4):
    """median SM clock (MHz) sampled while fn runs back to back"""
    samples, stop = [], [False]
    def sampler():
        while not stop[0]:
            samples.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.01)
    th = threading.Thread(target=sampler)
    t0 = time.time()
    th.start()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    stop[0] = True
    th.join()
    samples = sorted(samples[len(samples) // 3:]) or [0]
    return samples[len(samples) // 2]


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


planes_per_sm = N * C / 148.0
el = 2
for name, fn, nbytes in (("fwd", fwd, 2 * N * C * H * W * el), ("dx", bwd_dx, 2 * N * C * H * W * el),
                         ("dw", bwd_dw, 2 * N * C * H * W * el)):
    ms = timeit(fn)
    mhz = clock_under_load(fn)
    print("%-3s %.3f ms   %.0f GB/s   %.0f clk/plane @ %d MHz (sampled under load)" % (name, ms, nbytes / ms / 1e6, ms * 1e-3 * mhz * 1e6 / planes_per_sm, mhz))
