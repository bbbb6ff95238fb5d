"""Times the NHWC bf16 depthwise kernels through the C ABI (measurement tooling): python tools/time_dw_nhwc.py [N C H W k d p]"""
import sys, torch
sys.path.insert(0, ".")
import kdcc
from kdcc import _abi
args = [int(v) for v in sys.argv[1:8]] if len(sys.argv) >= 8 else [4, 4096, 128, 128, 3, 1, 1]
N, C, H, W, k, d, p = args
L = _abi.lib(); dev = "cuda"
x = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16); dy = torch.randn_like(x.float()).to(torch.bfloat16)
w = (torch.rand(C, k * k, device=dev) * 2 - 1) / k
y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.empty(C, k * k, device=dev)
ws = torch.empty(max(16, L.kdcc_dw_bwd_workspace_bytes(N, H, W, C, k, d, p, _abi.NHWC, _abi.BF16)), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr() if t is not None else None
fns = {"fwd": lambda: _abi.check(L.kdcc_dw_fwd(P(x), P(w), None, P(y), N, H, W, C, k, d, p, _abi.NHWC, _abi.BF16, st), "fwd"),
       "dx": lambda: _abi.check(L.kdcc_dw_bwd(P(x), P(w), P(dy), P(dx), None, None, P(ws), ws.numel(), N, H, W, C, k, d, p, _abi.NHWC, _abi.BF16, st), "dx"),
       "dw": lambda: _abi.check(L.kdcc_dw_bwd(P(x), P(w), P(dy), None, P(dw), None, P(ws), ws.numel(), N, H, W, C, k, d, p, _abi.NHWC, _abi.BF16, st), "dw")}
for name, fn in fns.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-3s %.3f ms  %.0f GB/s" % (name, ms, 2 * N * C * H * W * 2 / ms / 1e6))
