"""Dumps the hand-off trace CTA 0 of dw_tc_wgrad3_kernel records in a -DKDCC_DEBUG build (tools/build_variant.sh w3dbg
dw_tc_wgrad3.cu -DKDCC_DEBUG; run with KDCC_LIB=.../libkdcc_w3dbg.so) to gpurun_out/w3_trace.txt: warp, event, clock."""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("knowledge-distillation-by-replacing-cheap-conv_b200._abi")
lib = abi.lib()
n, c = 4, 4096
x = torch.randn(n, c, 128, 128, device="cuda").to(torch.bfloat16)
dy = torch.randn(n, c, 128, 128, device="cuda").to(torch.bfloat16)
dwg = torch.empty(c, 81, device="cuda")
ws_bytes = lib.kdcc_dw_bwd_workspace_bytes(n, 128, 128, c, 9, 5, 20, abi.NCHW, abi.BF16)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    rc = lib.kdcc_dw_bwd(x.data_ptr(), None, dy.data_ptr(), None, dwg.data_ptr(), None, ws.data_ptr(), ws_bytes, n, 128, 128, c, 9, 5, 20,
                         abi.NCHW, abi.BF16, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
N = 19 * 512
buf = (ctypes.c_longlong * N)()
fn = lib.kdcc_debug_w3_trace
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf, N) == 0
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/w3_trace.txt", "w") as f:
    for w in range(19):
        for i in range(512):
            v = buf[w * 512 + i]
            if v:
                f.write("%d %d %d\n" % (w, v & 255, v >> 8))
print("trace written")
