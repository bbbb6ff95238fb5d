"""Times the depthwise weight gradient of the 4096-channel site (and a 512-channel one) through the C ABI; run under
KDCC_LIB=<variant> KDCC_TC_DEBUG=<bits> for stage-skipping experiments (tools/build_variant.sh ... -DKDCC_DEBUG)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
abi = importlib.import_module("knowledge-distillation-by-replacing-cheap-conv_b200._abi")
lib = abi.lib()
dev = "cuda"
for (n, c) in ((4, 4096), (4, 512)):
    x = torch.randn(n, c, 128, 128, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, c, 128, 128, device=dev).to(torch.bfloat16)
    dwg = torch.empty(c, 81, device=dev)
    ws_bytes = lib.kdcc_dw_bwd_workspace_bytes(n, 128, 128, c, 9, 5, 20, abi.NCHW, abi.BF16)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.kdcc_dw_bwd(x.data_ptr(), None, dy.data_ptr(), None, dwg.data_ptr(), None, ws.data_ptr(), ws_bytes, n, 128, 128, c, 9, 5, 20,
                             abi.NCHW, abi.BF16, st)
        assert rc == 0, rc
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(30):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("dbg=%s C=%d: %.4f ms" % (os.environ.get("KDCC_TC_DEBUG", "0"), c, e0.elapsed_time(e1) / 30), flush=True)
