"""Times the TTA stitch at the shipped Cityscapes test setting (cfg/cityscapes/51M_deeplab_all.json:255-259: scales [1.0],
crop 1024 on 2048 x 1024 frames -> 3 windows + 3 mirrored, 19 classes) against the reference's host formulation
(device->host copy of the window outputs + the numpy oracle of utils/tta_process.py).  Measurement tooling."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import kdcc
from kdcc import tta
from oracle import oracle as orc
W, H, crop, C = 2048, 1024, 1024, 19
boxes = tta.window_coordinates(W, H, crop)
mapping = [[W, H, boxes]]
res = torch.randn(2 * len(boxes), C, crop, crop, device="cuda")
for _ in range(3): out = tta.reverse_mapping(mapping, res, (W, H))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = tta.reverse_mapping(mapping, res, (W, H))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
bytes_ = res.numel() * 4 + 3 * out.numel() * 4   # windows read once, output written, re-read and re-written by the mirrored pass
print("gpu reverse_mapping %.3f ms  %.0f GB/s (%d windows)" % (ms, bytes_ / ms / 1e6, 2 * len(boxes)))
t0 = time.time(); host = res.cpu().numpy(); t1 = time.time()
with np.errstate(divide="ignore", invalid="ignore"):
    want = orc.tta_reverse_mapping(mapping, host, (W, H))
t2 = time.time()
print("host: D2H %.0f ms + numpy stitch %.0f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
fin = np.isfinite(want)
print("max rel err", np.abs(out.cpu().numpy()[0][fin[0]] - want[0][fin[0]]).max() / np.abs(want[fin]).max())
