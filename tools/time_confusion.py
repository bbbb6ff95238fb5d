"""Times kdcc_confusion_update on BASELINE-size logits (4,19,1024,1024) fp32 + int64 labels (measurement tooling)."""
import sys, torch
sys.path.insert(0, ".")
import kdcc
N, C, H, W = 4, 19, 1024, 1024
logits = torch.randn(N, C, H, W, device="cuda"); labels = torch.randint(0, C, (N, H, W), device="cuda")
cm = kdcc.ConfusionMatrix(C, 255)
for _ in range(3): cm.update(logits, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): cm.update(logits, labels)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = N * H * W * (C * 4 + 8)
print("confusion_update: %.3f ms, %.0f GB/s algorithmic (%d MB)" % (ms, nbytes / ms / 1e6, nbytes >> 20))
import time
t0 = time.perf_counter()
pred = logits.argmax(1).cpu().numpy().ravel(); tg = labels.cpu().numpy().ravel()
import numpy as np
np.bincount(C * tg + pred, minlength=C * C)
print("reference formulation (D2H + numpy bincount): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
