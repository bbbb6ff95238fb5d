"""Can CUDA events recorded inside a captured graph time its kernels?  (measurement tooling)"""
import torch
x = torch.randn(64 << 20, device="cuda")
y = torch.empty_like(x)
evs = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    y.copy_(x); y.mul_(2.0)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    evs[0].record()
    y.copy_(x)
    evs[1].record()
    for _ in range(4):
        y.mul_(1.0001)
    evs[2].record()
for rep in range(3):
    g.replay()
    torch.cuda.synchronize()
    print("replay", rep, "copy %.3f ms, 4 muls %.3f ms" % (evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2])))
