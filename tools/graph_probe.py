"""Does CUDA-graph replay of the hot-path step beat stream launches?  (measurement tooling)"""
import sys, time
import torch
sys.path.insert(0, ".")
import kdcc
from kdcc.hotpath import HotPathStep
from bench import plan_51m

dev = torch.device("cuda", 0)
N = 4
hp = HotPathStep(plan_51m(), N, 128, 128, 9, 5, 20, dtype=torch.bfloat16, device=dev, logits_shape=(N, 19, 1024, 1024),
                 kd_temperature=1.0, hint_num_classes=1000.0, accumulation_steps=1, kd_grad=False, seed=0, layout="nchw")
xs, ts, ls, lt = hp.make_inputs(seed=100)
param = torch.nn.Parameter(hp.flat_params)
param.grad = hp.flat_grads
opt = torch.optim.RAdam([param], lr=5e-3, capturable=True)

def step():
    h, k = hp.step(xs, ts, ls, lt)
    opt.step()
    return h, k

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("stream launches: %.3f ms/step" % timed(step))
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("cpu enqueue time: %.3f ms/step" % ((t1 - t0) * 1e3 / 20))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = step()
print("graph replay:    %.3f ms/step" % timed(g.replay))
print("losses", float(out[0]), float(out[1]))
