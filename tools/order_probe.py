"""Measurement probe (not product): does the position of a timed region inside the process matter (clock / power state)?
Times the HotPathStep pass and the module-API pass alternately, several rounds, same process."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kdcc
from kdcc.hotpath import HotPathStep

dev = torch.device("cuda", 0)
plan = bench.plan_51m()
N, maps = 4, 128
need_dx = [False] + [True] * 8
hp = HotPathStep(plan, N, maps, maps, 9, 5, 20, dtype=torch.bfloat16, device=dev, logits_shape=(N, 19, 1024, 1024), need_dx=need_dx, order=os.environ.get("ORDER", "interleaved"))
xs, ts, ls, lt = hp.make_inputs(seed=100)
param = torch.nn.Parameter(hp.flat_params); param.grad = hp.flat_grads
opt = kdcc.optim.RAdam([param], lr=5e-3); opt.attach_lp_copy(param, hp.flat_lp); hp.refresh_lp(); hp.lp_maintained = True

def hot(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        hp.step(xs, ts, ls, lt); opt.step()
    e1.record(); torch.cuda.synchronize()
    return N * steps / (e0.elapsed_time(e1) * 1e-3)

for r in range(3):
    print("hot  %d: %.1f img/s" % (r, hot(20)), flush=True)
    a = bench.api_modules_run(plan, N, maps, (9, 5, 20), 1024, dev, 1, 20, seed=0)
    print("api  %d: %.1f img/s" % (r, a["value"]), flush=True)
print("hot  x: %.1f img/s (100 steps)" % hot(100))
