"""RAdam of the layerwise loop.  The numpy oracle is pinned on the reference's own optimizer class stepped on seeded
parameters / gradients (tests/golden/radam.npz, frozen by oracle/make_golden.py from utils/optim/radam.py); the fused
CUDA step must follow the same trajectory (p, exp_avg, exp_avg_sq after every step) to 1e-6 relative, through the
degenerated-to-SGD steps, the adaptive steps, weight decay and the moments-only branch."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "radam.npz")
TAGS = ["plain", "wd", "nosgd"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_follows_reference_radam(golden, tag):
    g = golden
    lr, b1, b2, eps, wd, sgd = g[tag + "/hyper"]
    p = g[tag + "/p0"]
    m, v = np.zeros_like(p), np.zeros_like(p)
    for t in range(g[tag + "/grads"].shape[0]):
        p, m, v = orc.radam_step(p, g[tag + "/grads"][t], m, v, t + 1, lr, (b1, b2), eps, wd, bool(sgd))
        assert _rel(p, g[tag + "/p"][t]) < 1e-6 and _rel(m, g[tag + "/m"][t]) < 1e-6 and _rel(v, g[tag + "/v"][t]) < 1e-6
    # the schedule the branches depend on (radam.py:65-84): SGD-like while N_sma < 5, adaptive afterwards
    assert orc.radam_scalars(1, 0.9, 0.999)[0] < 5 <= orc.radam_scalars(6, 0.9, 0.999)[0]
    assert orc.radam_scalars(1, 0.9, 0.999, False)[1] == -1


def test_host_mirror_schedule_and_state_layout():
    import torch
    from kdcc.optim import RAdam
    prm = torch.nn.Parameter(torch.zeros(4))
    opt = RAdam([prm], lr=5e-3)
    for step in (1, 3, 5, 6, 50, 1000):
        mine, want = opt.step_scalars(step, 0.9, 0.999), orc.radam_scalars(step, 0.9, 0.999)
        assert mine == want
    sd = opt.state_dict()
    assert set(sd["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "buffer"}   # the reference's group keys
    with pytest.raises(ValueError):
        RAdam([prm], lr=-1.0)
    prm.grad = torch.zeros(4)
    with pytest.raises(Exception):   # CPU parameters: no fallback
        opt.step()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_fused_step_follows_reference_trajectory(golden, tag):
    import torch
    from kdcc.optim import RAdam
    g = golden
    lr, b1, b2, eps, wd, sgd = g[tag + "/hyper"]
    prm = torch.nn.Parameter(torch.from_numpy(g[tag + "/p0"].copy()).cuda())
    lp = torch.zeros(prm.numel(), dtype=torch.bfloat16, device="cuda")
    opt = RAdam([prm], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd, degenerated_to_sgd=bool(sgd))
    opt.attach_lp_copy(prm, lp)
    for t in range(g[tag + "/grads"].shape[0]):
        prm.grad = torch.from_numpy(g[tag + "/grads"][t].copy()).cuda()
        opt.step()
        st = opt.state[prm]
        assert st["step"] == t + 1
        assert _rel(prm.detach().cpu().numpy(), g[tag + "/p"][t]) < 1e-6
        assert _rel(st["exp_avg"].cpu().numpy(), g[tag + "/m"][t]) < 1e-6
        assert _rel(st["exp_avg_sq"].cpu().numpy(), g[tag + "/v"][t]) < 1e-6
        assert torch.equal(lp, prm.detach().to(torch.bfloat16))   # the bf16 copy is the rounded new parameter


@pytest.mark.gpu
def test_fused_step_large_ragged_and_state_dict_roundtrip():
    import torch
    from kdcc.optim import RAdam
    rs = np.random.RandomState(2)
    n = 4 * 70001 + 3   # vector body + ragged tail
    p0 = rs.standard_normal(n).astype(np.float32)
    prm = torch.nn.Parameter(torch.from_numpy(p0.copy()).cuda())
    opt = RAdam([prm], lr=1e-2, weight_decay=1e-3)
    p, m, v = p0, np.zeros_like(p0), np.zeros_like(p0)
    for t in range(7):
        gr = rs.standard_normal(n).astype(np.float32)
        prm.grad = torch.from_numpy(gr).cuda()
        opt.step()
        p, m, v = orc.radam_step(p, gr, m, v, t + 1, 1e-2, (0.9, 0.999), 1e-8, 1e-3)
    assert _rel(prm.detach().cpu().numpy(), p) < 1e-6
    import copy
    sd = copy.deepcopy(opt.state_dict())   # load_state_dict keeps same-dtype state tensors by reference
    prm2 = torch.nn.Parameter(prm.detach().clone())
    opt2 = RAdam([prm2], lr=1e-2, weight_decay=1e-3)
    opt2.load_state_dict(sd)
    gr = rs.standard_normal(n).astype(np.float32)
    for o, q_ in ((opt, prm), (opt2, prm2)):
        q_.grad = torch.from_numpy(gr).cuda()
        o.step()
    assert torch.equal(prm.detach(), prm2.detach())
