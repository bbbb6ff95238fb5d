"""Multi-teacher KD (trainer/ensemble_trainer.py:76-83).  Golden vectors: the reference's own
KLDivergenceLoss module combined exactly as the ensemble trainer combines it (oracle/make_golden.py:ensemble_golden)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ensemble.npz")
TAGS = ["seg_T1", "cifar_T5", "seg_T2_w"]


def relerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


def _weights(K):
    return [1.0 / K] * K  # WEIGHT = 1: (sum_k KL_k + KL_teacher) / (K_ens + 1)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden(golden, tag):
    g = golden
    teachers = list(g[tag + "/teachers"])
    loss, ds = orc.kd_loss_multi(g[tag + "/s"], teachers, _weights(len(teachers)), float(g[tag + "/T"]))
    assert abs(loss - float(g[tag + "/loss"])) <= 1e-5 * abs(float(g[tag + "/loss"]))
    assert relerr(ds, g[tag + "/ds"]) < 1e-5


@pytest.mark.parametrize("tag", TAGS)
def test_torch_port_matches_reference_golden(golden, tag):
    import torch
    from oracle import torch_port as tp
    g = golden
    s = torch.from_numpy(g[tag + "/s"]).requires_grad_(True)
    ts = [torch.from_numpy(t) for t in g[tag + "/teachers"]]
    loss = tp.ensemble_kd_loss(s, ts[:-1], ts[-1], float(g[tag + "/T"]), 1.0)
    loss.backward()
    assert abs(float(loss) - float(g[tag + "/loss"])) <= 1e-6 * abs(float(g[tag + "/loss"]))
    assert relerr(s.grad.numpy(), g[tag + "/ds"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("tag", TAGS)
def test_kernel_matches_reference_golden(golden, tag, dtype):
    import torch
    import kdcc
    g = golden
    tol = 1e-5 if dtype == "float32" else 2e-2   # BASELINE.json tolerances
    dt = getattr(torch, dtype)
    s = torch.from_numpy(g[tag + "/s"]).cuda().to(dt).requires_grad_(True)
    ts = [torch.from_numpy(t).cuda().to(dt) for t in g[tag + "/teachers"]]
    crit = kdcc.MultiTeacherKLDivergenceLoss(temperature=float(g[tag + "/T"]), weight=1)
    loss = crit(s, ts[:-1], ts[-1])
    assert loss.dim() == 0 and loss.grad_fn is not None
    (2.0 * loss).backward()
    if dtype == "float32":
        ref_loss, ref_ds = float(g[tag + "/loss"]), g[tag + "/ds"]
    else:  # the oracle sees the same bf16-rounded inputs
        ref_loss, ref_ds = orc.kd_loss_multi(s.detach().float().cpu().numpy(), [t.float().cpu().numpy() for t in ts],
                                             _weights(len(ts)), float(g[tag + "/T"]))
    assert abs(float(loss) - ref_loss) <= tol * abs(ref_loss)
    assert relerr(s.grad.float().cpu().numpy(), 2.0 * ref_ds) < tol


@pytest.mark.gpu
def test_multi_equals_sum_of_single_and_full_size_properties():
    """(N,19,512,512): the fused pass equals the weighted sum of single-teacher kdcc_kd_loss calls; identical teachers
    reduce to the single-teacher loss; s == every teacher gives zero loss and zero gradient."""
    import torch
    import kdcc
    torch.manual_seed(9)
    s = (3 * torch.randn(2, 19, 512, 512, device="cuda")).requires_grad_(True)
    ts = [3 * torch.randn(2, 19, 512, 512, device="cuda") for _ in range(3)]
    w = [0.5, 0.3, 0.2]
    multi = kdcc.functional.kd_loss_multi(s, ts, w, 2.0)
    multi.backward()
    gm = s.grad.clone(); s.grad = None
    single = sum(wk * kdcc.functional.kd_loss(s, t, 2.0) for wk, t in zip(w, ts))
    single.backward()
    assert abs(float(multi) - float(single)) <= 1e-5 * abs(float(single))
    assert float((gm - s.grad).abs().max()) <= 1e-5 * float(s.grad.abs().max())
    same = kdcc.functional.kd_loss_multi(s.detach(), [ts[0]] * 4, [0.25] * 4, 1.0)
    one = kdcc.functional.kd_loss(s.detach(), ts[0], 1.0)
    assert abs(float(same) - float(one)) <= 1e-5 * abs(float(one))
    z = s.detach().clone().requires_grad_(True)
    l0 = kdcc.functional.kd_loss_multi(z, [z.detach()] * 2, [0.5, 0.5], 1.0)
    l0.backward()
    assert abs(float(l0)) < 1e-6 and float(z.grad.abs().max()) < 1e-9


def test_multi_errors_are_loud():
    import torch
    import kdcc
    with pytest.raises(kdcc.KdccError):
        kdcc.functional.kd_loss_multi(torch.zeros(1, 19, 2, 2), [torch.zeros(1, 19, 2, 2)], [1.0])
