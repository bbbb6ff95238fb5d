"""Size-independent identities the oracle itself must satisfy (it is the checker of every GPU test): the depthwise /
pointwise backward functions are the adjoints of the forward ones, the losses are consistent with their own gradients
(finite differences), and the KD loss follows the closed forms of losses/KLDiv.py on degenerate inputs.  Random
geometries through hypothesis.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc


def _dot(a, b):
    return float((np.asarray(a, np.float64) * np.asarray(b, np.float64)).sum())


geoms = st.tuples(st.integers(1, 2), st.integers(1, 4), st.integers(3, 11), st.integers(3, 12),
                  st.sampled_from([1, 3, 5]), st.integers(1, 3), st.integers(0, 4), st.integers(0, 2 ** 16))


@settings(max_examples=25, deadline=None)
@given(geoms)
def test_depthwise_backward_is_the_adjoint_of_forward(g):
    N, C, H, W, k, d, p, seed = g
    if H + 2 * p - d * (k - 1) < 1 or W + 2 * p - d * (k - 1) < 1:
        return
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((N, C, H, W)).astype(np.float32)
    w = rs.standard_normal((C, 1, k, k)).astype(np.float32)
    y = orc.dw_fwd(x, w, k, d, p)
    dy = rs.standard_normal(y.shape).astype(np.float32)
    dx, dw, _ = orc.dw_bwd(x, w, dy, k, d, p)
    lhs = _dot(y, dy)
    assert abs(lhs - _dot(x, dx)) <= 1e-4 * max(1.0, abs(lhs))   # <conv(x), dy> == <x, conv^T(dy)>
    assert abs(lhs - _dot(w, dw)) <= 1e-4 * max(1.0, abs(lhs))   # ... == <w, dW>  (the conv is bilinear)


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(1, 2), st.integers(1, 9), st.integers(1, 7), st.integers(1, 5), st.integers(1, 6), st.integers(0, 2 ** 16)))
def test_pointwise_backward_is_the_adjoint_of_forward(g):
    N, K, Co, H, W, seed = g
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((N, K, H, W)).astype(np.float32)
    w = rs.standard_normal((Co, K, 1, 1)).astype(np.float32)
    y = orc.pw_fwd(x, w)
    dy = rs.standard_normal(y.shape).astype(np.float32)
    dx, dw, _ = orc.pw_bwd(x, w, dy)
    lhs = _dot(y, dy)
    assert abs(lhs - _dot(x, dx)) <= 1e-4 * max(1.0, abs(lhs))
    assert abs(lhs - _dot(w, dw)) <= 1e-4 * max(1.0, abs(lhs))


def test_loss_gradients_match_finite_differences():
    rs = np.random.RandomState(3)
    s = rs.standard_normal((2, 5, 3, 4)).astype(np.float32)
    t = rs.standard_normal((2, 5, 3, 4)).astype(np.float32)
    wgt = rs.uniform(0.1, 1.0, 5).astype(np.float32)
    cases = [("kd T=2", lambda a: orc.kd_loss(a, t, 2.0)), ("hint", lambda a: orc.hint_loss(a, t, wgt, 1.0)),
             ("mse x19", lambda a: orc.hint_loss(a, t, None, 19.0))]
    for name, f in cases:
        loss, grad = f(s)
        for idx in [(0, 0, 0, 0), (1, 4, 2, 3), (0, 2, 1, 1)]:
            e = np.zeros_like(s)
            e[idx] = 1e-2
            num = (f(s + e)[0] - f(s - e)[0]) / 2e-2
            assert abs(num - grad[idx]) <= 2e-3 * max(1.0, abs(num)), (name, idx, num, grad[idx])


def test_kd_closed_forms():
    # equal logits -> 0; a uniform teacher against a one-hot-ish student -> T^2 * (mean log-ratio), losses/KLDiv.py:19-23
    s = np.random.RandomState(1).standard_normal((3, 7, 2, 2)).astype(np.float32)
    assert abs(orc.kd_loss(s, s, 3.0)[0]) < 1e-7
    assert np.abs(orc.kd_loss(s, s, 3.0)[1]).max() < 1e-7
    C, T = 7, 2.0
    t = np.zeros((1, C, 1, 1), np.float32)
    s1 = np.zeros((1, C, 1, 1), np.float32)
    s1[0, 0] = 4.0
    ps = np.exp(s1.reshape(C) / T) / np.exp(s1.reshape(C) / T).sum()
    want = T * T * float((np.full(C, 1 / C) * (np.log(1 / C) - np.log(ps))).sum())
    assert abs(orc.kd_loss(s1, t, T)[0] - want) < 1e-6 * max(1.0, want)
