"""numpy/ctypes front end of the C oracle (oracle/kdcc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

All arrays are the reference's layout: NCHW, contiguous, float32.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkdcc_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile kdcc_oracle.c with gcc via oracle/build.sh (a second or two)."""
    src = os.path.join(_HERE, "kdcc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        i, l, d, p = ctypes.c_int, ctypes.c_long, ctypes.c_double, _f32p
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_dw_fwd.argtypes = [p, p, p, p, i, i, i, i, i, i, i]
        L.orc_dw_bwd.argtypes = [p, p, p, p, p, p, i, i, i, i, i, i, i]
        L.orc_pw_fwd.argtypes = [p, p, p, p, i, i, i, l]
        L.orc_pw_bwd.argtypes = [p, p, p, p, p, p, i, i, i, l]
        L.orc_kd_loss.argtypes = [p, p, p, i, i, l, d, i]
        L.orc_kd_loss.restype = d
        L.orc_hint_loss.argtypes = [p, p, p, i, p, i, i, l, d]
        L.orc_hint_loss.restype = d
        L.orc_nchw_to_nhwc.argtypes = [p, p, i, i, l]
        L.orc_nhwc_to_nchw.argtypes = [p, p, i, i, l]
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def dw_fwd(x, w, k, d, p, bias=None):
    """x (N,C,H,W), w (C,1,k,k) -> y (N,C,Ho,Wo)."""
    x, w = _c(x), _c(w)
    N, C, H, W = x.shape
    Ho, Wo = H + 2 * p - d * (k - 1), W + 2 * p - d * (k - 1)
    y = np.empty((N, C, Ho, Wo), np.float32)
    b = _c(bias) if bias is not None else None
    lib().orc_dw_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, C, H, W, k, d, p)
    return y


def dw_bwd(x, w, dy, k, d, p, need_dx=True, need_dbias=False):
    """Returns (dx or None, dw (C,1,k,k), dbias or None)."""
    x, w, dy = _c(x), _c(w), _c(dy)
    N, C, H, W = x.shape
    dx = np.empty_like(x) if need_dx else None
    dw = np.empty((C, 1, k, k), np.float32)
    db = np.empty((C,), np.float32) if need_dbias else None
    lib().orc_dw_bwd(_ptr(x), _ptr(w), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), N, C, H, W, k, d, p)
    return dx, dw, db


def pw_fwd(x, w, bias=None):
    """x (N,Ci,H,W), w (Co,Ci,1,1) -> y (N,Co,H,W)."""
    x, w = _c(x), _c(w)
    N, Ci, H, W = x.shape
    Co = w.shape[0]
    y = np.empty((N, Co, H, W), np.float32)
    b = _c(bias) if bias is not None else None
    lib().orc_pw_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, Ci, Co, H * W)
    return y


def pw_bwd(x, w, dy, need_dx=True, need_dbias=False):
    x, w, dy = _c(x), _c(w), _c(dy)
    N, Ci, H, W = x.shape
    Co = w.shape[0]
    dx = np.empty_like(x) if need_dx else None
    dw = np.empty((Co, Ci, 1, 1), np.float32)
    db = np.empty((Co,), np.float32) if need_dbias else None
    lib().orc_pw_bwd(_ptr(x), _ptr(w), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), N, Ci, Co, H * W)
    return dx, dw, db


def kd_loss(s, t, T=1.0, target_is_prob=False, need_grad=True):
    """s, t (N,C,...) -> (loss, ds)."""
    s, t = _c(s), _c(t)
    N, C = s.shape[:2]
    HW = int(np.prod(s.shape[2:])) if s.ndim > 2 else 1
    ds = np.empty_like(s) if need_grad else None
    loss = lib().orc_kd_loss(_ptr(s), _ptr(t), _ptr(ds), N, C, HW, float(T), int(bool(target_is_prob)))
    return loss, ds


def hint_loss(s, t, w=None, scale=1.0, need_grad=True):
    """WeightedHintMSELoss (w given, scale=1) or MSELoss (w None, scale=num_classes)."""
    s, t = _c(s), _c(t)
    N, C = s.shape[:2]
    HW = int(np.prod(s.shape[2:])) if s.ndim > 2 else 1
    per_sample = 0
    if w is not None:
        w = _c(w)
        per_sample = int(w.ndim == 2)
        assert w.shape[-1] == C
    ds = np.empty_like(s) if need_grad else None
    loss = lib().orc_hint_loss(_ptr(s), _ptr(t), _ptr(w), per_sample, _ptr(ds), N, C, HW, float(scale))
    return loss, ds


def block_fwd_bwd(x, w_dw, w_pw, k, d, p, dy=None, need_dx=True):
    """DepthwiseSeparableBlock forward (+ backward when dy given), reference layout."""
    mid = dw_fwd(x, w_dw, k, d, p)
    y = pw_fwd(mid, w_pw)
    if dy is None:
        return y
    dmid, dw_pw, _ = pw_bwd(mid, w_pw, dy)
    dx, dw_dw, _ = dw_bwd(x, w_dw, dmid, k, d, p, need_dx=need_dx)
    return y, dx, dw_dw, dw_pw


def confusion(outputs, labels, num_classes=None, ignore_index=255):
    """Confusion matrix of argmax(outputs, axis 1) against labels, numpy restatement of
    utils/util.py:108-128 (CityscapesMetricTracker.update + confusion_for_batch): labels == ignore_index are
    rewritten to num_classes and masked out, hist[target][pred] = bincount(C * target + pred).  int64 (C, C)."""
    outputs, labels = np.asarray(outputs), np.asarray(labels).copy()
    C = outputs.shape[1] if num_classes is None else num_classes
    labels[labels == ignore_index] = C
    pred = outputs.argmax(axis=1).reshape(-1)
    target = labels.reshape(-1)
    mask = (target >= 0) & (target < C)
    return np.bincount(C * target[mask].astype(np.int64) + pred[mask], minlength=C * C).reshape(C, C).astype(np.int64)


def mean_iou(conf):
    """utils/util.py:113-118 (get_iou): nanmean of tp / (rows + cols - tp); 1.0 for an all-zero matrix."""
    conf = np.asarray(conf, np.float64)
    if not np.any(conf):
        return 1.0
    tp = np.diag(conf)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.nanmean(tp / (conf.sum(0) + conf.sum(1) - tp)))


def kd_loss_multi(s, teachers, weights, T=1.0, need_grad=True):
    """KD term of trainer/ensemble_trainer.py:76-83: sum_k weights[k] * KLDivergenceLoss(T)(s, teachers[k]) and its
    gradient, as the weighted sum of the single-teacher oracle (losses/KLDiv.py:19-23)."""
    loss, ds = 0.0, (np.zeros_like(_c(s)) if need_grad else None)
    for t, w in zip(teachers, weights):
        l, g = kd_loss(s, t, T, False, need_grad)
        loss += float(w) * l
        if need_grad:
            ds += np.float32(w) * g
    return loss, ds
