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


# ---- sliding-window test-time stitching -----------------------------------------------------
def tta_window_coordinates(w, h, tile, overlap=1 / 3):
    """Window list of utils/tta_process.py:81-105 (get_crops_image) for one scaled image of size (w, h): square
    tiles of `tile`, stride ceil(tile * (1 - overlap)), the last window of every axis pulled back inside the image,
    x outer / y inner.  Returns [(x1, y1, x2, y2)]."""
    import math
    stride = math.ceil(tile * (1 - overlap))
    nx = int(math.ceil((w - tile) / stride) + 1)
    ny = int(math.ceil((h - tile) / stride) + 1)
    out = []
    for ix in range(nx):
        for iy in range(ny):
            x2, y2 = min(ix * stride + tile, w), min(iy * stride + tile, h)
            out.append((max(x2 - tile, 0), max(y2 - tile, 0), x2, y2))
    return out


def tta_collect(w, h, coords, windows, count_mode="reference"):
    """Overlap-add of window outputs, restating utils/tta_process.py:39-52 (collect_windows_result).  The sum is the
    plain overlap-add (:49-51, a window is cropped to the part inside the image).  count_mode "reference" keeps the
    reference's counter as written (:46 indexes the (classes, h, w) counter with [y1:y2, x1:x2], i.e. classes by the
    y range and ROWS by the x range, all columns); "coverage" is the per-pixel coverage count.  float64 (C, h, w)."""
    windows = np.asarray(windows)
    C = windows.shape[1]
    full = np.zeros((C, h, w))
    cnt = np.zeros((C, h, w))
    for win, (x1, y1, x2, y2) in zip(windows, coords):
        if count_mode == "reference":
            cnt[y1:y2, x1:x2] += 1
        else:
            cnt[:, y1:y2, x1:x2] += 1
        full[:, y1:y2, x1:x2] += win[:, :y2 - y1, :x2 - x1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return full / cnt


def tta_resize_bilinear(planes, out_w, out_h):
    """cv2.resize(plane, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for every 2-D plane (utils/tta_process.py:
    29-36), restated: source coordinate (d + 0.5) * in/out - 0.5, taps floor and floor + 1 clamped to the plane, the
    fractional weight rounded to float32 (OpenCV keeps its interpolation table in float), arithmetic in float64.
    Same size = copy (as cv2); a tap of weight 0 is not read, so inf entries do not turn their neighbours into nan."""
    planes = np.asarray(planes, np.float64)
    C, h, w = planes.shape
    if (out_w, out_h) == (w, h):
        return planes.copy()

    def taps(n_in, n_out):
        f = ((np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5).astype(np.float32)
        i0 = np.floor(f).astype(np.int64)
        frac = (f - i0).astype(np.float32)
        frac[i0 < 0] = 0.0
        frac[i0 >= n_in - 1] = 0.0
        i0 = np.clip(i0, 0, n_in - 1)
        i1 = np.clip(i0 + 1, 0, n_in - 1)
        return i0, i1, frac.astype(np.float64)

    y0, y1, fy = taps(h, out_h)
    x0, x1, fx = taps(w, out_w)
    with np.errstate(invalid="ignore"):
        rows = np.where(fx > 0, planes[:, :, x0] * (1 - fx) + planes[:, :, x1] * fx, planes[:, :, x0])   # horizontal pass
        fy3 = fy[None, :, None]
        return np.where(fy3 > 0, rows[:, y0, :] * (1 - fy3) + rows[:, y1, :] * fy3, rows[:, y0, :])


def tta_reverse_mapping(mapping, results, ori_size, count_mode="reference"):
    """utils/tta_process.py:9-26 (reverse_mapping): per scale, stitch the plain and the mirrored windows, un-mirror
    the latter, resize both to ori_size = (w, h) and average.  results: all windows in order (plain then mirrored per
    scale).  Returns float64 (n_scales, C, h_ori, w_ori)."""
    results = np.asarray(results)
    out, idx = [], 0
    for w, h, coords in mapping:
        n = len(coords)
        plain = tta_collect(w, h, coords, results[idx:idx + n], count_mode)
        mirrored = tta_collect(w, h, coords, results[idx + n:idx + 2 * n], count_mode)[:, :, ::-1]
        out.append((tta_resize_bilinear(plain, *ori_size) + tta_resize_bilinear(mirrored, *ori_size)) / 2)
        idx += 2 * n
    return np.stack(out)


# ---- optimizer of the layerwise loop (cfg/cityscapes/*.json "optimizer": RAdam) -----------------------------------
def radam_scalars(step, beta1, beta2, degenerated_to_sgd=True):
    """(N_sma, step_size) of utils/optim/radam.py:65-84 for the 1-based step count (host scalars, float64)."""
    import math
    beta2_t = beta2 ** step
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma >= 5:
        size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - beta1 ** step)
    elif degenerated_to_sgd:
        size = 1.0 / (1 - beta1 ** step)
    else:
        size = -1
    return n_sma, size


def radam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, degenerated_to_sgd=True):
    """One step of utils/optim/radam.py:31-99 on fp32 arrays; `step` is the count AFTER the increment of :63.
    Returns the new (p, m, v)."""
    beta1, beta2 = betas
    p, g, m, v = (np.asarray(a, np.float32) for a in (p, g, m, v))
    v = (v * np.float32(beta2) + np.float32(1 - beta2) * g * g).astype(np.float32)
    m = (m * np.float32(beta1) + np.float32(1 - beta1) * g).astype(np.float32)
    n_sma, size = radam_scalars(step, beta1, beta2, degenerated_to_sgd)
    if n_sma >= 5:
        if weight_decay != 0:
            p = (p + np.float32(-weight_decay * lr) * p).astype(np.float32)
        p = (p + np.float32(-size * lr) * (m / (np.sqrt(v) + np.float32(eps)))).astype(np.float32)
    elif size > 0:
        if weight_decay != 0:
            p = (p + np.float32(-weight_decay * lr) * p).astype(np.float32)
        p = (p + np.float32(-size * lr) * m).astype(np.float32)
    return p, m, v


# ---- supervised loss that the layerwise loop logs twice per step (never back-propagated on this path) ------------
def cross_entropy_2d(logits, target, ignore_index=255):
    """losses/CrossEntropy.py:10-15 (NLLLoss2d(log_softmax(inputs)), mean over the pixels whose label is not
    ignore_index; nan when every pixel is ignored, as torch).  logits (N, C, H, W), target (N, H, W) int.  float64."""
    x = np.asarray(logits, np.float64)
    t = np.asarray(target).astype(np.int64)
    m = x.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(x - m).sum(axis=1))
    valid = t != ignore_index
    tc = np.where(valid, t, 0)
    picked = np.take_along_axis(x, tc[:, None], axis=1)[:, 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(((lse - picked) * valid).sum() / valid.sum())
