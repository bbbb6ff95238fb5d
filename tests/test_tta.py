"""Sliding-window test-time stitching.  The numpy oracle is pinned on the reference's own
utils/tta_process.py run end to end on seeded images (tests/golden/tta.npz, frozen by oracle/make_golden.py); the
host mirror must cut the same windows in the same order, and the CUDA kernels must reproduce the reference's maps
-- including the inf / nan its window counter produces when there are more classes than tile rows -- to 1e-5 relative
(fp32 on the device, float64 in the reference)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tta.npz")
CASES = range(5)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


def _case(g, i):
    W, H, crop, C, n_scales, seed, n_win = [int(v) for v in g["c%d/args" % i]]
    mapping = []
    for k in range(n_scales):
        m = g["c%d/map%d" % (i, k)]
        mapping.append([int(m[0, 0]), int(m[0, 1]), [tuple(int(v) for v in row) for row in m[1:]]])
    results = np.random.RandomState(seed).standard_normal((n_win, C, crop, crop)).astype(np.float32)
    return W, H, crop, C, [float(s) for s in g["c%d/scales" % i]], mapping, results, g["c%d/full" % i]


def _same(mine, want, tol=1e-5):
    """equal where the reference is finite (relative to its largest finite value), same inf / nan pattern elsewhere"""
    mine, want = np.asarray(mine, np.float64), np.asarray(want, np.float64)
    assert mine.shape == want.shape
    assert np.array_equal(np.isnan(mine), np.isnan(want))
    assert np.array_equal(np.isposinf(mine), np.isposinf(want)) and np.array_equal(np.isneginf(mine), np.isneginf(want))
    fin = np.isfinite(want)
    scale = np.abs(want[fin]).max() if fin.any() else 1.0
    assert np.abs(mine[fin] - want[fin]).max() <= tol * scale


@pytest.mark.parametrize("i", CASES)
def test_oracle_matches_reference_reverse_mapping(golden, i):
    W, H, crop, C, scales, mapping, results, full = _case(golden, i)
    for (w, h, boxes), s in zip(mapping, scales):
        assert orc.tta_window_coordinates(w, h, int(s * crop)) == boxes
    with np.errstate(divide="ignore", invalid="ignore"):
        _same(orc.tta_reverse_mapping(mapping, results, (W, H)), full, 1e-6)   # golden is stored as fp32


def test_oracle_coverage_mode_is_a_partition_of_unity():
    # windows that all hold the same constant stitch back to that constant when divided by the true coverage
    boxes = orc.tta_window_coordinates(50, 30, 28)
    ones = np.full((len(boxes), 3, 28, 28), 2.5, np.float32)
    assert np.allclose(orc.tta_collect(50, 30, boxes, ones, "coverage"), 2.5)
    ref = orc.tta_collect(50, 30, boxes, ones, "reference")
    assert not np.allclose(ref, 2.5)   # the reference counter does not normalise the overlaps


def test_oracle_resize_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    x = np.random.RandomState(3).standard_normal((3, 37, 53))
    for ow, oh in [(53, 37), (100, 80), (20, 11), (53, 74)]:
        want = np.stack([cv2.resize(p, (ow, oh), interpolation=cv2.INTER_LINEAR) for p in x])
        assert np.abs(want - orc.tta_resize_bilinear(x, ow, oh)).max() < 2e-5   # identity / 2x: exact


def test_host_mirror_cuts_the_reference_windows(golden):
    """kdcc.tta.scale_and_flip_image / get_crops_image on the stored image: same boxes, same crops, same order."""
    pytest.importorskip("torchvision")
    from PIL import Image
    from kdcc import tta
    g = golden
    W, H, crop, C, scales, mapping, results, full = _case(g, 0)
    img = Image.fromarray(g["c0/image"])
    data = tta.scale_and_flip_image(img, ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]), scales)
    ori, mine_map, windows = tta.get_crops_image(data, scales, crop_size=crop)
    assert ori == (W, H)
    assert [[w, h, list(b)] for w, h, b in mine_map] == [[w, h, list(b)] for w, h, b in mapping]
    assert np.array_equal(windows.numpy(), g["c0/windows"])
    for i in CASES:   # window lists of every case
        W, H, crop, C, scales, mapping, _, _ = _case(g, i)
        for (w, h, boxes), s in zip(mapping, scales):
            assert tta.window_coordinates(w, h, int(s * crop)) == boxes


@pytest.mark.gpu
@pytest.mark.parametrize("i", CASES)
def test_reverse_mapping_kernels_match_reference(golden, i):
    import torch
    from kdcc import tta
    W, H, crop, C, scales, mapping, results, full = _case(golden, i)
    out = tta.reverse_mapping(mapping, torch.from_numpy(results).cuda(), (W, H))
    _same(out.cpu().numpy(), full)


@pytest.mark.gpu
def test_stitch_modes_resize_and_accumulate_match_oracle():
    import torch
    import kdcc
    rs = np.random.RandomState(5)
    w, h, tile, C = 70, 44, 32, 6
    boxes = orc.tta_window_coordinates(w, h, tile)
    win = rs.standard_normal((len(boxes), C, tile, tile)).astype(np.float32)
    coords = torch.tensor(boxes, dtype=torch.int32).cuda()
    dwin = torch.from_numpy(win).cuda()
    for mode in ("reference", "coverage"):
        want = orc.tta_collect(w, h, boxes, win, mode)
        out = torch.empty((C, h, w), device="cuda")
        kdcc.functional.tta_stitch(dwin, coords, h, w, out, count_mode=mode)
        _same(out.cpu().numpy(), want)
        kdcc.functional.tta_stitch(dwin, coords, h, w, out, flip=True, count_mode=mode, alpha=0.25, accumulate=True)
        _same(out.cpu().numpy(), want + 0.25 * want[:, :, ::-1])
    # resize to a different size (a scale != 1 pass) against the oracle's cv2 restatement, then through reverse_mapping
    src = rs.standard_normal((C, h, w)).astype(np.float32)
    for OW, OH in [(w, h), (100, 63), (35, 22), (70, 88)]:
        dst = torch.zeros((C, OH, OW), device="cuda")
        kdcc.functional.resize_bilinear(torch.from_numpy(src).cuda(), dst)
        _same(dst.cpu().numpy(), orc.tta_resize_bilinear(src, OW, OH))
    res = rs.standard_normal((2 * len(boxes), C, tile, tile)).astype(np.float32)
    mapping = [[w, h, boxes]]
    out = kdcc.tta.reverse_mapping(mapping, torch.from_numpy(res).cuda(), (105, 66), count_mode="coverage")
    _same(out.cpu().numpy(), orc.tta_reverse_mapping(mapping, res, (105, 66), "coverage"))


@pytest.mark.gpu
def test_inference_test_runs_student_windows_on_device(golden):
    """DepthwiseStudent.inference_test end to end with a toy student (1x1 conv): equals the oracle fed with the same
    window outputs."""
    pytest.importorskip("torchvision")
    import torch
    from kdcc import tta
    torch.manual_seed(0)
    net = torch.nn.Conv2d(3, 5, 1).cuda()
    data = torch.rand(2, 3, 36, 60, device="cuda")   # ToPILImage expects [0, 1]
    args = {"scales": [1.0], "crop_size": 24}
    out = tta.inference_test(net, data, args)
    assert out.shape == (2, 5, 36, 60) and out.is_cuda
    from torchvision import transforms
    for n in range(2):
        image_data = tta.scale_and_flip_image(transforms.ToPILImage()(data[n].cpu()), ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]), [1.0])
        ori, mapping, windows = tta.get_crops_image(image_data, [1.0], crop_size=24)
        with torch.no_grad():
            res = net(windows.cuda()).cpu().numpy()
        want = orc.tta_reverse_mapping(mapping, res, ori).mean(axis=0)
        _same(out[n].cpu().numpy(), want)


def test_tta_entry_points_refuse_host_tensors():
    """No CPU fallback: the stitching kernels are CUDA only and say so."""
    import torch
    import kdcc
    win = torch.zeros(2, 3, 4, 4)
    coords = torch.zeros(2, 4, dtype=torch.int32)
    out = torch.zeros(3, 6, 6)
    with pytest.raises(kdcc.KdccError):
        kdcc.functional.tta_stitch(win, coords, 6, 6, out)
    with pytest.raises(kdcc.KdccError):
        kdcc.functional.resize_bilinear(out, torch.zeros(3, 8, 8))
