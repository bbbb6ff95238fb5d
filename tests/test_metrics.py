"""Segmentation metric: the numpy oracle is pinned on the reference's own CityscapesMetricTracker
(tests/golden/metrics.npz, frozen by oracle/make_golden.py from utils/util.py:57-128); the CUDA kernel must reproduce
the oracle bit for bit (integer counts)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


def test_oracle_confusion_matches_reference_tracker(golden):
    g = golden
    acc = np.zeros((19, 19), np.int64)
    for i in range(2):
        acc += orc.confusion(g["in%d/logits" % i], g["in%d/labels" % i], 19, 255)
        assert np.array_equal(acc, g["conf_after%d" % i])
        assert abs(orc.mean_iou(acc) - float(g["iou_after%d" % i])) < 1e-15
    assert orc.mean_iou(np.zeros((19, 19))) == float(g["iou_empty"]) == 1.0


def test_oracle_confusion_edge_cases():
    # every label ignored -> empty matrix; labels outside [0, C) are dropped like the reference's mask
    logits = np.random.RandomState(0).randn(1, 19, 4, 4).astype(np.float32)
    assert orc.confusion(logits, np.full((1, 4, 4), 255), 19, 255).sum() == 0
    lab = np.array([[[0, 18, 19, -1]]])
    assert orc.confusion(logits[:, :, :1, :], lab, 19, 255).sum() == 2


@pytest.mark.gpu
def test_confusion_kernel_matches_golden_and_tracker_api(golden):
    import torch
    import kdcc
    g = golden
    tr = kdcc.CityscapesMetricTracker()
    for i in range(2):
        logits, labels = torch.from_numpy(g["in%d/logits" % i]).cuda(), torch.from_numpy(g["in%d/labels" % i]).cuda()
        before = labels.clone()
        tr.update(logits, labels)
        assert torch.equal(labels, before)  # kdcc does not rewrite the caller's labels
        assert np.array_equal(tr.conf.astype(np.int64), g["conf_after%d" % i])
        assert abs(tr.get_iou() - float(g["iou_after%d" % i])) < 1e-12
    tr.reset()
    assert tr.get_iou() == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dtype", [((3, 19, 33, 47), "float32"), ((2, 10, 1, 5), "float32"), ((2, 19, 64, 64), "bfloat16"),
                                          ((1, 32, 16, 24), "float32"), ((4, 19, 256, 512), "float32")])
def test_confusion_kernel_matches_oracle_seeded(shape, dtype):
    import torch
    import kdcc
    rs = np.random.RandomState(sum(shape))
    N, C = shape[:2]
    logits = torch.from_numpy((3 * rs.standard_normal(shape)).astype(np.float32)).to(getattr(torch, dtype))
    logits[:, 1] = logits[:, 0]  # ties: first index wins
    labels = rs.randint(0, C, (N,) + shape[2:])
    labels[rs.rand(*labels.shape) < 0.1] = 255
    cm = kdcc.ConfusionMatrix(C, 255, device="cuda")
    cm.update(logits.cuda(), torch.from_numpy(labels).cuda())
    cm.update(logits.cuda(), torch.from_numpy(labels).cuda())
    ref = 2 * orc.confusion(logits.float().numpy(), labels, C, 255)
    assert np.array_equal(cm.numpy(), ref)
    assert abs(cm.iou() - float(np.nanmean(np.where(ref.sum(0) + ref.sum(1) - np.diag(ref) > 0,
                                                    np.diag(ref) / np.maximum(ref.sum(0) + ref.sum(1) - np.diag(ref), 1), np.nan)))) < 1e-12


@pytest.mark.gpu
def test_confusion_full_size_properties():
    """BASELINE size (4,19,1024,1024): counts add up to the valid pixels, a prediction that equals the labels gives a
    diagonal matrix, and two halves of the batch sum to the whole (additivity)."""
    import torch
    import kdcc
    torch.manual_seed(3)
    N, C, H, W = 4, 19, 1024, 1024
    labels = torch.randint(0, C, (N, H, W), device="cuda")
    labels[torch.rand(N, H, W, device="cuda") < 0.05] = 255
    logits = torch.randn(N, C, H, W, device="cuda")
    whole = kdcc.ConfusionMatrix(C, 255)
    whole.update(logits, labels)
    assert int(whole.mat.sum()) == int((labels != 255).sum())
    parts = kdcc.ConfusionMatrix(C, 255)
    parts.update(logits[:2], labels[:2])
    parts.update(logits[2:], labels[2:])
    assert torch.equal(parts.mat, whole.mat)
    onehot = torch.nn.functional.one_hot(labels.clamp(max=C - 1), C).permute(0, 3, 1, 2).float().contiguous()
    diag = kdcc.ConfusionMatrix(C, 255)
    diag.update(onehot, labels)
    m = diag.numpy()
    assert np.array_equal(m, np.diag(np.diag(m))) and diag.iou() == 1.0


def test_confusion_needs_cuda():
    import torch
    import kdcc
    cm_args = (torch.zeros(361, dtype=torch.long), torch.zeros(1, 19, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long))
    with pytest.raises(kdcc.KdccError):
        kdcc.functional.confusion_update(*cm_args)
