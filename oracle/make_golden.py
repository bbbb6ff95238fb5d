"""Freeze golden vectors from the REFERENCE's own modules (run in the build container only).

    python oracle/make_golden.py          # writes tests/golden/*.npz

The reference ships no tests or fixtures, so parity is pinned on outputs of
its modules executed here on the torch CPU backend: DepthwiseSeparableBlock
(models/students/transform_blocks/depthwise_separable_conv.py), KLDivergenceLoss,
EnsembleKLDivergenceLoss, WeightedHintMSELoss, MSELoss (losses/*.py), plus real CIFAR-10
ResNet44 teacher logits from checkpoints/cifar10/resnet44.th.  Modules are loaded by file
path so none of the reference's package-level side effects run.  /root/reference is not
available on the GPU box; only the .npz files travel.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("KDCC_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def block_case(ref_block_cls, tag, N, Ci, Co, H, W, k, d, p, seed, dtype=torch.float32):
    torch.manual_seed(seed)
    blk = ref_block_cls(Ci, Co, k, p, d, Ci, None).to(dtype)
    # spread the weights a little so that every tap matters
    with torch.no_grad():
        blk.separable_conv.weight.mul_(2.0)
    x = torch.randn(N, Ci, H, W, dtype=dtype, requires_grad=True)
    y = blk(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    f = lambda t: t.detach().to(torch.float32).numpy()
    return {
        f"{tag}/geom": np.array([N, Ci, Co, H, W, k, d, p], np.int64),
        f"{tag}/x": f(x), f"{tag}/w_dw": f(blk.separable_conv.weight), f"{tag}/w_pw": f(blk.pointwise_conv.weight),
        f"{tag}/y": f(y), f"{tag}/dy": f(dy), f"{tag}/dx": f(x.grad),
        f"{tag}/dw_dw": f(blk.separable_conv.weight.grad), f"{tag}/dw_pw": f(blk.pointwise_conv.weight.grad),
    }


def loss_case(tag, module, args, grads_of=0):
    args = [a.clone().requires_grad_(i == grads_of) for i, a in enumerate(args)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = module(*args)
    loss.backward()
    out = {f"{tag}/loss": np.array(loss.item(), np.float64), f"{tag}/grad": args[grads_of].grad.numpy()}
    for i, a in enumerate(args):
        out[f"{tag}/arg{i}"] = a.detach().numpy()
    return out


def ensemble_golden():
    """The reference has no module for the multi-teacher KD term: it is the three lines
    trainer/ensemble_trainer.py:81-83 over its KLDivergenceLoss module, reproduced here with that module."""
    from functools import reduce
    kl = _load("ref_kl", "losses/KLDiv.py").KLDivergenceLoss
    WEIGHT = 1  # trainer/ensemble_trainer.py:10
    out = {}
    for tag, shape, T, K in (("seg_T1", (2, 19, 9, 11), 1, 3), ("cifar_T5", (32, 10), 5, 2), ("seg_T2_w", (1, 19, 8, 8), 2, 4)):
        torch.manual_seed(41 + K)
        crit = kl(temperature=T)
        s = (2 * torch.randn(shape)).requires_grad_(True)
        outputs = [2 * torch.randn(shape) for _ in range(K)]
        tc = 2 * torch.randn(shape)
        kd = reduce(lambda acc, elem: acc + WEIGHT * crit(s, elem), outputs, 0)
        kd = kd + crit(s, tc)
        kd = kd / (WEIGHT * len(outputs) + 1)
        kd.backward()
        out[tag + "/s"] = s.detach().numpy()
        out[tag + "/teachers"] = np.stack([o.numpy() for o in outputs] + [tc.numpy()])
        out[tag + "/T"] = np.array(float(T))
        out[tag + "/loss"] = np.array(float(kd))
        out[tag + "/ds"] = s.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ensemble.npz"), **out)


def metrics_golden():
    """utils/util.py CityscapesMetricTracker run on seeded logits/labels (ignore pixels, exact ties, two updates)."""
    util = _load("ref_util", "utils/util.py")
    out = {}
    torch.manual_seed(31)
    tr = util.CityscapesMetricTracker()
    cases = []
    for i, (n, h, w) in enumerate([(2, 37, 53), (1, 64, 64)]):
        logits = 3 * torch.randn(n, 19, h, w)
        logits[:, 5] = logits[:, 3]            # exact ties between two classes: the first index must win
        logits[0, :, 0, :7] = 0.25             # all classes tied
        labels = torch.randint(0, 19, (n, h, w))
        labels[torch.rand(n, h, w) < 0.07] = 255
        cases.append((logits, labels))
        out["in%d/logits" % i], out["in%d/labels" % i] = logits.numpy(), labels.numpy().copy()
        tr.update(logits, labels.clone())
        out["conf_after%d" % i] = tr.conf.astype(np.int64)
        out["iou_after%d" % i] = np.array(tr.get_iou())
    tr.reset()
    out["iou_empty"] = np.array(tr.get_iou())
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


def tta_golden():
    """utils/tta_process.py run end to end on small seeded images: window list, crops and reverse_mapping of seeded
    window outputs (scale 1.0 as in every shipped config, one equal-tile two-scale case).  The module predates
    numpy 1.24 (`np.float`, :51), so that alias is restored before it is loaded; nothing else is shimmed."""
    from PIL import Image
    np.float = float
    ref = _load("ref_tta", "utils/tta_process.py")
    mean_std = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    rs = np.random.RandomState(77)
    out = {}
    #          W   H  crop scales     classes
    cases = [(60, 36, 24, [1.0], 5),       # 4 x 2 windows, fewer classes than tile rows: the reference counter covers every class
             (48, 48, 24, [1.0], 19),      # square, 3 x 3 windows
             (50, 30, 28, [1.0], 7),       # windows pulled back at both borders
             (40, 20, 20, [1.0, 1.0], 4),  # two scales (equal tiles, the only multi-scale form the reference can batch)
             (36, 18, 18, [1.0], 24)]      # more classes than the tile is high: rows of the counter stay 0 -> inf / nan
    for i, (W, H, crop, scales, C) in enumerate(cases):
        img = Image.fromarray(rs.randint(0, 255, (H, W, 3)).astype(np.uint8))
        ori, mapping, windows = ref.get_crops_image(ref.scale_and_flip_image(img, mean_std, scales), scales, crop_size=crop)
        # window outputs are NOT stored: tests regenerate them from this seed (legacy RandomState streams are stable)
        seed = 1000 + i
        res = np.random.RandomState(seed).standard_normal((windows.shape[0], C, windows.shape[2], windows.shape[3])).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            full = ref.reverse_mapping(mapping, res, ori)
        out["c%d/image" % i] = np.asarray(img)
        out["c%d/args" % i] = np.array([W, H, crop, C, len(scales), seed, windows.shape[0]], np.int64)
        out["c%d/scales" % i] = np.array(scales, np.float64)
        if i == 0:
            out["c%d/windows" % i] = windows.numpy()   # pins the crop order / mirroring of get_crops_image
        out["c%d/full" % i] = full.astype(np.float32)
        for k, (w, h, boxes) in enumerate(mapping):
            out["c%d/map%d" % (i, k)] = np.array([[w, h, 0, 0]] + [list(b) for b in boxes], np.int64)
    np.savez_compressed(os.path.join(OUT, "tta.npz"), **out)


def radam_golden():
    """utils/optim/radam.py stepped 9 times on seeded parameters / gradients (the first 5 steps take its
    degenerated-to-SGD branch), with and without weight decay and with degenerated_to_sgd off."""
    RAdam = _load("ref_radam", "utils/optim/radam.py").RAdam
    out = {}
    cfgs = [("plain", dict(lr=5e-3)),                                   # cfg/cityscapes/51M_deeplab_all.json:64-69
            ("wd", dict(lr=1e-2, weight_decay=1e-2, betas=(0.8, 0.99))),
            ("nosgd", dict(lr=5e-3, degenerated_to_sgd=False))]
    for tag, kw in cfgs:
        torch.manual_seed(5)
        prm = torch.nn.Parameter(torch.randn(257))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            opt = RAdam([prm], **kw)
            out[tag + "/p0"] = prm.detach().numpy().copy()
            grads, ps, ms, vs = [], [], [], []
            for t in range(9):
                prm.grad = torch.randn(257) * (1 + t)
                grads.append(prm.grad.numpy().copy())
                opt.step()
                ps.append(prm.detach().numpy().copy())
                ms.append(opt.state[prm]["exp_avg"].numpy().copy())
                vs.append(opt.state[prm]["exp_avg_sq"].numpy().copy())
        out[tag + "/grads"], out[tag + "/p"], out[tag + "/m"], out[tag + "/v"] = map(np.stack, (grads, ps, ms, vs))
        out[tag + "/hyper"] = np.array([kw.get("lr"), kw.get("betas", (0.9, 0.999))[0], kw.get("betas", (0.9, 0.999))[1],
                                        1e-8, kw.get("weight_decay", 0.0), float(kw.get("degenerated_to_sgd", True))], np.float64)
    np.savez_compressed(os.path.join(OUT, "radam.npz"), **out)


def ce_golden():
    """losses/CrossEntropy.py CrossEntropyLoss2d on seeded logits / labels with ignored pixels (the supervised and
    teacher losses that trainer/layerwise_trainer.py:225,232 evaluates every step for logging)."""
    CE = _load("ref_ce", "losses/CrossEntropy.py").CrossEntropyLoss2d
    out = {}
    torch.manual_seed(41)
    for i, (n, c, h, w, frac) in enumerate([(2, 19, 17, 23, 0.1), (1, 19, 32, 32, 0.0), (3, 10, 5, 7, 0.5)]):
        logits = 4 * torch.randn(n, c, h, w)
        labels = torch.randint(0, c, (n, h, w))
        labels[torch.rand(n, h, w) < frac] = 255
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss = CE(ignore_index=255)(logits, labels)
        out["c%d/logits" % i], out["c%d/labels" % i] = logits.numpy(), labels.numpy()
        out["c%d/loss" % i] = np.array(loss.item(), np.float64)
    np.savez_compressed(os.path.join(OUT, "ce.npz"), **out)


def block_extra_golden():
    """Round 2: Cityscapes-geometry cases the NCHW tensor-core kernels accept (W % 8 == 0), so that path is compared
    with the reference's own fp32 block -- taps NOT pre-rounded to bf16 -- and one full 128 x 128 plane.
        python oracle/make_golden.py extra"""
    Block = _load("ref_dsc", "models/students/transform_blocks/depthwise_separable_conv.py").DepthwiseSeparableBlock
    blocks = {}
    #                 tag              N  Ci  Co    H    W  k  d   p  seed
    for spec in [("city_k9d5_w32",     2, 16, 24,  40,  32, 9, 5, 20, 17),   # small planes, image pair
                 ("city_k9d5_plane",   1,  8,  8, 128, 128, 9, 5, 20, 18)]:  # a whole 128 x 128 plane (1024^2 crop)
        blocks.update(block_case(Block, *spec))
    np.savez_compressed(os.path.join(OUT, "block_extra.npz"), **blocks)
    print("block_extra.npz", os.path.getsize(os.path.join(OUT, "block_extra.npz")), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic reduction order
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        return block_extra_golden()
    blockmod = _load("ref_dsc", "models/students/transform_blocks/depthwise_separable_conv.py")
    Block = blockmod.DepthwiseSeparableBlock
    kl = _load("ref_kl", "losses/KLDiv.py").KLDivergenceLoss
    ekl = _load("ref_ekl", "losses/EnsembleKLDiv.py").EnsembleKLDivergenceLoss
    whm = _load("ref_whm", "losses/WeightedHintMSELoss.py").WeightedHintMSELoss
    mse = _load("ref_mse", "losses/MSELoss.py").MSELoss

    # ---- cheap-conv block: forward + backward -------------------------------------------
    blocks = {}
    #                 tag            N  Ci  Co   H   W  k  d   p  seed
    for spec in [("cifar_k3",        4, 64, 64,  8,  8, 3, 1,  1, 11),   # cfg/cifar10/resnet44/config1.json:84-89
                 ("city_k9d5",       1, 16, 24, 26, 31, 9, 5, 20, 12),   # cfg/cityscapes/51M_deeplab_all.json:117-122
                 ("ragged_k3d2",     2,  8, 16,  9, 12, 3, 2,  2, 13),
                 ("k5",              1,  8,  8, 10, 10, 5, 1,  2, 14),
                 ("shrink_k3p0",     2,  8, 24,  7, 11, 3, 1,  0, 15),   # output smaller than input
                 ("onepix_k1",       3, 16,  8,  5,  4, 1, 1,  0, 16)]:
        blocks.update(block_case(Block, *spec))
    # fp64 run of the Cityscapes geometry: tolerance anchor for the fp32 comparisons
    blocks.update(block_case(Block, "city_k9d5_f64", 1, 16, 24, 26, 31, 9, 5, 20, 12, dtype=torch.float64))
    np.savez_compressed(os.path.join(OUT, "block.npz"), **blocks)

    # ---- losses ---------------------------------------------------------------------------
    losses = {}
    torch.manual_seed(21)
    s, t = 3 * torch.randn(2, 19, 12, 13), 3 * torch.randn(2, 19, 12, 13)
    for T in (1, 2, 5):
        losses.update(loss_case(f"kl_T{T}", kl(temperature=T), [s, t]))
    losses[f"kl_T1/T"] = np.array(1.0); losses["kl_T2/T"] = np.array(2.0); losses["kl_T5/T"] = np.array(5.0)
    # extreme logits: softmax must be max-subtracted
    big_s, big_t = 40 * torch.randn(3, 19, 5, 7), 40 * torch.randn(3, 19, 5, 7)
    losses.update(loss_case("kl_big", kl(temperature=1), [big_s, big_t])); losses["kl_big/T"] = np.array(1.0)

    # real CIFAR-10 ResNet44 teacher logits (cfg/cifar10/resnet44/config1.json: KLDivergenceLoss T=5, batch 32)
    resnet = _load("ref_cifar_resnet", "models/cifar_models/resnet.py")
    teacher = resnet.resnet44()
    ck = torch.load(os.path.join(REF, "checkpoints/cifar10/resnet44.th"), map_location="cpu", weights_only=False)
    teacher.load_state_dict({k.replace("module.", "", 1): v for k, v in ck["state_dict"].items()})
    teacher.eval()
    torch.manual_seed(0)
    with torch.no_grad():
        t_logits = teacher(torch.randn(32, 3, 32, 32))
    s_logits = t_logits + 0.5 * torch.randn_like(t_logits)
    losses.update(loss_case("kl_cifar_T5", kl(temperature=5), [s_logits, t_logits])); losses["kl_cifar_T5/T"] = np.array(5.0)

    # ensemble KL: targets are probabilities (mean of 3 teachers' softmaxes); one variant has exact zeros
    torch.manual_seed(22)
    s = 2 * torch.randn(2, 19, 8, 8)
    probs = torch.stack([torch.softmax(2 * torch.randn(2, 19, 8, 8), 1) for _ in range(3)]).mean(0)
    losses.update(loss_case("ekl", ekl(), [s, probs]))
    onehot = torch.zeros(2, 19, 8, 8).scatter_(1, torch.randint(0, 19, (2, 1, 8, 8)), 1.0)
    losses.update(loss_case("ekl_onehot", ekl(), [s, onehot]))

    # hint losses
    torch.manual_seed(23)
    s, t = torch.randn(2, 16, 8, 9), torch.randn(2, 16, 8, 9)
    w1, w2 = torch.rand(16), torch.rand(2, 16)
    losses.update(loss_case("whint_vec", whm(), [s, t, w1]))
    losses.update(loss_case("whint_tab", whm(), [s, t, w2]))
    losses.update(loss_case("mse_nc1000", mse(num_classes=1000), [s, t])); losses["mse_nc1000/nc"] = np.array(1000.0)
    losses.update(loss_case("mse_nc1", mse(num_classes=1), [s, t])); losses["mse_nc1/nc"] = np.array(1.0)
    s4, t4 = torch.randn(3, 24, 1, 1), torch.randn(3, 24, 1, 1)          # degenerate 1x1 maps
    losses.update(loss_case("mse_1x1", mse(num_classes=19), [s4, t4])); losses["mse_1x1/nc"] = np.array(19.0)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **losses)

    metrics_golden()
    ensemble_golden()
    tta_golden()
    radam_golden()
    ce_golden()
    for fn in ("block.npz", "losses.npz", "metrics.npz", "ensemble.npz", "tta.npz", "radam.npz", "ce.npz"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)), "bytes")


if __name__ == "__main__":
    sys.exit(main())
