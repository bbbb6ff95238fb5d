"""Step-level parity against fixtures frozen from the reference's OWN student wrapper, teachers and losses
(oracle/make_golden_steps.py): the F10 residual-unit aliasing case and BASELINE config 1 end to end.

Each case runs twice: on the CPU with the replaced blocks restated by the oracle's torch port (pins the host mirror --
kdcc.DepthwiseStudent, hooks, prepare_train_epoch, ClassificationStep -- in the `-m "not gpu"` suite), and on the GPU
with the real kdcc blocks and losses through libkdcc.so.  fp32 throughout; the frozen convolutions around the blocks
are stock torch (cuDNN on the GPU), hence 1e-4 instead of 1e-5 on the GPU side."""
import os
from functools import reduce

import numpy as np
import pytest
import torch
from torch import nn

import kdcc
from conftest import GOLDEN
from harness_models import CifarResNet, ResidualTeacher
from oracle import torch_port as tp
from test_student_trainer import RefBlock


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _block_grads(blk):
    if isinstance(blk, RefBlock):
        return blk.w_dw.grad, blk.w_pw.grad
    return blk.separable_conv.weight.grad, blk.pointwise_conv.weight.grad


def _install_blocks(model, names, weights, device):
    """Copy the golden block weights into the replaced blocks; on the CPU swap them for the oracle port."""
    for n in names:
        blk = model.get_block(n, model.student)
        with torch.no_grad():
            blk.separable_conv.weight.copy_(torch.from_numpy(weights[n + ".separable_conv.weight"]))
            blk.pointwise_conv.weight.copy_(torch.from_numpy(weights[n + ".pointwise_conv.weight"]))
        if device == "cpu":
            model._set_block(n, RefBlock(blk), model.student)


# ---------------------------------------------------------------------------------------------------------------
# F10: hook-captured block output is mutated by the residual add_ (wider_resnet.py:180-182, depthwise_student.py:63-76)
# ---------------------------------------------------------------------------------------------------------------
def _residual_case(device):
    g = _load("residual.npz")
    k, p, d = [int(v) for v in g["geom"]]
    names = ["block1.convs.conv2", "block2.convs.conv2"]
    teacher = ResidualTeacher()
    teacher.load_state_dict({key[len("state/teacher."):]: torch.from_numpy(g[key]) for key in g.files if key.startswith("state/teacher.")})
    model = kdcc.DepthwiseStudent(teacher.to(device), {"trainer": {"verbosity": 2}})
    model.replace([{"name": n, "epoch": 1} for n in names], kernel_size=k, padding=p, dilation=d)
    model.register_hint_layers(names)
    model.unfreeze(names)
    weights = {key[len("state/student."):]: g[key] for key in g.files if key.startswith("state/student.")}
    _install_blocks(model, names, weights, device)
    if device == "cpu":                      # the hooks were attached to the blocks that just got swapped out
        model.register_hint_layers(names)
    x = torch.from_numpy(g["x"]).to(device)
    out_st, out_tc = model(x)
    if device == "cpu":
        crit = lambda s, t: tp.mse_loss(s, t, 1000)
    else:
        crit = kdcc.MSELoss(num_classes=1000)
    pairs = list(zip(model.student_hidden_outputs, model.teacher_hidden_outputs))
    assert len(pairs) == 2
    hint = reduce(lambda acc, e: acc + crit(e[0], e[1]), pairs, 0)
    hint.backward()
    tol = 1e-5 if device == "cpu" else 1e-4
    for i, (s, t) in enumerate(pairs):
        assert _rel(s.detach().cpu().numpy(), g["hidden_st/%d" % i]) < tol, i
        assert _rel(t.detach().cpu().numpy(), g["hidden_tc/%d" % i]) < tol, i
    assert abs(float(hint) - float(g["hint_loss"])) <= tol * abs(float(g["hint_loss"]))
    assert _rel(out_st.detach().cpu().numpy(), g["out_st"]) < tol
    for n in names:
        gd, gp = _block_grads(model.get_block(n, model.student))
        assert _rel(gd.cpu().numpy(), g["grad/student.%s.separable_conv.weight" % n]) < tol, n
        assert _rel(gp.cpu().numpy(), g["grad/student.%s.pointwise_conv.weight" % n]) < tol, n
    return model, pairs


def test_residual_alias_host_mirror_matches_reference():
    _residual_case("cpu")


@pytest.mark.gpu
def test_residual_alias_kdcc_blocks_match_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, pairs = _residual_case("cuda")
    # aliasing, explicitly: the tensor the hook stored is the tensor the next unit consumed, i.e. it already contains
    # the shortcut -- it differs from the raw block output by exactly that shortcut
    blk = model.get_block("block1.convs.conv2", model.student)
    assert isinstance(blk, kdcc.DepthwiseSeparableBlock)
    with torch.no_grad():
        stem = model.student.stem(torch.from_numpy(_load("residual.npz")["x"]).cuda())
        unit = model.student.block1
        pre = unit.bn1(stem.clone())
        raw = blk(unit.convs.bn2(unit.convs.conv1(pre)))
    hooked = pairs[0][0].detach()
    assert float((hooked - raw - stem).abs().max()) < 1e-4 * float(hooked.abs().max())
    assert float((hooked - raw).abs().max()) > 0.1 * float(stem.abs().max())


@pytest.mark.gpu
def test_residual_alias_bf16_tensor_core_path():
    """The same unit in bf16 NCHW (tensor-core depthwise, tcgen05 pointwise): the in-place add_ on the kernel's output
    and the backward through it stay within the bf16 tolerance of the reference's fp32 run."""
    g = _load("residual.npz")
    k, p, d = [int(v) for v in g["geom"]]
    names = ["block1.convs.conv2", "block2.convs.conv2"]
    teacher = ResidualTeacher()
    teacher.load_state_dict({key[len("state/teacher."):]: torch.from_numpy(g[key]) for key in g.files if key.startswith("state/teacher.")})
    model = kdcc.DepthwiseStudent(teacher.cuda(), {"trainer": {"verbosity": 2}})
    model.replace([{"name": n, "epoch": 1} for n in names], kernel_size=k, padding=p, dilation=d)
    model.register_hint_layers(names)
    model.unfreeze(names)
    weights = {key[len("state/student."):]: g[key] for key in g.files if key.startswith("state/student.")}
    _install_blocks(model, names, weights, "cuda")
    model = model.to(torch.bfloat16)
    for n in names:   # the blocks keep fp32 master weights (the checkpoint layout); activations are bf16
        model.get_block(n, model.student).float()
    x = torch.from_numpy(g["x"]).cuda().to(torch.bfloat16)
    model(x)
    pairs = list(zip(model.student_hidden_outputs, model.teacher_hidden_outputs))
    crit = kdcc.MSELoss(num_classes=1000)
    hint = reduce(lambda acc, e: acc + crit(e[0], e[1]), pairs, 0)
    hint.backward()
    for i, (s, _) in enumerate(pairs):
        assert not s.is_contiguous(memory_format=torch.channels_last) or s.shape[1] == 1
        assert _rel(s.detach().float().cpu().numpy(), g["hidden_st/%d" % i]) < 3e-2, i
    assert abs(float(hint) - float(g["hint_loss"])) <= 3e-2 * abs(float(g["hint_loss"]))
    for n in names:
        gd, gp = _block_grads(model.get_block(n, model.student))
        assert _rel(gd.float().cpu().numpy(), g["grad/student.%s.separable_conv.weight" % n]) < 5e-2, n
        assert _rel(gp.float().cpu().numpy(), g["grad/student.%s.pointwise_conv.weight" % n]) < 5e-2, n


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 1: cfg/cifar10/resnet44/config1.json, one step of trainer/classification_trainer.py:24-39
# ---------------------------------------------------------------------------------------------------------------
def _cifar_case(device):
    from kdcc.trainer import prepare_train_epoch
    g = _load("cifar_step.npz")
    k, p, d = [int(v) for v in g["geom"]]
    teacher = CifarResNet(7)
    sd = {key[len("teacher/"):]: torch.from_numpy(g[key]) for key in g.files if key.startswith("teacher/")}
    missing = teacher.load_state_dict(sd, strict=False)
    assert all(m.endswith("num_batches_tracked") for m in missing.missing_keys) and not missing.unexpected_keys
    model = kdcc.DepthwiseStudent(teacher.to(device), {"trainer": {"verbosity": 2}})
    pruning = {"pruning_plan": [{"name": str(n), "epoch": 1} for n in g["plan"]],
               "hint": [{"name": str(n), "epoch": 1} for n in g["hint"]],
               "unfreeze": [{"name": str(n), "epoch": 1} for n in g["unfreeze"]],
               "args": {"kernel_size": k, "padding": p, "dilation": d}}
    opt = prepare_train_epoch(model, pruning, 1, None, lambda ps: torch.optim.SGD(ps, lr=0.0))
    names = [str(n) for n in g["plan"]]
    assert len(names) == 8 and len(opt.param_groups[0]["params"]) == 32     # 8 blocks x 2 + 4 residual blocks x 2 BN x 2
    weights = {key[len("student/"):]: g[key] for key in g.files if key.startswith("student/")}
    _install_blocks(model, names, weights, device)
    if device == "cpu":   # the swapped-in port blocks are fresh modules: unfreeze them and rebuild the optimizer
        model.unfreeze([str(n) for n in g["unfreeze"]])
        opt = torch.optim.SGD(model.trainable_parameters(), lr=0.0)
    model.train()                                                            # classification_trainer.py:21
    assert model.student.training and not model.teacher.training
    if device == "cpu":
        crit = [nn.CrossEntropyLoss(ignore_index=255), lambda s, t: tp.kl_div_loss(s, t, float(g["T"])),
                lambda s, t: tp.mse_loss(s, t, 1)]
    else:
        crit = [nn.CrossEntropyLoss(ignore_index=255), kdcc.KLDivergenceLoss(temperature=float(g["T"])), kdcc.MSELoss(num_classes=1)]
    step = kdcc.ClassificationStep(model, crit, opt, accumulation_steps=2)   # 2: no optimizer step on index 0, so the
    x = torch.from_numpy(g["x"]).to(device)                                  # gradients (halved) are still in the bucket
    target = torch.randint(0, 10, (x.shape[0],), generator=torch.Generator().manual_seed(1)).to(device)
    out = step(x, target, batch_idx=0)
    tol = 1e-5 if device == "cpu" else 1e-4
    assert _rel(out["output_tc"].detach().cpu().numpy(), g["out_tc"]) < tol
    assert _rel(out["output_st"].detach().cpu().numpy(), g["out_st"]) < tol
    assert abs(2 * float(out["kd_loss"]) - float(g["kd_loss"])) <= tol * float(g["kd_loss"])
    assert abs(2 * float(out["hint_loss"]) - float(g["hint_loss"])) <= tol * float(g["hint_loss"])
    checked = 0
    named = dict(model.named_parameters())
    for key in g.files:
        if not key.startswith("grad/"):
            continue
        name = key[len("grad/"):]
        if name in named:
            grad = named[name].grad
        else:   # CPU run: the port block holds w_dw / w_pw instead of the two conv modules
            base, leaf = name.rsplit(".", 2)[0], name.rsplit(".", 2)[1]
            grad = named[base + (".w_dw" if leaf == "separable_conv" else ".w_pw")].grad
        assert grad is not None, name
        assert _rel(2 * grad.cpu().numpy(), g[key]) < 10 * tol, name
        checked += 1
    assert checked == 32
    return step


def test_cifar_config1_step_host_mirror_matches_reference():
    _cifar_case("cpu")


@pytest.mark.gpu
def test_cifar_config1_step_kdcc_matches_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    step = _cifar_case("cuda")
    # and the optimizer the config names (RAdam, utils/optim/radam.py) steps through the flat bucket, odd-sized BN
    # vectors included (every gradient view is 128-byte aligned)
    opt = kdcc.optim.RAdam(step.model.trainable_parameters(), lr=1e-3)
    step.optimizer = opt
    before = [p.detach().clone() for p in step.model.trainable_parameters()]
    x = torch.from_numpy(_load("cifar_step.npz")["x"]).cuda()
    step(x, torch.zeros(x.shape[0], dtype=torch.long, device="cuda"), batch_idx=1)   # (1 + 1) % 2 == 0: steps
    torch.cuda.synchronize()
    moved = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, step.model.trainable_parameters()))
    assert moved == len(before)
    assert float(step.bucket.flat.abs().sum()) == 0.0
