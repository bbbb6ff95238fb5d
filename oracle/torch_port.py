"""torch-CPU restatement of the reference modules on the hot path.  TEST INFRASTRUCTURE ONLY.

The reference keeps its arithmetic in PyTorch library calls, so "the reference's own CPU
implementation of the path" is those same calls executed by the torch CPU backend.  This
module restates them as plain functions (it does not import /root/reference, which does
not exist on the GPU box) and is what bench.py times for `cpu_baseline` and
`--impl reference` with every host thread.  tests/test_oracle_golden.py pins it, together
with the C oracle, against fixtures produced by the reference modules themselves.

Call sites restated (reference file:line):
  block_forward      models/students/transform_blocks/depthwise_separable_conv.py:11-14
  kl_div_loss        losses/KLDiv.py:19-23
  ensemble_kl_loss   losses/EnsembleKLDiv.py:18-22
  weighted_hint_mse  losses/WeightedHintMSELoss.py:12-16
  mse_loss           losses/MSELoss.py:14-16
  layerwise_step     trainer/layerwise_trainer.py:223-235 (hot-path part of the loop body)
"""
import warnings

import torch
import torch.nn.functional as F


def block_forward(x, w_dw, w_pw, padding, dilation, b_dw=None, b_pw=None):
    channels = x.shape[1]
    mid = F.conv2d(x, w_dw, b_dw, stride=1, padding=padding, dilation=dilation, groups=channels)
    return F.conv2d(mid, w_pw, b_pw)


def kl_div_loss(inputs, targets, temperature=1.0):
    log_ps = F.log_softmax(inputs / temperature, dim=1)
    pt = F.softmax(targets / temperature, dim=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # reduction='mean' deprecation notice, as in the reference
        kl = F.kl_div(log_ps, pt, reduction="mean")
    return kl * (temperature ** 2) * targets.shape[1]


def ensemble_kl_loss(inputs, target_probs):
    log_ps = F.log_softmax(inputs, dim=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kl = F.kl_div(log_ps, target_probs, reduction="mean")
    return kl * target_probs.shape[1]


def weighted_hint_mse(inputs, targets, filter_weight):
    per_channel = ((inputs - targets) ** 2).mean(dim=(-1, -2))
    per_sample = (filter_weight * per_channel).sum(dim=-1) / filter_weight.sum(dim=-1)
    return per_sample.mean()


def mse_loss(inputs, targets, num_classes=19):
    return F.mse_loss(inputs, targets, reduction="mean") * num_classes


def layerwise_step(sites, logits_s, logits_t, kd_temperature=1.0, hint_num_classes=1000):
    """One pass of the hot path as LayerwiseTrainer drives it, on CPU tensors.

    sites: list of dicts {x, w_dw, w_pw, teacher, padding, dilation}; w_* require grad.
    Runs every cheap-conv block forward, the hint loss against the teacher feature,
    the (logged) KD loss on the logits, and backward of the summed hint loss
    (trainer/layerwise_trainer.py:229-235 back-propagates the hint loss only).
    Returns (hint_loss, kd_loss) as floats.
    """
    hint = 0
    for s in sites:
        y = block_forward(s["x"], s["w_dw"], s["w_pw"], s["padding"], s["dilation"])
        hint = hint + mse_loss(y, s["teacher"], hint_num_classes)
    with torch.no_grad():
        kd = kl_div_loss(logits_s, logits_t, kd_temperature)
    hint.backward()
    return float(hint), float(kd)


def ensemble_kd_loss(output_st, ensemble_outputs, output_tc, temperature=1.0, weight=1.0, accumulation_steps=1):
    """trainer/ensemble_trainer.py:81-83: kd_loss over the ensemble members and the own teacher."""
    from functools import reduce
    kd = reduce(lambda acc, elem: acc + weight * kl_div_loss(output_st, elem, temperature), ensemble_outputs, 0)
    kd = kd + kl_div_loss(output_st, output_tc, temperature)
    return kd / (weight * len(ensemble_outputs) + 1) / accumulation_steps
