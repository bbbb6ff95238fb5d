"""DepthwiseSeparableBlock -- the student's cheap-conv replacement block on libkdcc (sm_100a).

Drop-in for models/students/transform_blocks/depthwise_separable_conv.py:4-14 of the reference:
same constructor signature, same sub-module names, so `state_dict()` keys
(`separable_conv.weight (C,1,k,k)`, `pointwise_conv.weight (Co,C,1,1)`, `+ .bias`) and default
initialisation match existing checkpoints and `forgiving_state_restore`.  Only `forward` differs: the
two convolutions run in the hand-written kernels instead of ATen/cuDNN.
"""
import torch
from torch import nn

from . import functional as F_kdcc


class DepthwiseSeparableBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding, dilation, groups, bias, use_cuda=True):
        super().__init__()
        # nn.Conv2d instances hold (and initialise) the parameters exactly like the reference; passing a
        # Tensor as `bias` raises the same RuntimeError the reference raises.
        self.separable_conv = nn.Conv2d(in_channels, in_channels, kernel_size, padding=padding, dilation=dilation,
                                        groups=groups, bias=bias)
        self.pointwise_conv = nn.Conv2d(in_channels, out_channels, 1, bias=bias)
        if groups != in_channels:
            raise ValueError("DepthwiseSeparableBlock: the kdcc kernels implement the depthwise case groups == in_channels "
                             "(the only one DepthwiseStudent.replace builds), got groups=%d" % groups)
        k, p, d = self.separable_conv.kernel_size, self.separable_conv.padding, self.separable_conv.dilation
        if k[0] != k[1] or p[0] != p[1] or d[0] != d[1]:
            raise ValueError("DepthwiseSeparableBlock: square kernel / symmetric padding and dilation only")
        self._k, self._p, self._d = k[0], p[0], d[0]

    def forward(self, x):
        dw, pw = self.separable_conv, self.pointwise_conv
        # a channels_last caller gets a channels_last result whatever layout the kernels ran in between
        cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        x = F_kdcc.depthwise_conv(x, dw.weight, dw.bias, self._k, self._d, self._p)
        x = F_kdcc.pointwise_conv(x, pw.weight, pw.bias, out_channels_last=cl)
        return x

    @torch.no_grad()
    def forward_fused_bn_relu(self, x, bn, relu=True):
        """Inference-only: pointwise conv with a following eval-mode BatchNorm (+ReLU) folded into the
        GEMM epilogue (valid on the Layerwise path where the student stays in eval mode)."""
        if bn.training:
            raise RuntimeError("BN fold needs eval-mode BatchNorm (running statistics)")
        dw, pw = self.separable_conv, self.pointwise_conv
        inv = torch.rsqrt(bn.running_var.float() + bn.eps)
        gamma = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
        beta = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
        scale = (gamma * inv).contiguous()
        shift = beta - bn.running_mean.float() * scale
        if pw.bias is not None:
            shift = shift + pw.bias.float() * scale
        cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        x = F_kdcc.depthwise_conv(x, dw.weight, dw.bias, self._k, self._d, self._p)
        return F_kdcc.pointwise_conv(x, pw.weight, None, scale, shift.contiguous(), relu, out_channels_last=cl)
