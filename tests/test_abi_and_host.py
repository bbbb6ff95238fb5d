"""CPU-side checks: the C-ABI library loads and exports every symbol include/kdcc.h declares, the
ctypes table mirrors the header, and the host mirror keeps the reference's module interface.  No
compute calls are made (there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kdcc():
    import __graft_entry__ as entry
    entry.build()
    import kdcc as pkg
    return pkg


def header_symbols():
    text = open(os.path.join(ROOT, "include", "kdcc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kdcc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(kdcc):
    lib = ctypes.CDLL(kdcc.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 17
    for name in syms:
        assert hasattr(lib, name), name
    assert sorted(kdcc._abi.SIGNATURES) == syms  # the ctypes table and the header agree


def test_version_and_error_strings(kdcc):
    L = kdcc._abi.lib()
    assert L.kdcc_version() == 104
    assert "success" in kdcc._abi.strerror(0)
    for code in (-1, -2, -3, -4, -5):
        assert kdcc._abi.strerror(code).startswith("kdcc:")
    assert L.kdcc_loss_workspace_bytes() > 0
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert L.kdcc_dw_fwd(None, None, None, None, 1, 8, 8, 16, 3, 1, 1, 0, 1, None) == -1
    assert L.kdcc_dw_fwd(None, None, None, None, 1, 8, 8, 6, 3, 1, 1, 0, 1, None) == -2     # C % 8 (NHWC)
    assert L.kdcc_dw_fwd(None, None, None, None, 1, 8, 8, 6, 3, 1, 1, 1, 0, None) == -2     # NCHW is bf16-only
    assert L.kdcc_pw_fwd(None, None, None, None, 0, None, None, 128, 64, 64, 1, 0, 7, None) == -1  # bad dtype
    assert L.kdcc_dw_bwd_workspace_bytes(1, 128, 128, 512, 9, 5, 20, 0, 1) > 0
    assert L.kdcc_dw_bwd_workspace_bytes(1, 128, 128, 512, 9, 5, 20, 1, 1) > 0


def test_dispatch_names(kdcc):
    d = kdcc._abi.dispatch_name
    NHWC, NCHW, BF16, F32 = kdcc._abi.NHWC, kdcc._abi.NCHW, kdcc._abi.BF16, kdcc._abi.F32
    assert d(0, 1, 128, 128, 512, 512, 9, 5, 20, NHWC, BF16) == "dw_conv_tma_k9"
    assert d(1, 1, 128, 128, 512, 512, 9, 5, 20, NHWC, BF16) == "dw_wgrad_tma_k9"
    assert d(0, 32, 8, 8, 64, 64, 3, 1, 1, NHWC, BF16) == "dw_conv_tma_k3"
    assert d(0, 1, 128, 128, 512, 512, 9, 5, 20, NHWC, F32) == "dw_direct"
    assert d(2, 1, 128, 128, 512, 512, 9, 5, 20, NHWC, BF16) == "pw_gemm_sm100_fwd"
    assert d(2, 1, 128, 128, 512, 512, 9, 5, 20, NHWC, F32) == "pw_simt"
    # the reference's NCHW layout: depthwise on the tensor cores, pointwise as W.X per image
    assert d(0, 1, 128, 128, 512, 512, 9, 5, 20, NCHW, BF16) == "dw_tc_conv"
    assert d(1, 1, 128, 128, 512, 512, 9, 5, 20, NCHW, BF16) == "dw_tc_wgrad"
    assert d(2, 1, 128, 128, 512, 512, 9, 5, 20, NCHW, BF16) == "pw_gemm_sm100_fwd"
    assert d(0, 1, 128, 128, 512, 512, 9, 5, 20, NCHW, F32) == "unsupported"
    assert d(0, 1, 30, 30, 16, 16, 3, 1, 1, NCHW, BF16) == "unsupported"       # W % 8 != 0 -> channels_last path


def test_block_keeps_reference_interface(kdcc):
    blk = kdcc.DepthwiseSeparableBlock(in_channels=16, out_channels=24, kernel_size=9, padding=20, dilation=5,
                                       groups=16, bias=None)
    sd = blk.state_dict()
    assert list(sd) == ["separable_conv.weight", "pointwise_conv.weight"]
    assert tuple(sd["separable_conv.weight"].shape) == (16, 1, 9, 9)
    assert tuple(sd["pointwise_conv.weight"].shape) == (24, 16, 1, 1)
    with_bias = kdcc.DepthwiseSeparableBlock(8, 8, 3, 1, 1, 8, True)
    assert "separable_conv.bias" in with_bias.state_dict() and "pointwise_conv.bias" in with_bias.state_dict()
    with pytest.raises(RuntimeError):
        kdcc.DepthwiseSeparableBlock(8, 8, 3, 1, 1, 8, torch.zeros(8))
    # no CPU fallback: the product path refuses host tensors instead of silently computing elsewhere
    with pytest.raises(kdcc.KdccError):
        blk(torch.randn(1, 16, 8, 8))


def test_losses_keep_reference_interface(kdcc):
    import kdcc.losses as losses
    for name, kwargs in (("KLDivergenceLoss", {"temperature": 5}), ("MSELoss", {"reduction": "mean", "num_classes": 1000}),
                         ("WeightedHintMSELoss", {"reduction": "mean", "num_classes": 19}), ("EnsembleKLDivergenceLoss", {})):
        mod = getattr(losses, name)(**kwargs)      # parse_config.ConfigParser.init_obj style construction
        assert isinstance(mod, torch.nn.Module)
    crit = torch.nn.ModuleList([losses.KLDivergenceLoss(2), losses.MSELoss(num_classes=1)])  # layerwise_trainer.py:62-63
    assert len(crit) == 2
    with pytest.raises(kdcc.KdccError):
        losses.KLDivergenceLoss()(torch.randn(2, 10), torch.randn(2, 10))
