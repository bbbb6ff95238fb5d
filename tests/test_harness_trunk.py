"""The DeepLabV3+/WideResNet38 harness trunk (harness/deeplab_wrn38.py) keeps the reference's checkpoint layout and block
names: key/shape listing frozen from models/deeplabv3/deeplabv3.py:DeepWV3Plus (tests/golden/deepwv3plus_keys.json, written
by oracle/make_golden_steps.py), the survey's parameter-count known answers, and the 51M plan's surgery through the
drop-in student wrapper.  In the build container (where /root/reference exists) the forward pass is also compared with
the reference class itself."""
import json
import os
import sys
import types

import pytest
import torch

import kdcc
from conftest import GOLDEN
from harness.deeplab_wrn38 import DeepWV3Plus, PRUNING_51M, pruning_section


@pytest.fixture(scope="module")
def trunk():
    torch.manual_seed(0)
    return DeepWV3Plus(19).eval()


def test_state_dict_layout_matches_the_reference_class(trunk):
    want = json.load(open(os.path.join(GOLDEN, "deepwv3plus_keys.json")))
    have = {k: list(v.shape) for k, v in trunk.state_dict().items()}
    assert list(have) == list(want) and have == want
    assert sum(p.numel() for p in trunk.parameters()) == 137_103_936          # 137.10 M


def test_51m_plan_surgery_gives_the_survey_parameter_counts(trunk):
    from kdcc.trainer import prepare_train_epoch
    model = kdcc.DepthwiseStudent(trunk, {"trainer": {"verbosity": 2}})
    opt = prepare_train_epoch(model, pruning_section(), 1, None, lambda ps: torch.optim.SGD(ps, lr=0.1))
    assert model.replaced_block_names == PRUNING_51M["names"]
    sites = [(model.get_block(n, model.student).separable_conv.in_channels, model.get_block(n, model.student).pointwise_conv.out_channels)
             for n in PRUNING_51M["names"]]
    assert sites == [(512, 512)] * 5 + [(1024, 2048)] + [(4096, 256)] * 3       # what bench.py's PLAN_51M times
    assert sum(p.numel() for p in model.student.parameters()) == 85_960_768     # 85.96 M student
    assert sum(p.numel() for p in opt.param_groups[0]["params"]) == 7_839_232   # trainable
    assert not model.student.training                                            # F9: the student stays in eval mode


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference sources only exist in the build container")
def test_forward_matches_the_reference_class(trunk):
    cwd, path = os.getcwd(), list(sys.path)
    try:
        for name, attrs in (("beautifultable", {"BeautifulTable": type("BeautifulTable", (), {})}),
                            ("torchsummary", {"summary": lambda *a, **k: None}), ("tensorboardX", {"SummaryWriter": object})):
            if name not in sys.modules:
                m = types.ModuleType(name)
                m.__dict__.update(attrs)
                sys.modules[name] = m
        os.chdir("/root/reference")
        sys.path.insert(0, "/root/reference")
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
        from models.deeplabv3.deeplabv3 import DeepWV3Plus as Ref
        ref = Ref(num_classes=19).eval()
        ref.load_state_dict(trunk.state_dict())
        x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            assert torch.equal(ref(x), trunk(x))
    finally:
        os.chdir(cwd)
        sys.path[:] = path
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
