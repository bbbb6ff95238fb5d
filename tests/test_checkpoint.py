"""Checkpoint layout and forgiving restore: same dict keys as base/base_trainer.py:170-177, same
partial-load rules as models/__init__.py:61-88, resume = surgery replay + restore (trainer/layerwise_trainer.py:404-427)."""
import os

import torch
from torch import nn

import kdcc
from kdcc import checkpoint as ck
from test_student_trainer import GEOM, PLAN, TinyTeacher, make_student


def test_checkpoint_dict_has_the_reference_keys(tmp_path):
    st = make_student()
    opt = torch.optim.SGD(st.trainable_parameters(), lr=0.1)
    path = ck.save_checkpoint(str(tmp_path / "checkpoint-epoch3.pth"), st, opt, epoch=3, monitor_best=0.5, config={"name": "x"}, save_best=True)
    state = torch.load(path, weights_only=False)
    assert set(state) == {"arch", "epoch", "state_dict", "optimizer", "monitor_best", "config"}
    assert state["arch"] == "DepthwiseStudent" and state["epoch"] == 3
    assert "student.body.0.separable_conv.weight" in state["state_dict"] and "teacher.body.0.weight" in state["state_dict"]
    assert os.path.exists(tmp_path / "model_best.pth")


def test_resume_replays_surgery_then_restores(tmp_path):
    st = make_student()
    with torch.no_grad():
        for p in st.trainable_parameters():
            p.add_(1.0)
    path = ck.save_checkpoint(str(tmp_path / "c.pth"), st, None, epoch=1)
    torch.manual_seed(0)
    fresh = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})   # un-operated student: teacher-shaped
    epoch = ck.resume(fresh, None, path, lambda i: [b for b in PLAN if b["epoch"] == i], **GEOM)
    assert epoch == 1 and fresh.replaced_block_names == st.replaced_block_names
    for (k, a), (_, b) in zip(st.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k


def test_forgiving_restore_skips_mismatches_and_strips_dataparallel_prefix():
    net = nn.Sequential(nn.Linear(4, 3), nn.Linear(3, 2))
    ref = {k: torch.full_like(v, 7.0) for k, v in net.state_dict().items()}
    ref["1.weight"] = torch.zeros(5, 3)                      # wrong size: skipped (different number of classes)
    ref["extra.weight"] = torch.zeros(1)                     # unknown key: ignored
    before = net.state_dict()["1.weight"].clone()
    ck.forgiving_state_restore(net, ref)
    assert float(net.state_dict()["0.weight"].mean()) == 7.0 and torch.equal(net.state_dict()["1.weight"], before)
    par = {"module." + k: torch.full_like(v, 3.0) for k, v in net.state_dict().items()}
    ck.forgiving_state_restore(net, par)                      # saved from nn.DataParallel: every key starts with module.
    assert float(net.state_dict()["1.bias"].mean()) == 3.0
    net2, _ = ck.restore_snapshot(nn.Sequential(nn.Linear(4, 3), nn.Linear(3, 2)), None, {"state_dict": net.state_dict()})
    assert torch.equal(net2.state_dict()["0.weight"], net.state_dict()["0.weight"])
