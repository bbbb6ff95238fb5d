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


def _pruning(two_epochs=False):
    """The config's "pruning" section (cfg/cityscapes/*.json): plan / hint / unfreeze lists with epochs + default args."""
    plan = [dict(b) for b in PLAN]
    if two_epochs:
        plan[1]["epoch"] = 2
    names = [{"name": b["name"], "epoch": b["epoch"]} for b in plan]
    unfreeze = [dict(n) for n in names]
    if two_epochs:
        unfreeze[1]["lr"] = 0.02      # per-layer learning rate (layerwise_trainer.py:171-172)
    return {"pruning_plan": plan, "hint": names, "unfreeze": unfreeze, "args": dict(GEOM)}


def test_resume_replays_surgery_then_restores(tmp_path):
    st = make_student()
    with torch.no_grad():
        for p in st.trainable_parameters():
            p.add_(1.0)
    path = ck.save_checkpoint(str(tmp_path / "c.pth"), st, None, epoch=1, config={"pruning": _pruning(), "optimizer": {"type": "SGD"}})
    torch.manual_seed(0)
    fresh = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})   # un-operated student: teacher-shaped
    epoch, opt, best = ck.resume(fresh, None, path, lambda ps: torch.optim.SGD(ps, lr=0.1))
    assert epoch == 1 and best is None and fresh.replaced_block_names == st.replaced_block_names
    assert len(opt.param_groups) == 1 and len(opt.param_groups[0]["params"]) == 4   # fresh optimizer of epoch 1
    for (k, a), (_, b) in zip(st.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k


def test_resume_rebuilds_param_groups_and_restores_the_optimizer_state(tmp_path):
    """A two-epoch plan: epoch 1 creates the optimizer, epoch 2 adds a param group with its own lr.  The reference
    replays prepare_train_epoch for every saved epoch so that optimizer.load_state_dict fits (layerwise_trainer.py:
    413-427); the RAdam moments must come back, and a changed optimizer type must leave the fresh state alone."""
    from kdcc.trainer import prepare_train_epoch
    pruning = _pruning(two_epochs=True)
    make = lambda ps: kdcc.optim.RAdam(ps, lr=0.1)
    torch.manual_seed(0)
    st = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    args = {"lr": 0.1}
    opt = None
    for ep in (1, 2):
        opt = prepare_train_epoch(st, pruning, ep, opt, make, args)
    assert [len(g["params"]) for g in opt.param_groups] == [2, 2] and opt.param_groups[1]["lr"] == 0.02
    for i, p in enumerate(st.trainable_parameters()):       # hand-made optimizer state: resume must bring it back
        opt.state[p] = {"step": 3, "exp_avg": torch.full_like(p, float(i + 1)), "exp_avg_sq": torch.full_like(p, 0.5)}
    path = ck.save_checkpoint(str(tmp_path / "c2.pth"), st, opt, epoch=2, monitor_best=0.25,
                              config={"pruning": pruning, "optimizer": {"type": "RAdam", "args": {"lr": 0.1}}})

    torch.manual_seed(0)
    fresh = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    epoch, ropt, best = ck.resume(fresh, None, path, make, optimizer_args={"lr": 0.1}, optimizer_type="RAdam")
    assert (epoch, best) == (2, 0.25) and fresh.replaced_block_names == ["body.0", "body.3"]
    assert [len(g["params"]) for g in ropt.param_groups] == [2, 2] and ropt.param_groups[1]["lr"] == 0.02
    for i, p in enumerate(fresh.trainable_parameters()):
        assert ropt.state[p]["step"] == 3 and float(ropt.state[p]["exp_avg"].mean()) == float(i + 1)

    torch.manual_seed(0)
    other = kdcc.DepthwiseStudent(TinyTeacher(), {"trainer": {"verbosity": 2}})
    _, sopt, _ = ck.resume(other, None, path, lambda ps: torch.optim.SGD(ps, lr=0.1), optimizer_args={"lr": 0.1}, optimizer_type="SGD")
    assert len(sopt.state) == 0                              # type changed: the saved moments are not loaded (:420-423)


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
