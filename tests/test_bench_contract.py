"""bench.py's reference arm runs without a GPU (it times the torch-CPU port of the path), so its JSON line -- the
contract the driver parses -- is checked here; the kdcc arm needs a B200 and must refuse to run without one."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--crop", "256"], 600)   # small crop: seconds, same line
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "img/s"
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_kdcc_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--e2e-steps", "0"], 300)
    assert r.returncode != 0
    assert "cuda" in (r.stderr + r.stdout).lower()


def test_tensor_issue_floor_model():
    """roofline.tensor_issue_floor: DESIGN.md 4.0's MMA cost model applied to the default workload."""
    sys.path.insert(0, ROOT)
    import bench
    plan = bench.plan_51m()
    planes = 4 * sum(ci for ci, _ in plan)
    f = bench.tensor_issue_floor(plan, 4, (9, 5, 20), "nchw", 128, 3.25, 1900.0)
    assert abs(f["ms"] - planes * (90 * 59 + 72 * 74) / 148 / 1.9e6) < 1e-3 and 0 < f["frac"] < 1
    assert bench.tensor_issue_floor(plan, 4, (9, 5, 20), "nhwc", 128, 3.25, 1900.0) is None
    assert bench.tensor_issue_floor(plan, 4, (9, 5, 20), "nchw", 128, 3.25, None) is None
