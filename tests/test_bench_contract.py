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


def test_smem_floor_model():
    """roofline.smem_floor: DESIGN.md 4.0's shared-memory byte model applied to the default workload."""
    sys.path.insert(0, ROOT)
    import bench
    plan = bench.plan_51m()
    need_dx = [False] + [True] * (len(plan) - 1)
    planes = 4 * sum(ci for ci, _ in plan)
    f = bench.smem_floor(plan, 4, (9, 5, 20), "nchw", 128, need_dx, 2.66, 1900.0)
    want = ((planes - 4 * plan[0][0]) * 618 * 1024 + planes * 378 * 1024) / 128 / 148 / 1.9e6
    assert abs(f["ms"] - want) < 1e-3 and 0 < f["frac"] < 1
    assert bench.smem_floor(plan, 4, (9, 5, 20), "nhwc", 128, need_dx, 3.15, 1900.0) is None
    assert bench.smem_floor(plan, 4, (9, 5, 20), "nchw", 128, need_dx, 3.15, None) is None
    assert bench.smem_floor(plan, 4, (3, 1, 1), "nchw", 128, need_dx, 3.15, 1900.0) is None
