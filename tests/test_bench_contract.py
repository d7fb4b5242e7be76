"""bench.py's reference arm (the reference's NumPy/BLAS path timed on the host cores, oracle port) runs without a GPU
and prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000", "--dim", "64",
                          "--queries", "64", "--cpu-queries", "16", "--k", "10", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "exact top-k QPS" and d["unit"] == "queries/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["rows_total"] == 20000 and d["config"]["k"] == 10


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "1000",
                          "--dim", "16", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line_on_a_small_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--rows", "30000", "--dim", "64", "--queries", "300",
                          "--k", "10", "--steps", "2", "--warmup", "3", "--no-regimes", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 300 * 64 * 4
    assert d["e2e"]["d2h_bytes_per_step"] == 300 * 10 * (4 + 8)
    assert d["roofline"]["bound"] in ("tensor", "hbm") and 0 < d["roofline"]["frac"]
    assert d["config"]["exact_fallback_fraction"] <= 0.25
