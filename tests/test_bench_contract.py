"""The reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) runs
without a GPU; its JSON line must carry the keys the measurement contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


@pytest.mark.parametrize("workload", ["c1", "c2", "sparse"])
def test_reference_arm_line(workload):
    # c2's own sample is n = m = 1000 (minutes of CPU work): the contract is checked on a smaller one
    d = run_reference("--workload", workload, *(["--cpu-size", "160"] if workload == "c2" else []))
    assert d["impl"] == "reference" and d["metric"] == "newton_step_ms" and d["unit"] == "ms"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] == d["value"] and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if workload in ("c1", "c2"):
        assert cb["blas3_gram"]["value"] > 0           # the BLAS-3 variant of the port beside the as-written one
    if workload == "c1":
        assert "extrapolat" not in cb["sample"]        # config 1 is measured at its own size
    if workload == "c2":
        # the work model is validated on a second sample size, and the bounded sample is declared
        assert cb["model_check"]["predicted_ms_per_step"] > 0 and cb["model_check"]["measured_ms_per_step"] > 0
        assert d["sample_steps"] == {"timed": 2, "warmup": 1} and "extrapolat" in d["config"]["measured_on"]


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
