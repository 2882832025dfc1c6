"""bench.py contract, the part that runs without a GPU: the reference arm prints ONE JSON line with the agreed keys, and the
fused arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-points", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train-step collocation pts/sec" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    # "reference" when baseline/_ref (the unmodified reference, scripts/install_reference.py) is present, else the oracle port
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "deepxde"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-points", "64", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fused_arm_has_no_cpu_fallback():
    r = _run("--steps", "1", "--warmup", "1", "--points", "64", "--no-cpu-baseline")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in (r.stderr + r.stdout)
