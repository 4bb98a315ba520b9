"""bench.py's contract on a CPU-only box: the reference arm runs (rank 0 only under torchrun-style env),
prints ONE JSON line with the agreed keys, and stays bounded whatever --steps says; the CUDA arm refuses to
run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=timeout)


def test_reference_arm_line_and_bound():
    t0 = time.time()
    r = run(["--impl", "reference", "--steps", "40", "--warmup", "3"])
    assert r.returncode == 0, r.stderr
    assert time.time() - t0 < 200
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Msamples/s IQ through decimate+FEC")
    assert d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 40 and d["warmup"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
