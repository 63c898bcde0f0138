"""bench.py's contract where it can be exercised without a GPU: the reference arm (`--impl reference`, the CPU restatement on the host cores)
prints ONE JSON line with the keys the driver reads and the same `config` object the GPU arm reports; the GPU arm refuses to run without CUDA
(no CPU fallback for the product path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--envs-per-gpu", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the config object is built by the same Spec for both arms (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    cfg = d["config"]
    assert cfg["envs_per_gpu"] == 256 and cfg["substeps"] == 4 and cfg["solver"] == "cg" and "workload" in cfg
    a = argparse.Namespace(workload="c2", substeps=4, envs_per_gpu=256, ctrl_dt=0.02, solver="cg", solver_iterations=8, ls_iterations=8, no_obs=False, seed=0,
                           flush="write")
    assert bench.Spec(a).config(1) == cfg


def test_gpu_arm_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-extras")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
