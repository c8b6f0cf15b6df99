"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
(`--impl reference`, the CPU implementation of the path on the host cores) prints exactly ONE JSON
line with the agreed keys, alone and under a 2-rank torchrun launch (rank 0 prints, rank 1 exits 0
without work), and the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
            "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _one_json_line(stdout: str) -> dict:
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, stdout
    return json.loads(lines[0])


def _check_reference_line(d: dict, n_gpus: int) -> None:
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec_incl_obs"
    assert d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == n_gpus and d["steps"] == 3 and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    _check_reference_line(_one_json_line(res.stdout), 1)


def test_reference_arm_under_torchrun_rank0_only():
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1"],
        capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    _check_reference_line(_one_json_line(res.stdout), 2)


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode != 0          # loud failure, no CPU fallback, no JSON line
    assert not [ln for ln in res.stdout.splitlines() if ln.strip().startswith("{")]
