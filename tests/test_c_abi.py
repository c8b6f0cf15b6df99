"""The C ABI from plain C: compile tests/c_abi_smoke.c against include/pacbot_b200.h with gcc,
link the in-tree shared library, run it.  On the CPU box it checks the no-device refusal; on a
GPU box the same binary does a reset / step / obs round trip."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, pb):
    lib = pb._lib.LIB_PATH
    exe = str(tmp_path / "c_abi_smoke")
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    subprocess.run(
        [cc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
         os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, lib,
         "-Wl,-rpath," + os.path.dirname(lib)],
        check=True, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_c_abi_from_plain_c_without_gpu(tmp_path, pb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked variant")
    res = _build_and_run(tmp_path, pb)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "refused as designed" in res.stdout


@pytest.mark.gpu
def test_c_abi_from_plain_c_on_gpu(tmp_path, pb):
    res = _build_and_run(tmp_path, pb)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "ok on a GPU" in res.stdout
