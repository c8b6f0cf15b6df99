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


def test_every_entry_point_refuses_null_handles(pb):
    """Error behaviour at the boundary: no entry point may crash or throw across the C ABI when it
    is handed a NULL engine / NULL buffers; status-returning ones answer with a negative PB_ERR_*
    code and leave a message in pb_last_error()."""
    import ctypes as C
    import re

    hdr = open(os.path.join(ROOT, "include", "pacbot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\bint\s+(pb_\w+)\s*\(([^;]*?)\)\s*;", hdr)
    assert len(protos) >= 40
    lib = C.CDLL(pb._lib.LIB_PATH)  # a private handle: argtypes set here must not leak
    lib.pb_last_error.restype = C.c_char_p
    checked = 0
    for name, args in protos:
        params = [a.strip() for a in args.split(",")]
        if params == ["void"] or not any("*" in a for a in params):
            continue  # pb_version and the pure table lookups take no pointers
        f = getattr(lib, name)
        f.restype = C.c_int
        rc = f(*[C.c_void_p(0) if "*" in a else C.c_int64(0) for a in params])
        if name in ("pb_destroy", "pb_mcts_destroy", "pb_mcts_max_nodes"):
            assert rc == 0, name  # destroying nothing is fine (free(NULL)); no trees -> 0 nodes
            continue
        assert rc < 0, (name, rc)
        assert lib.pb_last_error(), name
        checked += 1
    assert checked >= 35
