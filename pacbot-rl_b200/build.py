"""Build the sm_100a CUDA extension in-tree with nvcc (no JIT cache: the .so must travel with the
repo snapshot to the GPU box).  Usage: python pacbot-rl_b200/build.py [--force] [--verbose]"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpacbot_b200.so")
SOURCES = [os.path.join(CSRC, "pb_kernels.cu")]
HEADERS = [
    os.path.join(CSRC, "pb_tables.cuh"),
    os.path.join(CSRC, "pb_state.cuh"),
    os.path.join(CSRC, "pb_engine.cuh"),
    os.path.join(CSRC, "pb_encode.cuh"),
    os.path.join(CSRC, "pb_encode_small.cuh"),
    os.path.join(CSRC, "pb_mcts.cuh"),
    os.path.join(CSRC, "pb_mcts_abi.inl"),
    os.path.join(CSRC, "pb_rollout_abi.inl"),
    os.path.join(CSRC, "pb_replay_abi.inl"),
    os.path.join(CSRC, "pb_qnet_abi.inl"),
    os.path.join(CSRC, "pb_dqn_abi.inl"),
    os.path.join(os.path.dirname(HERE), "include", "pacbot_b200.h"),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return cand


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [__file__])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # PB_NVCC_EXTRA: extra flags for A/B experiments (e.g. "-DPB_ENC_PREFETCH_DIST=512")
    extra = os.environ.get("PB_NVCC_EXTRA", "").split()
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", LIB, *SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
