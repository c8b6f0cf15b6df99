"""CPU differential test of the observation encoder: pacbot-rl_b200/csrc/pb_encode.cuh
(`k_encode_obs`, PacmanGym::obs, env.rs:410-500) compiled for the host and executed by one OS thread
per lane (tests/host_engine/host_encode.cpp) against the oracle's observation, bit for bit
(float32 compared as uint32).  Each warp encodes hundreds of envs one after the other, so the
un-patching of the previous env's cells, the ticket hand-out and every lane role are exercised on
the GPU suite's fuzzed states (frightened ghosts stacked on one cell, last pellet, fruit, empty
locations, ...).  Test infrastructure only; the GPU tests of the encoder (TMA stores, ring slots,
65 536- and 2 M-env launches) remain the parity tests proper."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_obs_equal
from test_parity_gpu import fuzz_states, gpu_to_oracle_state

HE_DIR = os.path.join(ROOT, "tests", "host_engine")
SO = os.path.join(HE_DIR, "_build", "libhost_encode.so")
CSRC = os.path.join(ROOT, "pacbot-rl_b200", "csrc")


@pytest.fixture(scope="module")
def henc(pb):
    srcs = [os.path.join(HE_DIR, f) for f in ("host_encode.cpp", "host_common.h", "cuda_runtime.h")]
    srcs += [os.path.join(CSRC, f) for f in ("pb_encode.cuh", "pb_engine.cuh", "pb_state.cuh", "pb_tables.cuh")]
    if not (os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(
            ["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
             "-pthread", "-shared", "-fPIC", "-I", HE_DIR, "-I", CSRC,
             os.path.join(HE_DIR, "host_encode.cpp"), "-o", SO], check=True, capture_output=True, text=True)
    lib = C.CDLL(SO)
    lib.he_encode_obs.restype = C.c_int
    lib.he_encode_obs.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]

    def encode(states, first=0, count=None):
        st = np.ascontiguousarray(states, pb.ENV_STATE_DTYPE)
        count = len(st) - first if count is None else count
        out = np.full((count, 17, 28, 31), np.nan, np.float32)    # every float must be written
        assert lib.he_encode_obs(st.ctypes.data, len(st), first, count, out.ctypes.data) == 0
        return out

    return encode


def oracle_obs(oracle, st):
    ob = oracle.OracleBatch(len(st), np.zeros(len(st), np.uint8), None)
    ob.import_state(gpu_to_oracle_state(oracle, st))
    return ob.obs(), ob.export_state()


@pytest.mark.parametrize("seed", [0, 1])
def test_encoder_source_matches_oracle_on_fuzzed_states(henc, pb, oracle, seed):
    rng = np.random.default_rng(300 + seed)
    st = fuzz_states(rng, pb, 2048)
    # more of what the single-cell lanes fight over: several frightened ghosts on one cell, on a
    # super-pellet cell, on the fruit cell; Pac-Man and ghosts off the board
    k = np.arange(len(st))
    stack = k % 7 == 0
    st["g_row"][stack, 1:3] = st["g_row"][stack, :1]
    st["g_col"][stack, 1:3] = st["g_col"][stack, :1]
    st["g_fright"][stack, :3] = rng.integers(1, 41, (int(stack.sum()), 3))
    sup = k % 11 == 0
    st["g_row"][sup, 3], st["g_col"][sup, 3] = 3, 26
    st["pellets"][sup, 3] |= np.uint32(1 << 26)
    fruit = k % 13 == 0
    st["g_row"][fruit, 0], st["g_col"][fruit, 0], st["fruit_steps"][fruit] = 17, 13, 9
    off = k % 17 == 0
    st["pac_row"][off], st["pac_col"][off] = 32, 32
    want, ost = oracle_obs(oracle, st)
    assert_obs_equal(want, henc(st), f"fuzz seed {seed}", ost)


def test_encoder_source_on_a_rollout_and_ranges(henc, pb, oracle):
    """States a game actually visits (oracle rollout with deaths, frightened phases, resets),
    ragged env counts and a sub-range launch (pb_encode_obs_range)."""
    n = 333
    rng = np.random.default_rng(7)
    tape = rng.integers(0, 2**32, size=(n, 509), dtype=np.uint32)
    ob = oracle.OracleBatch(n, np.full(n, 3, np.uint8), tape)
    ob.reset()
    for t in range(120):
        m = ob.action_mask()
        a = (rng.random((n, 5)) * m).argmax(1).astype(np.uint8)
        ob.step(a, auto_reset=True)
        if t % 40 == 39:
            flat = ob.export_state()
            st = np.zeros(n, pb.ENV_STATE_DTYPE)
            for f in st.dtype.names:
                if f in flat.dtype.names:
                    st[f] = flat[f]
            assert_obs_equal(ob.obs(), henc(st), f"rollout step {t}", flat)
    for cnt in (1, 31, 33):
        assert_obs_equal(ob.obs()[:cnt], henc(st[:cnt]), f"{cnt} envs")
    assert_obs_equal(ob.obs()[100:177], henc(st, first=100, count=77), "sub-range")


def test_host_builds_reproduce_golden_trajectory_including_observations(henc, pb):
    """tests/golden/trajectory_v1.npz end to end on the CPU with product sources only: the host
    build of the engine steps the envs, the host-emulated encoder writes their observations; the
    fixture's per-step digests of the full state AND of the float32 observations must come out
    (the GPU test of the same fixture does this through the C ABI)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_trajectory import obs_digest, state_digest
    from test_host_engine import HE_SO, HostEngineBatch, _build

    _build()
    lib = C.CDLL(HE_SO)
    lib.he_step.restype = C.c_int64
    lib.he_step.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_uint64, C.c_int64, C.c_void_p, C.c_int64]
    lib.he_reset.restype = None
    lib.he_reset.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_uint64, C.c_int64,
                             C.c_void_p, C.c_int64]
    lib.he_action_mask.restype = None
    lib.he_action_mask.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "trajectory_v1.npz")))
    hb = HostEngineBatch(lib, pb, len(g["cfg"]), g["cfg"], g["tape"])
    hb.reset()
    for t in range(g["actions"].shape[0]):
        r, d, _, _ = hb.step(g["actions"][t], auto_reset=True)
        assert np.array_equal(r, g["reward"][t]) and np.array_equal(d, g["done"][t]), t
        assert state_digest(hb.st) == g["state_digest"][t], f"state differs at step {t}"
        assert obs_digest(henc(hb.st)) == g["obs_digest"][t], f"observation differs at step {t}"
