"""CPU differential campaign: the DEVICE game logic (pacbot-rl_b200/csrc/pb_engine.cuh, the source
k_step / k_reset execute per thread) compiled for the host through tests/host_engine/ and run
against the scalar C oracle on identical action tapes and identical injected draw tapes.

This is test infrastructure (see tests/host_engine/host_engine.cpp): it lets the container
without a GPU check every change to the engine source bit for bit, on far more env-steps than
the GPU budget allows.  The GPU parity tests (test_parity_gpu.py) remain the parity tests proper:
they exercise the compiled sm_100a code through the C ABI.  Integer work: tolerance none."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_states_equal

HE_DIR = os.path.join(ROOT, "tests", "host_engine")
HE_SO = os.path.join(HE_DIR, "_build", "libhost_engine.so")
CSRC = os.path.join(ROOT, "pacbot-rl_b200", "csrc")


def _build():
    srcs = [os.path.join(HE_DIR, f) for f in ("host_engine.cpp", "cuda_runtime.h")]
    srcs += [os.path.join(CSRC, f) for f in ("pb_engine.cuh", "pb_state.cuh", "pb_tables.cuh")]
    srcs.append(os.path.join(ROOT, "include", "pacbot_b200.h"))
    if os.path.exists(HE_SO) and all(os.path.getmtime(HE_SO) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(HE_SO), exist_ok=True)
    subprocess.run(
        ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", HE_DIR, "-I", CSRC,
         os.path.join(HE_DIR, "host_engine.cpp"), "-o", HE_SO],
        check=True, capture_output=True, text=True)


@pytest.fixture(scope="module")
def he(pb):
    _build()
    lib = C.CDLL(HE_SO)
    lib.he_sizeof_env_state.restype = C.c_size_t
    assert lib.he_sizeof_env_state() == pb.ENV_STATE_DTYPE.itemsize
    lib.he_step.restype = C.c_int64
    lib.he_step.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_uint64, C.c_int64, C.c_void_p, C.c_int64]
    lib.he_reset.restype = None
    lib.he_reset.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_uint64, C.c_int64,
                             C.c_void_p, C.c_int64]
    lib.he_action_mask.restype = None
    lib.he_action_mask.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    return lib


class HostEngineBatch:
    """n envs stepped by the host build of the device engine; state kept as pb_env_state records."""

    def __init__(self, lib, pb, n, cfg, tape=None, seed=0, gid_base=0):
        self.lib, self.n, self.seed, self.gid_base = lib, n, seed, gid_base
        self.tape = None if tape is None else np.ascontiguousarray(tape, np.uint32)
        self.st = np.zeros(n, pb.ENV_STATE_DTYPE)
        self.st["cfg_flags"] = np.asarray(cfg, np.uint8) & 0x0F
        self._reset(None, 0)  # pb_create: zeroed state + cfg flags, k_reset(randomise = 0)

    def _tape_args(self):
        if self.tape is None:
            return None, 0
        return self.tape.ctypes.data, self.tape.shape[1]

    def _reset(self, mask, randomise):
        t, tl = self._tape_args()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.lib.he_reset(self.st.ctypes.data, self.n, None if m is None else m.ctypes.data,
                          randomise, self.seed, self.gid_base, t, tl)

    def reset(self, mask=None):
        self._reset(mask, 1)

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, np.uint8)
        r = np.zeros(self.n, np.int32)
        d = np.zeros(self.n, np.uint8)
        m = np.zeros((self.n, 5), np.uint8)
        t, tl = self._tape_args()
        bad = self.lib.he_step(self.st.ctypes.data, self.n, a.ctypes.data, r.ctypes.data,
                               d.ctypes.data, m.ctypes.data, int(auto_reset), self.seed,
                               self.gid_base, t, tl)
        return r, d.astype(bool), m.astype(bool), bad

    def action_mask(self):
        m = np.zeros((self.n, 5), np.uint8)
        self.lib.he_action_mask(self.st.ctypes.data, self.n, m.ctypes.data)
        return m.astype(bool)


def random_actions(rng, mask, p_wild=0.1):
    n = mask.shape[0]
    a = (rng.random((n, 5)) * mask).argmax(1).astype(np.uint8)
    wild = rng.random(n) < p_wild
    a[wild] = rng.integers(0, 5, wild.sum(), dtype=np.uint8)
    return a


def run_campaign(he, pb, oracle, n, steps, seed, auto_reset, tape_len=509):
    rng = np.random.default_rng(seed)
    tape = rng.integers(0, 2**32, size=(n, tape_len), dtype=np.uint32)
    cfg = (np.arange(n) % 16).astype(np.uint8)
    ob = oracle.OracleBatch(n, cfg, tape)
    hb = HostEngineBatch(he, pb, n, cfg, tape)
    assert_states_equal(ob.export_state(), hb.st, "constructor")
    ob.reset()
    hb.reset()
    assert_states_equal(ob.export_state(), hb.st, "reset")
    episodes = 0
    for t in range(steps):
        mask = ob.action_mask()
        assert np.array_equal(mask, hb.action_mask()), f"mask before step {t}"
        a = random_actions(rng, mask)
        r_o, d_o = ob.step(a, auto_reset=auto_reset)
        r_h, d_h, m_h, bad = hb.step(a, auto_reset=auto_reset)
        assert bad == -1
        assert np.array_equal(r_o, r_h), f"reward differs at step {t}"
        assert np.array_equal(d_o.astype(bool), d_h), f"done differs at step {t}"
        assert np.array_equal(ob.action_mask(), m_h), f"mask differs at step {t}"
        assert_states_equal(ob.export_state(), hb.st, f"step {t}")
        episodes += int(d_h.sum())
    return episodes


@pytest.mark.parametrize("auto_reset", [True, False])
def test_device_engine_source_matches_oracle_on_cpu(he, pb, oracle, auto_reset):
    """~0.5 M env-steps per case (all 16 config-flag combinations, deaths, frightened phases,
    auto resets / stepping done envs), every field of every env compared after every step."""
    episodes = run_campaign(he, pb, oracle, n=2048, steps=260, seed=11 + auto_reset,
                            auto_reset=auto_reset)
    assert episodes > 100, episodes


# PB_EXTRA_FUZZ_SEEDS="3,4,5" widens both campaigns for a one-off run (same switch as the GPU suite)
EXTRA = [int(x) for x in os.environ.get("PB_EXTRA_FUZZ_SEEDS", "").split(",") if x]


@pytest.mark.parametrize("seed", [0, 1, 2] + EXTRA)
def test_device_engine_source_fuzzed_states(he, pb, oracle, seed):
    """The fuzz generator of the GPU suite (test_parity_gpu.fuzz_states: random well-formed states
    next to every rare branch -- anger thresholds, fruit spawn / expiry, last pellet, mode flip,
    level-time penalty, fright expiry, ghost-eating combos, every ticks_per_step and update
    period), imported on both sides and stepped; everything compared after every step."""
    from test_parity_gpu import fuzz_states, gpu_to_oracle_state

    n, steps = 2048, 12
    rng = np.random.default_rng(100 + seed)
    tape = rng.integers(0, 2**32, size=(n, 211), dtype=np.uint32)
    st = fuzz_states(rng, pb, n)
    ob = oracle.OracleBatch(n, st["cfg_flags"], tape)
    hb = HostEngineBatch(he, pb, n, st["cfg_flags"], tape)
    ob.import_state(gpu_to_oracle_state(oracle, st))
    hb.st[:] = st
    assert_states_equal(ob.export_state(), hb.st, "import")
    seen_done = 0
    for t in range(steps):
        a = random_actions(rng, ob.action_mask(), p_wild=0.2)
        r_o, d_o = ob.step(a, auto_reset=False)
        r_h, d_h, m_h, _ = hb.step(a, auto_reset=False)
        assert np.array_equal(r_o, r_h), f"reward differs at step {t}"
        assert np.array_equal(d_o.astype(bool), d_h), f"done differs at step {t}"
        assert np.array_equal(ob.action_mask(), m_h), f"mask differs at step {t}"
        assert_states_equal(ob.export_state(), hb.st, f"fuzz step {t}")
        seen_done += int(d_h.sum())
    final = hb.st
    assert seen_done > 0 and (final["curr_level"] > 1).any() and (final["curr_lives"] < 3).any()
    assert (final["update_period"] < 12).any()


@pytest.mark.parametrize("seed", EXTRA)
def test_device_engine_source_extra_campaign(he, pb, oracle, seed):
    run_campaign(he, pb, oracle, n=2048, steps=400, seed=1000 + seed, auto_reset=bool(seed & 1))


def test_invalid_action_leaves_env_untouched(he, pb, oracle):
    n = 64
    cfg = np.zeros(n, np.uint8)
    hb = HostEngineBatch(he, pb, n, cfg, None, seed=3)
    before = hb.st.copy()
    a = np.zeros(n, np.uint8)
    a[17] = 5
    a[40] = 200
    r, d, m, bad = hb.step(a)
    assert bad == 17
    for i in (17, 40):
        assert hb.st[i] == before[i] and r[i] == 0 and not d[i]


# ---- the GPU suite's own scenario tests, re-run with the host build in place of the GPU ----------
class _HostEnv:
    """The few methods test_parity_gpu's helpers call on a BatchedPacmanGym."""

    def __init__(self, hb, auto_reset):
        self.hb, self.auto_reset = hb, auto_reset

    def reset(self):
        self.hb.reset()

    def set_state(self, st):
        self.hb.st[:] = st


def _patch_parity_helpers(monkeypatch, he, pb):
    import test_parity_gpu as tp

    def make_pair(pb_, oracle_, n, cfg, tape, auto_reset=False, seed=0):
        cfg = np.asarray(cfg, np.uint8)
        return (oracle_.OracleBatch(n, cfg, tape),
                _HostEnv(HostEngineBatch(he, pb, n, cfg, tape, seed=seed), auto_reset))

    def compare_step(ob, env, actions, t, auto_reset, check_obs):
        r_o, d_o = ob.step(actions, auto_reset=auto_reset)
        r_h, d_h, m_h, _ = env.hb.step(actions, auto_reset=auto_reset)
        assert np.array_equal(r_o, r_h), f"reward differs at step {t}"
        assert np.array_equal(d_o.astype(bool), d_h), f"done differs at step {t}"
        assert np.array_equal(ob.action_mask(), m_h), f"mask differs at step {t}"
        assert_states_equal(ob.export_state(), env.hb.st, f"step {t}")
        return d_o

    monkeypatch.setattr(tp, "make_pair", make_pair)
    monkeypatch.setattr(tp, "compare_step", compare_step)
    return tp


def test_scripted_scenarios_on_the_host_build(he, pb, oracle, monkeypatch):
    """test_parity_gpu.test_scripted_scenarios (deaths both ways, 1-4 ghost
    combos, fright expiry, mode flips, level-time penalty, super pellet, fruit, anger, last pellet,
    every ticks_per_step) with its own scenario list and guards, the host build standing in for
    the GPU engine (observations are not compared here: the encoder is GPU-only)."""
    tp = _patch_parity_helpers(monkeypatch, he, pb)
    tp.test_scripted_scenarios(pb, oracle)


@pytest.mark.parametrize("n", [1, 31, 33])
def test_odd_batch_sizes_on_the_host_build(he, pb, oracle, monkeypatch, n):
    tp = _patch_parity_helpers(monkeypatch, he, pb)
    tp.test_odd_batch_sizes(pb, oracle, n)


def test_host_build_reproduces_golden_trajectory(he, pb):
    """The committed trajectory fixture (tests/golden/trajectory_v1.npz), replayed by the host build
    of the device engine: reward, done, mask and the digest of the full state at every step (the
    observation digest needs the encoder, i.e. the GPU test of the same fixture)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_trajectory import state_digest

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "trajectory_v1.npz")))
    hb = HostEngineBatch(he, pb, len(g["cfg"]), g["cfg"], g["tape"])
    hb.reset()
    for t in range(g["actions"].shape[0]):
        assert np.array_equal(hb.action_mask(), g["masks"][t]), f"mask differs at step {t}"
        r, d, _, _ = hb.step(g["actions"][t], auto_reset=True)
        assert np.array_equal(r, g["reward"][t]), f"reward differs at step {t}"
        assert np.array_equal(d, g["done"][t]), f"done differs at step {t}"
        assert state_digest(hb.st) == g["state_digest"][t], f"state differs at step {t}"
