"""CPU differential tests of the DEVICE MCTS kernels (pacbot-rl_b200/csrc/pb_mcts.cuh) compiled for
the host (tests/host_engine/host_mcts.cpp) against the scalar MCTS oracle (oracle/mcts_oracle.c, a
restatement of pacbot_rs/src/mcts.rs): the GPU suite's own test bodies
(test_mcts_gpu.test_search_take_action_and_reset_match_oracle, ..ponder..) are re-run with a host
stand-in for `BatchedPacmanGym` / `BatchedMCTS`, so tree statistics, rewards and env states are
compared bit for bit over search / re-root / fresh-root / compaction / reset cycles without a GPU.

Test infrastructure only (like tests/test_host_engine.py): `HostMCTS` below re-implements the
host-side orchestration of pacbot-rl_b200/mcts.py with NumPy, the leaf observations come from the
env oracle (the encoder is GPU-only), and the GPU tests remain the parity tests proper."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from test_parity_gpu import gpu_to_oracle_state

HE_DIR = os.path.join(ROOT, "tests", "host_engine")
HM_SO = os.path.join(HE_DIR, "_build", "libhost_mcts.so")
CSRC = os.path.join(ROOT, "pacbot-rl_b200", "csrc")
Q = dict(best_action=0, action_distribution=1, policy_prior=2, action_values=3, value=4, max_depth=5,
         node_count=6, root_visits=7)                       # MCTS_Q_* of pb_mcts.cuh
ST_REROOTED, ST_NEEDS_ROOT, ST_WAS_DONE, ST_BAD_ACTION = 2, 3, 4, 5
ERR = {1: "node pool full", 2: "tried to expand the same child twice", 4: "search from a finished env"}


def _build():
    srcs = [os.path.join(HE_DIR, f) for f in ("host_mcts.cpp", "host_common.h", "cuda_runtime.h")]
    srcs += [os.path.join(CSRC, f) for f in ("pb_mcts.cuh", "pb_engine.cuh", "pb_state.cuh", "pb_tables.cuh")]
    srcs.append(os.path.join(ROOT, "include", "pacbot_b200.h"))
    if os.path.exists(HM_SO) and all(os.path.getmtime(HM_SO) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(HM_SO), exist_ok=True)
    # -ffp-contract=off / no fast-math: one IEEE f32 operation per __f*_rn intrinsic, as on the GPU
    subprocess.run(
        ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
         "-shared", "-fPIC", "-I", HE_DIR, "-I", CSRC, os.path.join(HE_DIR, "host_mcts.cpp"), "-o", HM_SO],
        check=True, capture_output=True, text=True)


@pytest.fixture(scope="module")
def hm(pb):
    _build()
    L = C.CDLL(HM_SO)
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.hm_engine_create.restype = vp
    L.hm_engine_create.argtypes = [i64, vp, u64, i64]
    L.hm_mcts_create.restype = vp
    L.hm_mcts_create.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.hm_mcts_check.restype = C.c_uint
    L.hm_mcts_check.argtypes = [vp, vp]
    L.hm_mcts_collect_act.restype = C.c_int
    L.hm_mcts_collect_act.argtypes = [vp, vp, C.c_int, C.c_int] + [vp] * 7
    L.hm_engine_copy_envs.restype = C.c_int
    L.hm_engine_copy_envs.argtypes = [vp, vp, vp, vp, i64]
    for name, args in dict(
            hm_engine_destroy=[vp], hm_engine_reset=[vp, vp], hm_engine_get_state=[vp, vp],
            hm_engine_set_state=[vp, vp], hm_engine_action_mask=[vp, vp], hm_mcts_destroy=[vp],
            hm_mcts_reset_root=[vp, vp, vp, vp], hm_mcts_root_noise=[vp, vp, vp, C.c_float],
            hm_mcts_select=[vp, vp, vp, vp], hm_mcts_backprop=[vp, vp, vp],
            hm_mcts_take_action=[vp, vp, vp, vp, vp, vp], hm_mcts_query=[vp, C.c_int, vp]).items():
        getattr(L, name).restype = None
        getattr(L, name).argtypes = args
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def _mask_np(mask, n):
    if mask is None:
        return None
    m = np.ascontiguousarray(np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask)).astype(np.uint8)
    assert m.shape == (n,)
    return m


class HostEnv:
    """The part of BatchedPacmanGym the MCTS tests use, on the host build of the engine."""

    device = torch.device("cpu")
    oracle = None   # set by the fixtures: observations come from the env oracle

    def __init__(self, L, pb, n, random_start, random_ticks, *, device=None, seed=0, env_id_base=0):
        self.L, self.pb, self.num_envs, self.seed, self.env_id_base = L, pb, n, seed, env_id_base
        rs = np.broadcast_to(np.asarray(random_start, bool), (n,))
        rt = np.broadcast_to(np.asarray(random_ticks, bool), (n,))
        self.cfg = (rs.astype(np.uint8) | (rt.astype(np.uint8) << 1))
        self.h = L.hm_engine_create(n, _ptr(np.ascontiguousarray(self.cfg)), seed, env_id_base)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hm_engine_destroy(self.h)
            self.h = None

    def reset(self, mask=None):
        m = _mask_np(mask, self.num_envs)
        self.L.hm_engine_reset(self.h, _ptr(m))

    def get_state(self):
        st = np.zeros(self.num_envs, self.pb.ENV_STATE_DTYPE)
        self.L.hm_engine_get_state(self.h, st.ctypes.data)
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, self.pb.ENV_STATE_DTYPE)
        self.L.hm_engine_set_state(self.h, st.ctypes.data)

    def action_mask(self):
        m = np.zeros((self.num_envs, 5), np.uint8)
        self.L.hm_engine_action_mask(self.h, m.ctypes.data)
        return torch.from_numpy(m.astype(bool))

    @property
    def _h(self):          # the name the product's host code uses for the engine handle
        return self.h

    def obs_range(self, first, count, out):
        """What k_encode_obs writes for envs [first, first + count): the oracle's observation of
        the same states (the encoder itself is GPU-only and has its own parity tests)."""
        st = gpu_to_oracle_state(self.oracle, self.get_state()[first:first + count])
        ob = self.oracle.OracleBatch(count, np.zeros(count, np.uint8), None)
        ob.import_state(st)
        out.copy_(torch.from_numpy(ob.obs()))
        return out

    def obs(self, out=None):
        if out is None:
            out = torch.empty((self.num_envs, 17, 28, 31), dtype=torch.float32)
        return self.obs_range(0, self.num_envs, out)

    def copy_envs_from(self, src, src_index=None, dst_index=None, count=None):
        def ptr(t):
            return None if t is None else np.ascontiguousarray(t.cpu().numpy(), np.int64)

        si, di = ptr(src_index), ptr(dst_index)
        if count is None:
            count = len(di) if di is not None else len(si) if si is not None else min(self.num_envs, src.num_envs)
        assert self.L.hm_engine_copy_envs(self.h, _ptr(di), src.h, _ptr(si), int(count)) == 0


class _CollectActABI:
    """`mcts._L` as pacbot-rl_b200/alphazero.py uses it: pb_mcts_collect_act with the C ABI's
    argument list, on host pointers."""

    def __init__(self, L):
        self._lib = L

    def pb_mcts_collect_act(self, mcts_h, snap_h, max_len, tree_size, remaining, ep_len, ep_mask, ep_dist,
                            ep_value, status, flags, stream):
        return self._lib.hm_mcts_collect_act(mcts_h, snap_h, max_len, tree_size, remaining, ep_len, ep_mask,
                                             ep_dist, ep_value, status, flags)


class HostMCTS:
    """pacbot-rl_b200/mcts.py::BatchedMCTS re-stated over the host build of the kernels."""

    def __init__(self, L, pb, oracle, env, evaluator, max_nodes, *, noise_fn=None, noise_mix=0.25):
        self.L, self.pb, self.oracle, self.env, self.evaluator = L, pb, oracle, env, evaluator
        self.max_nodes, self.noise_fn, self.noise_mix = int(max_nodes), noise_fn, float(noise_mix)
        n = self.num_envs = env.num_envs
        self.sim = HostEnv(L, pb, n, False, False, seed=env.seed, env_id_base=env.env_id_base)
        factor = int(os.environ.get("PB_MCTS_POOL_FACTOR", "8"))
        tpw = int(os.environ.get("PB_MCTS_TREES_PER_WARP", "1"))
        self.h = L.hm_mcts_create(env.h, self.sim.h, self.max_nodes, factor, tpw)
        self.device = env.device
        self._h = self.h
        self._L = _CollectActABI(L)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hm_mcts_destroy(self.h)
            self.h = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _stream(self):
        return None

    def _obs_of(self, engine):
        return engine.obs()

    def _evaluate(self, obs, mask):
        value, policy = self.evaluator(obs, mask)
        return (np.ascontiguousarray(value.cpu().numpy(), np.float32).reshape(-1),
                np.ascontiguousarray(policy.cpu().numpy(), np.float32).reshape(-1, 5))

    def check(self):
        epoch = C.c_uint64(0)
        err = self.L.hm_mcts_check(self.h, C.addressof(epoch))
        if err:
            raise ValueError("pb_mcts: " + "; ".join(v for k, v in ERR.items() if err & k))
        return int(epoch.value)

    def add_root_noise(self, mask=None):
        m = _mask_np(mask, self.num_envs)
        num_valid = self.env.action_mask().sum(dim=1)
        if self.noise_fn is not None:
            noise = self.noise_fn(num_valid, None if m is None else torch.from_numpy(m))
        else:   # the product's sampler (a torch function; runs on CPU tensors as well)
            noise = self.pb.mcts.dirichlet_root_noise(num_valid)
        noise = np.ascontiguousarray(noise.cpu().numpy(), np.float32)
        self.L.hm_mcts_root_noise(self.h, _ptr(m), noise.ctypes.data, self.noise_mix)

    def reset_root(self, mask=None, index=None):
        m = _mask_np(mask, self.num_envs)
        value, policy = self._evaluate(self._obs_of(self.env), self.env.action_mask())
        self.L.hm_mcts_reset_root(self.h, _ptr(m), value.ctypes.data, policy.ctypes.data)
        self.add_root_noise(m)

    def reset(self, mask=None, index=None):
        self.env.reset(mask)
        self.reset_root(mask)

    def search_iteration(self, active=None):
        n = self.num_envs
        a = _mask_np(active, n)
        kind = np.zeros(n, np.uint8)
        lmask = np.zeros((n, 5), np.uint8)
        self.L.hm_mcts_select(self.h, _ptr(a), kind.ctypes.data, lmask.ctypes.data)
        value, policy = self._evaluate(self._obs_of(self.sim), torch.from_numpy(lmask.astype(bool)))
        self.L.hm_mcts_backprop(self.h, value.ctypes.data, policy.ctypes.data)

    def ponder(self, max_tree_size):
        if max_tree_size > self.max_nodes:
            raise ValueError(f"max_tree_size {max_tree_size} exceeds max_nodes {self.max_nodes}")
        remaining = (max_tree_size - self.node_count()).clamp(min=0)
        for it in range(int(remaining.max().item())):
            self.search_iteration(remaining > it)
        return self.best_action()

    def take_action(self, actions, mask=None):
        n = self.num_envs
        m = _mask_np(mask, n)
        a = np.ascontiguousarray(actions.cpu().numpy(), np.uint8)
        reward, done, status = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.L.hm_mcts_take_action(self.h, _ptr(m), a.ctypes.data, reward.ctypes.data, done.ctypes.data,
                                   status.ctypes.data)
        if (status == ST_WAS_DONE).any():
            raise RuntimeError("take_action on a finished env")
        if (status == ST_BAD_ACTION).any():
            raise ValueError("Invalid action")
        if (status == ST_NEEDS_ROOT).any():
            self.reset_root(status == ST_NEEDS_ROOT)
        if (status == ST_REROOTED).any():
            self.add_root_noise(status == ST_REROOTED)
        return torch.from_numpy(reward), torch.from_numpy(done.astype(bool)), torch.from_numpy(status)

    def _query(self, what, shape, dtype):
        out = np.zeros(shape, dtype)
        self.L.hm_mcts_query(self.h, Q[what], out.ctypes.data)
        return torch.from_numpy(out)


for _name, _shape, _dtype in [("best_action", 1, np.int32), ("action_distribution", 5, np.float32),
                              ("policy_prior", 5, np.float32), ("action_values", 5, np.float32),
                              ("value", 1, np.float32), ("max_depth", 1, np.int32),
                              ("node_count", 1, np.int32), ("root_visits", 1, np.int32)]:
    def _make(name=_name, width=_shape, dtype=_dtype):
        def query(self):
            return self._query(name, (self.num_envs,) if width == 1 else (self.num_envs, width), dtype)
        return query
    setattr(HostMCTS, _name, _make())


@pytest.fixture
def host_pb(hm, pb, oracle, monkeypatch):
    """`pb` with BatchedPacmanGym / mcts.BatchedMCTS replaced by the host stand-ins, and
    Tensor.cuda() as the identity (the re-used GPU test bodies call it on their masks)."""
    monkeypatch.setattr(pb, "BatchedPacmanGym",
                        lambda n, rs=False, rt=False, **kw: HostEnv(hm, pb, n, rs, rt, **kw))
    monkeypatch.setattr(pb.mcts, "BatchedMCTS",
                        lambda env, ev, max_nodes, **kw: HostMCTS(hm, pb, oracle, env, ev, max_nodes, **kw))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(HostEnv, "oracle", oracle)

    # pacbot-rl_b200/alphazero.py resolves these two names in its own namespace; with the stand-ins
    # in place the product's BatchedExperienceCollector (its host-side episode bookkeeping
    # included) runs UNMODIFIED over the host build of the kernels
    class MCTSFactory:
        _p = staticmethod(HostMCTS._p)

        def __new__(cls, env, ev, max_nodes, **kw):
            return HostMCTS(hm, pb, oracle, env, ev, max_nodes, **kw)

    monkeypatch.setattr(pb.alphazero, "BatchedPacmanGym", pb.BatchedPacmanGym)
    monkeypatch.setattr(pb.alphazero, "BatchedMCTS", MCTSFactory)
    return pb


@pytest.fixture(scope="module")
def pymcts(oracle):
    from oracle import pymcts

    pymcts.lib()
    return pymcts


@pytest.mark.parametrize("pool_factor", ["8", "1"], ids=["pool8x", "pool1x"])
@pytest.mark.parametrize("evaluator", ["hashed", "uniform"])
def test_device_mcts_source_matches_oracle_on_cpu(host_pb, pymcts, evaluator, pool_factor, monkeypatch):
    import test_mcts_gpu as tg
    from mcts_common import np_evaluator, uniform_evaluator

    tg.test_search_take_action_and_reset_match_oracle(
        host_pb, pymcts, np_evaluator if evaluator == "hashed" else uniform_evaluator, pool_factor, monkeypatch)


def test_device_mcts_ponder_on_cpu(host_pb, pymcts):
    import test_mcts_gpu as tg

    tg.test_ponder_grows_each_tree_to_the_requested_size(host_pb, pymcts)


def test_device_mcts_reports_search_from_finished_env(host_pb):
    import test_mcts_gpu as tg

    tg.test_search_from_a_finished_env_is_reported(host_pb)


@pytest.mark.parametrize("tpw", ["4", "32"])
def test_trees_per_warp_mapping_does_not_change_results(host_pb, pymcts, tpw, monkeypatch):
    """mcts_tree_of_thread: the lane -> tree map used when trees outnumber warps."""
    import test_mcts_gpu as tg

    monkeypatch.setenv("PB_MCTS_TREES_PER_WARP", tpw)
    tg.test_ponder_grows_each_tree_to_the_requested_size(host_pb, pymcts)


def test_product_collector_runs_over_the_host_kernels_and_matches_oracle(host_pb, pymcts):
    """test_mcts_gpu.test_collector_matches_oracle: the product's BatchedExperienceCollector
    (pacbot-rl_b200/alphazero.py, unmodified -- search budget, episode records, truncation
    bootstrap, discounted returns, hand-out order) over the host build of k_mcts_collect_act and
    friends, against the oracle's ExperienceCollector: items bit for bit."""
    import test_mcts_gpu as tg

    tg.test_collector_matches_oracle(host_pb, pymcts)
