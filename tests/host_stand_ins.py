"""Host stand-ins shared by tests/test_host_mcts.py and tests/test_host_replay.py: the host build
of the product's device code (tests/host_engine/host_mcts.cpp), `HostEnv` (the engine object the
product's Python host code is handed instead of `BatchedPacmanGym`, which refuses to exist
without a CUDA device), `HostABI` (the pb_mcts_* entry points routed to the host build) and the
`host_pb` fixture that wires them in.  TEST INFRASTRUCTURE ONLY."""
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
    L.hm_mcts_backprop_collect.restype = C.c_int
    L.hm_mcts_backprop_collect.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int] + [vp] * 7 + [C.c_float, C.c_float]
    L.hm_engine_copy_envs.restype = C.c_int
    L.hm_engine_copy_envs.argtypes = [vp, vp, vp, vp, i64]
    L.hm_engine_step.restype = i64
    L.hm_engine_step.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    L.hm_engine_set_tape.restype = None
    L.hm_engine_set_tape.argtypes = [vp, vp, i64]
    for name, args in dict(
            hm_engine_destroy=[vp], hm_engine_reset=[vp, vp], hm_engine_get_state=[vp, vp],
            hm_engine_set_state=[vp, vp], hm_engine_action_mask=[vp, vp], hm_mcts_destroy=[vp],
            hm_mcts_reset_root=[vp, vp, vp, vp], hm_mcts_root_noise=[vp, vp, vp, C.c_float],
            hm_mcts_root_noise_dirichlet=[vp, vp, C.c_float, C.c_float],
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

    @property
    def _L(self):          # `env._L`: the library object the product's host code calls into
        return HostABI(self.L)

    def __init__(self, L, pb, n, random_start=False, random_ticks=False, *, device=None, seed=0,
                 env_id_base=0, randomize_ghosts=False, remove_super_pellets=False, auto_reset=False):
        self.L, self.pb, self.num_envs, self.seed, self.env_id_base = L, pb, n, seed, env_id_base
        self.auto_reset, self._tape = bool(auto_reset), None

        def bits(v, shift):
            return np.broadcast_to(np.asarray(v, bool), (n,)).astype(np.uint8) << shift

        self.cfg = bits(random_start, 0) | bits(random_ticks, 1) | bits(randomize_ghosts, 2) | bits(
            remove_super_pellets, 3)
        self.h = L.hm_engine_create(n, _ptr(np.ascontiguousarray(self.cfg)), seed, env_id_base)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hm_engine_destroy(self.h)
            self.h = None

    def reset(self, mask=None):
        m = _mask_np(mask, self.num_envs)
        self.L.hm_engine_reset(self.h, _ptr(m))

    def get_state(self, first=0, count=None):
        st = np.zeros(self.num_envs, self.pb.ENV_STATE_DTYPE)
        self.L.hm_engine_get_state(self.h, st.ctypes.data)
        return st[first:(self.num_envs if count is None else first + count)].copy()

    def set_state(self, st, first=0):
        full = self.get_state()
        full[first:first + len(st)] = np.ascontiguousarray(st, self.pb.ENV_STATE_DTYPE)
        self.L.hm_engine_set_state(self.h, full.ctypes.data)
        self.cfg = (full["cfg_flags"] & 0x0F).astype(np.uint8)

    # ---- what pacbot-rl_b200/gym.py (the per-env PacmanGym drop-in) asks of its engine -------------
    @property
    def _cfg(self):
        return self.cfg

    @property
    def random_start(self):
        return (self.cfg & 1).astype(bool)

    @random_start.setter
    def random_start(self, value):   # pb_set_cfg_flags
        st = self.get_state()
        v = np.broadcast_to(np.asarray(value, bool), (self.num_envs,)).astype(np.uint8)
        st["cfg_flags"] = (st["cfg_flags"] & ~np.uint8(1)) | v
        self.set_state(st)

    def step_host(self, actions, obs_out=None, want_mask=True, out=None):   # pb_step_host
        n = self.num_envs
        a = np.ascontiguousarray(actions, np.uint8)
        reward, done, mask = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros((n, 5), np.uint8)
        bad = self.L.hm_engine_step(self.h, a.ctypes.data, reward.ctypes.data, done.ctypes.data,
                                    mask.ctypes.data, int(self.auto_reset))
        if bad >= 0:
            raise ValueError("Invalid action")   # PB_ERR_INVALID_ACTION through _lib.check
        return reward, done.astype(bool), (mask.astype(bool) if want_mask else None)

    def step_snapshot(self, actions=None, want_obs=True):   # pb_step_snapshot_host
        n = self.num_envs
        out = {}
        if actions is not None:
            r, d, _ = self.step_host(actions, want_mask=False)
            out["reward"], out["done"] = r, d
        st = self.get_state()
        info = np.zeros((n, 8), np.int32)
        info[:, 0], info[:, 1] = st["curr_score"], st["curr_lives"]
        info[:, 2] = (st["curr_lives"] < 3) | (st["curr_level"] != 1)
        info[:, 3] = self.first_ai_done().numpy()
        info[:, 4] = st["num_pellets"]
        out["info"] = info
        out["mask"] = self.action_mask().numpy()
        if want_obs:
            out["obs"] = self.obs_numpy()
        return out

    def obs_numpy(self, first=0, count=None):   # pb_obs_host
        count = self.num_envs - first if count is None else count
        out = torch.empty((count, 17, 28, 31), dtype=torch.float32)
        return self.obs_range(first, count, out).numpy().copy()

    def clone_env_from(self, dst_index, src, src_index):   # pb_clone_env
        self.copy_envs_from(src, torch.tensor([src_index]), torch.tensor([dst_index]))
        self.cfg[dst_index] = src.cfg[src_index]

    def view(self, index):
        return self.pb.gym.PacmanGymView(self, index)

    def score(self):
        return self._scalar(lambda s: int(s["curr_score"])).to(torch.int32)

    def lives(self):
        return self._scalar(lambda s: int(s["curr_lives"])).to(torch.int32)

    def action_mask(self, out=None):
        m = np.zeros((self.num_envs, 5), np.uint8)
        self.L.hm_engine_action_mask(self.h, m.ctypes.data)
        t = torch.from_numpy(m.astype(bool))
        if out is not None:
            out.copy_(t)
            return out
        return t

    def set_draw_tape(self, tape):
        """pb_set_draw_tape: raw u32 draws [n, tape_len] (the GPU API takes them as int32 bits)."""
        self._tape = None if tape is None else np.ascontiguousarray(tape.cpu().numpy()).view(np.uint32)
        self.L.hm_engine_set_tape(self.h, _ptr(self._tape), 0 if self._tape is None else self._tape.shape[1])

    def step(self, actions, *, check_actions=True, mask_out=None):
        n = self.num_envs
        a = actions.cpu().numpy() if isinstance(actions, torch.Tensor) else np.asarray(actions)
        a = np.ascontiguousarray(np.where((a < 0) | (a > 4), 255, a).astype(np.uint8))
        reward, done, mask = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros((n, 5), np.uint8)
        bad = self.L.hm_engine_step(self.h, a.ctypes.data, reward.ctypes.data, done.ctypes.data,
                                    mask.ctypes.data, int(self.auto_reset))
        if check_actions and bad >= 0:
            raise ValueError("Invalid action")
        if mask_out is not None:
            mask_out.copy_(torch.from_numpy(mask.astype(bool)))
        return torch.from_numpy(reward), torch.from_numpy(done.astype(bool))

    def step_obs(self, actions, obs_out, *, check_actions=False, mask_out=None):
        r, d = self.step(actions, check_actions=check_actions, mask_out=mask_out)
        self.obs(out=obs_out)
        return r, d

    def obs_ring(self, ring, slot_index):
        self.obs(out=ring[int(slot_index)])

    def _scalar(self, fn):
        return torch.from_numpy(np.array([fn(s) for s in self.get_state()]))

    def first_ai_done(self):   # env.rs:281-285
        def f(s):
            sup = any((int(s["pellets"][r]) >> c) & 1 for r, c in ((3, 1), (3, 26), (23, 1), (23, 26)))
            return (not sup) and not (s["g_fright"] > 0).any()
        return self._scalar(f).bool()

    def remaining_pellets(self):
        return self._scalar(lambda s: int(s["num_pellets"])).to(torch.int32)

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


class HostABI:
    """The pb_mcts_* entry points of include/pacbot_b200.h (argument lists as in the C ABI, stream
    ignored) on top of the host build: what pacbot-rl_b200/mcts.py and alphazero.py call through
    `env._L`.  Handles are the harness's pointers; tensors are CPU tensors, so `data_ptr()` is a
    host pointer."""

    last_error = ""

    def __init__(self, L):
        self._lib = L

    def pb_mcts_create(self, out_ref, root_h, sim_h, max_nodes):
        factor = int(os.environ.get("PB_MCTS_POOL_FACTOR", "8"))     # as pb_mcts_create reads them
        tpw = int(os.environ.get("PB_MCTS_TREES_PER_WARP", "1"))
        out_ref._obj.value = self._lib.hm_mcts_create(root_h, sim_h, max_nodes, factor, tpw)
        return 0

    def pb_mcts_destroy(self, h):
        self._lib.hm_mcts_destroy(h)
        return 0

    def pb_mcts_reset_root(self, h, mask, value, policy, stream):
        self._lib.hm_mcts_reset_root(h, mask, value, policy)
        return 0

    def pb_mcts_root_noise(self, h, mask, noise, mix, stream):
        self._lib.hm_mcts_root_noise(h, mask, noise, mix)
        return 0

    def pb_mcts_root_noise_dirichlet(self, h, mask, alpha, mix, stream):
        self._lib.hm_mcts_root_noise_dirichlet(h, mask, alpha, mix)
        return 0

    def pb_mcts_select(self, h, active, leaf_kind, leaf_mask, stream):
        self._lib.hm_mcts_select(h, active, leaf_kind, leaf_mask)
        return 0

    def pb_mcts_backprop(self, h, value, policy, stream):
        self._lib.hm_mcts_backprop(h, value, policy)
        return 0

    def pb_mcts_take_action(self, h, mask, actions, reward, done, status, stream):
        self._lib.hm_mcts_take_action(h, mask, actions, reward, done, status)
        return 0

    def pb_mcts_query(self, h, what, out, stream):
        self._lib.hm_mcts_query(h, what, out)
        return 0

    def pb_mcts_check(self, h, epoch_ref, stream):
        err = self._lib.hm_mcts_check(h, C.addressof(epoch_ref._obj))
        if err:   # the message pb_mcts_check composes (pb_mcts_abi.inl)
            HostABI.last_error = "pb_mcts: " + "; ".join(v for k, v in ERR.items() if err & k)
            return -1
        return 0

    def pb_mcts_collect_act(self, h, snap_h, max_len, tree_size, remaining, ep_len, ep_mask, ep_dist,
                            ep_value, status, flags, stream):
        return self._lib.hm_mcts_collect_act(h, snap_h, max_len, tree_size, remaining, ep_len, ep_mask,
                                             ep_dist, ep_value, status, flags)


    def pb_mcts_backprop_collect(self, h, value, policy, snap_h, max_len, tree_size, remaining, ep_len,
                                 ep_mask, ep_dist, ep_value, status, flags, alpha, mix, publish, stream):
        assert not publish     # the pinned host word is a GPU facility
        return self._lib.hm_mcts_backprop_collect(h, value, policy, snap_h, max_len, tree_size, remaining,
                                                  ep_len, ep_mask, ep_dist, ep_value, status, flags,
                                                  alpha, mix)


def _check(rc):   # pacbot-rl_b200/_lib.py::check with the adapter's error text
    if rc != 0:
        raise ValueError(HostABI.last_error)


@pytest.fixture
def host_pb(hm, pb, oracle, monkeypatch):
    """`pb` with `BatchedPacmanGym` replaced by the host stand-in wherever the product's host code
    resolves that name.  pacbot-rl_b200/mcts.py (`BatchedMCTS`) and alphazero.py
    (`BatchedExperienceCollector`) then run UNMODIFIED -- tensors on the CPU, C calls through
    HostABI into the host build of the kernels.  `Tensor.cuda()` becomes the identity (the re-used
    GPU test bodies call it on their masks) and the CUDA stream argument is dropped."""
    def make_env(n, rs=False, rt=False, **kw):
        return HostEnv(hm, pb, n, rs, rt, **kw)

    monkeypatch.setattr(HostEnv, "oracle", oracle)
    for mod in (pb, pb.mcts, pb.alphazero, pb.replay_buffer, pb.gym):
        monkeypatch.setattr(mod, "BatchedPacmanGym", make_env)
    monkeypatch.setattr(pb.mcts.BatchedMCTS, "_stream", lambda self: None)
    monkeypatch.setattr(pb.mcts, "check", _check)
    monkeypatch.setattr(pb.alphazero, "check", _check)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return pb


