"""ctypes binding of the MCTS / AlphaZero-collector part of the CPU ORACLE (oracle/mcts_oracle.c).

TEST INFRASTRUCTURE ONLY (same rules as pyoracle.py).  Evaluator and Dirichlet noise are Python
callables so that tests can feed the oracle and the CUDA path the very same numbers:

    evaluator(obs f32 [B,17,28,31], mask bool [B,5]) -> (value f32 [B], policy f32 [B,5])
    noise(tree: int, num_valid: int) -> f32 [num_valid]
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import pyoracle
from .pyoracle import OBS_FLOATS, OBS_SHAPE

_EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_uint8),
                       C.POINTER(C.c_float), C.POINTER(C.c_float))
_NOISE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float))


class _Config(C.Structure):
    _fields_ = [("tree_size", C.c_int64), ("max_episode_length", C.c_int64),
                ("discount_factor", C.c_float), ("num_parallel_envs", C.c_int64)]


ITEM_DTYPE = np.dtype([("obs", "<f4", OBS_SHAPE), ("action_mask", "u1", (5,)), ("value", "<f4"),
                       ("action_distribution", "<f4", (5,))], align=True)

_bound = False


def lib() -> C.CDLL:
    global _bound
    L = pyoracle.lib()
    if not _bound:
        vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
        L.pm_draw_raw.restype = C.c_uint32
        L.pm_draw_raw.argtypes = [u64, u64, C.c_uint32]
        L.pm_sim_seed.restype = u64
        L.pm_sim_seed.argtypes = [u64, u64]
        L.pm_batch_create.restype = vp
        L.pm_batch_create.argtypes = [i64, vp, vp, u64, u64, _EVAL_FN, vp, _NOISE_FN, vp]
        L.pm_batch_destroy.argtypes = [vp]
        L.pm_batch_env.restype = vp
        L.pm_batch_env.argtypes = [vp, i64]
        L.pm_batch_epoch.restype = u64
        L.pm_batch_epoch.argtypes = [vp]
        L.pm_batch_reset.argtypes = [vp, vp, C.c_int]
        L.pm_batch_search_iteration.argtypes = [vp, vp]
        L.pm_batch_take_action.argtypes = [vp, vp, vp, vp, vp]
        L.pm_best_action.restype = C.c_int
        L.pm_best_action.argtypes = [vp, i64]
        for name in ("pm_action_distribution", "pm_policy_prior", "pm_action_values"):
            getattr(L, name).argtypes = [vp, i64, vp]
        L.pm_value.restype = C.c_float
        L.pm_value.argtypes = [vp, i64]
        L.pm_max_depth.restype = i64
        L.pm_max_depth.argtypes = [vp, i64]
        L.pm_node_count.restype = i64
        L.pm_node_count.argtypes = [vp, i64]
        L.pm_root_visits.restype = C.c_uint32
        L.pm_root_visits.argtypes = [vp, i64]
        L.pm_collector_create.restype = vp
        L.pm_collector_create.argtypes = [_Config, u64, u64, _EVAL_FN, vp, _NOISE_FN, vp]
        L.pm_collector_destroy.argtypes = [vp]
        L.pm_collector_generate.restype = i64
        L.pm_collector_generate.argtypes = [vp, C.POINTER(vp)]
        L.pm_collector_batch.restype = vp
        L.pm_collector_batch.argtypes = [vp]
        L.po_gym_obs.argtypes = [vp, vp]
        L.po_gym_action_mask.argtypes = [vp, vp]
        L.po_gym_score.restype = C.c_uint32
        L.po_gym_score.argtypes = [vp]
        L.po_gym_export.argtypes = [vp, vp]
        assert ITEM_DTYPE.itemsize == OBS_FLOATS * 4 + 8 + 4 + 20, ITEM_DTYPE.itemsize
        _bound = True
    return L


def _callbacks(evaluator, noise):
    def _eval(_ctx, tree, obs_p, mask_p, value_p, policy_p):
        obs = np.ctypeslib.as_array(obs_p, shape=(1,) + OBS_SHAPE).copy()
        mask = np.ctypeslib.as_array(mask_p, shape=(1, 5)).astype(bool)
        value, policy = evaluator(obs, mask)
        value_p[0] = np.float32(np.asarray(value).reshape(-1)[0])
        pol = np.asarray(policy, np.float32).reshape(-1)
        for a in range(5):
            policy_p[a] = pol[a]

    def _noise(_ctx, tree, num_valid, out_p):
        w = np.asarray(noise(int(tree), int(num_valid)), np.float32).reshape(-1)
        assert w.shape[0] == num_valid
        for k in range(num_valid):
            out_p[k] = w[k]

    return _EVAL_FN(_eval), _NOISE_FN(_noise)


class _TreeQueries:
    """Per-tree getters shared by OracleMCTS and OracleCollector (handle in self._b)."""

    def best_action(self, i: int) -> int:
        return int(lib().pm_best_action(self._b, i))

    def _vec5(self, fn, i):
        out = np.zeros(5, np.float32)
        fn(self._b, i, out.ctypes.data_as(C.c_void_p))
        return out

    def action_distribution(self, i: int) -> np.ndarray:
        return self._vec5(lib().pm_action_distribution, i)

    def policy_prior(self, i: int) -> np.ndarray:
        return self._vec5(lib().pm_policy_prior, i)

    def action_values(self, i: int) -> np.ndarray:
        return self._vec5(lib().pm_action_values, i)

    def value(self, i: int) -> np.float32:
        return np.float32(lib().pm_value(self._b, i))

    def max_depth(self, i: int) -> int:
        return int(lib().pm_max_depth(self._b, i))

    def node_count(self, i: int) -> int:
        return int(lib().pm_node_count(self._b, i))

    def root_visits(self, i: int) -> int:
        return int(lib().pm_root_visits(self._b, i))

    def root_obs(self, i: int) -> np.ndarray:
        out = np.empty(OBS_SHAPE, np.float32)
        lib().po_gym_obs(lib().pm_batch_env(self._b, i), out.ctypes.data_as(C.c_void_p))
        return out

    def root_mask(self, i: int) -> np.ndarray:
        out = np.zeros(5, np.uint8)
        lib().po_gym_action_mask(lib().pm_batch_env(self._b, i), out.ctypes.data_as(C.c_void_p))
        return out.astype(bool)

    def score(self, i: int) -> int:
        return int(lib().po_gym_score(lib().pm_batch_env(self._b, i)))

    def export_state(self) -> np.ndarray:
        out = np.zeros(self.n, pyoracle.FLAT_DTYPE)
        for i in range(self.n):
            lib().po_gym_export(lib().pm_batch_env(self._b, i), out[i:i + 1].ctypes.data_as(C.c_void_p))
        return out


class OracleMCTS(_TreeQueries):
    """n lock-stepped MCTSContext restatements (mcts.rs) with injected evaluator / noise."""

    def __init__(self, n, random_start, random_ticks, seed, gid_base, evaluator, noise):
        self.n = int(n)
        self._cbs = _callbacks(evaluator, noise)  # keep alive
        rs = np.ascontiguousarray(np.broadcast_to(np.asarray(random_start, np.uint8), (self.n,)))
        rt = np.ascontiguousarray(np.broadcast_to(np.asarray(random_ticks, np.uint8), (self.n,)))
        self._b = lib().pm_batch_create(self.n, rs.ctypes.data_as(C.c_void_p), rt.ctypes.data_as(C.c_void_p),
                                        seed, gid_base, self._cbs[0], None, self._cbs[1], None)

    def __del__(self):
        b, self._b = getattr(self, "_b", None), None
        if b:
            lib().pm_batch_destroy(b)

    @staticmethod
    def _mask(mask):
        if mask is None:
            return None, None
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        return m, m.ctypes.data_as(C.c_void_p)

    def reset(self, mask=None, reset_env: bool = True) -> None:
        m, p = self._mask(mask)
        lib().pm_batch_reset(self._b, p, int(reset_env))

    def search_iteration(self, active=None) -> None:
        m, p = self._mask(active)
        lib().pm_batch_search_iteration(self._b, p)

    def take_action(self, actions, mask=None):
        m, p = self._mask(mask)
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        reward = np.zeros(self.n, np.int32)
        done = np.zeros(self.n, np.uint8)
        lib().pm_batch_take_action(self._b, p, a.ctypes.data_as(C.c_void_p),
                                   reward.ctypes.data_as(C.c_void_p), done.ctypes.data_as(C.c_void_p))
        return reward, done.astype(bool)

    @property
    def epoch(self) -> int:
        return int(lib().pm_batch_epoch(self._b))


class OracleCollector(_TreeQueries):
    """ExperienceCollector restatement (alphazero.rs)."""

    def __init__(self, tree_size, max_episode_length, discount_factor, num_parallel_envs, seed,
                 gid_base, evaluator, noise):
        self.n = int(num_parallel_envs)
        self._cbs = _callbacks(evaluator, noise)
        cfg = _Config(tree_size, max_episode_length, discount_factor, num_parallel_envs)
        self._c = lib().pm_collector_create(cfg, seed, gid_base, self._cbs[0], None, self._cbs[1], None)
        self._b = lib().pm_collector_batch(self._c)

    def __del__(self):
        c, self._c = getattr(self, "_c", None), None
        if c:
            lib().pm_collector_destroy(c)

    def generate_experience(self) -> np.ndarray:
        """Structured array of ITEM_DTYPE (a copy)."""
        p = C.c_void_p()
        k = lib().pm_collector_generate(self._c, C.byref(p))
        buf = (C.c_char * (k * ITEM_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=ITEM_DTYPE, count=k).copy()
