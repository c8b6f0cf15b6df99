"""ctypes binding of the CPU ORACLE (oracle/libpacbot_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, `__graft_entry__.smoke()` and bench.py's
`cpu_baseline` / `--impl reference` legs; never by the product package.  PARITY UNPINNED at the
engine boundary -- see oracle/pacbot_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpacbot_oracle.so")

OBS_SHAPE = (17, 28, 31)
OBS_FLOATS = 17 * 28 * 31

FLAT_DTYPE = np.dtype(
    [
        ("curr_ticks", "<u4"),
        ("rng_ctr", "<u4"),
        ("pellets", "<u4", (31,)),
        ("level_steps", "<u2"),
        ("curr_score", "<u2"),
        ("num_pellets", "<u2"),
        ("last_score", "<u2"),
        ("update_period", "u1"),
        ("mode", "u1"),
        ("paused", "u1"),
        ("pause_on_update", "u1"),
        ("mode_steps", "u1"),
        ("curr_level", "u1"),
        ("curr_lives", "u1"),
        ("ghost_combo", "u1"),
        ("fruit_steps", "u1"),
        ("last_action", "u1"),
        ("ticks_per_step", "u1"),
        ("random_start", "u1"),
        ("random_ticks", "u1"),
        ("randomize_ghosts", "u1"),
        ("remove_super_pellets", "u1"),
        ("pac_row", "i1"),
        ("pac_col", "i1"),
        ("pac_dir", "u1"),
        ("g_row", "i1", (4,)),
        ("g_col", "i1", (4,)),
        ("g_dir", "u1", (4,)),
        ("g_next_row", "i1", (4,)),
        ("g_next_col", "i1", (4,)),
        ("g_next_dir", "u1", (4,)),
        ("g_fright", "u1", (4,)),
        ("g_trapped", "u1", (4,)),
        ("g_spawning", "u1", (4,)),
        ("g_eaten", "u1", (4,)),
    ],
    align=True,
)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc); returns the .so path."""
    srcs = [
        os.path.join(_HERE, f)
        for f in ("pacbot_oracle.c", "pacbot_oracle.h", "mcts_oracle.c", "mcts_oracle.h", "Makefile")
    ]
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, u8p, i32p, u32p, f32p = (
            C.c_void_p,
            C.c_int64,
            C.POINTER(C.c_uint8),
            C.POINTER(C.c_int32),
            C.POINTER(C.c_uint32),
            C.POINTER(C.c_float),
        )
        L.po_batch_create.restype = vp
        L.po_batch_create.argtypes = [i64, vp, vp, i64]
        L.po_batch_destroy.argtypes = [vp]
        L.po_batch_reset.argtypes = [vp, vp]
        L.po_batch_step.restype = i64
        L.po_batch_step.argtypes = [vp, vp, vp, vp, C.c_int]
        L.po_batch_action_mask.argtypes = [vp, vp]
        L.po_batch_obs.argtypes = [vp, vp]
        L.po_batch_export.argtypes = [vp, vp]
        L.po_batch_import.argtypes = [vp, vp]
        L.po_batch_obs_range.argtypes = [vp, i64, i64, vp]
        L.po_batch_set_counter_rng.argtypes = [vp, C.c_uint64, i64]
        L.po_batch_sample_actions.argtypes = [vp, C.c_uint64, i64, C.c_uint64, vp]
        L.po_counter_draw_raw.restype = C.c_uint32
        L.po_counter_draw_raw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.po_counter_action_raw.restype = C.c_uint32
        L.po_counter_action_raw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.po_batch_env.restype = vp
        L.po_batch_env.argtypes = [vp, i64]
        L.po_gym_first_ai_done.argtypes = [vp]
        L.po_gym_first_ai_done.restype = C.c_int
        L.po_gym_is_done.argtypes = [vp]
        L.po_gym_is_done.restype = C.c_int
        L.po_wall_at.argtypes = [C.c_int, C.c_int]
        L.po_wall_at.restype = C.c_int
        L.po_sizeof_flat.restype = C.c_size_t
        L.po_sizeof_gym.restype = C.c_size_t
        L.po_rollout.restype = C.c_double
        L.po_rollout.argtypes = [
            i64,
            i64,
            i64,
            C.c_uint64,
            C.c_int,
            C.POINTER(C.c_uint64),
            C.POINTER(C.c_int64),
        ]
        assert L.po_sizeof_flat() == FLAT_DTYPE.itemsize, (L.po_sizeof_flat(), FLAT_DTYPE.itemsize)
        _lib = L
    return _lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


CFG_RANDOM_START = 1
CFG_RANDOM_TICKS = 2
CFG_RANDOMIZE_GHOSTS = 4
CFG_REMOVE_SUPER_PELLETS = 8


class OracleBatch:
    """N independent scalar oracle envs with an injected raw-draw tape (uint32 [N, tape_len])."""

    def __init__(self, n: int, cfg_flags=None, tape: np.ndarray | None = None):
        self.n = int(n)
        L = lib()
        if cfg_flags is None:
            cfg_flags = np.zeros(self.n, np.uint8)
        self.cfg = np.ascontiguousarray(cfg_flags, dtype=np.uint8)
        assert self.cfg.shape == (self.n,)
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint32)
            assert tape.ndim == 2 and tape.shape[0] == self.n
        self.tape = tape  # keep alive: the C side borrows the pointer
        self._h = L.po_batch_create(
            self.n, _ptr(self.cfg), _ptr(tape), 0 if tape is None else tape.shape[1]
        )

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # module globals are already torn down at interpreter exit
            lib().po_batch_destroy(h)

    def reset(self, mask: np.ndarray | None = None) -> None:
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        lib().po_batch_reset(self._h, _ptr(mask))

    def step(self, actions: np.ndarray, auto_reset: bool = False):
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        assert actions.shape == (self.n,)
        reward = np.zeros(self.n, np.int32)
        done = np.zeros(self.n, np.uint8)
        bad = lib().po_batch_step(
            self._h, _ptr(actions), _ptr(reward), _ptr(done), int(auto_reset)
        )
        if bad >= 0:
            raise ValueError("Invalid action")
        return reward, done.astype(bool)

    def action_mask(self) -> np.ndarray:
        out = np.zeros((self.n, 5), np.uint8)
        lib().po_batch_action_mask(self._h, _ptr(out))
        return out.astype(bool)

    def obs(self) -> np.ndarray:
        out = np.empty((self.n,) + OBS_SHAPE, np.float32)
        lib().po_batch_obs(self._h, _ptr(out))
        return out

    def obs_range(self, first: int, count: int) -> np.ndarray:
        """obs() of envs [first, first + count) (large batches are compared chunk by chunk)."""
        assert 0 <= first and first + count <= self.n
        out = np.empty((count,) + OBS_SHAPE, np.float32)
        lib().po_batch_obs_range(self._h, first, count, _ptr(out))
        return out

    def set_counter_rng(self, seed: int, gid_base: int = 0) -> None:
        """Replace the tape by the product's counter RNG keyed (seed, gid_base + i): what the
        CUDA engine draws when no tape is installed (include/pacbot_b200.h "Counter RNG")."""
        lib().po_batch_set_counter_rng(self._h, seed & 0xFFFFFFFFFFFFFFFF, gid_base)

    def sample_actions(self, seed: int, gid_base: int, call_idx: int) -> np.ndarray:
        """pb_sample_actions / pb_step_random's choice: uniformly random valid action per env."""
        out = np.zeros(self.n, np.uint8)
        lib().po_batch_sample_actions(self._h, seed & 0xFFFFFFFFFFFFFFFF, gid_base, call_idx, _ptr(out))
        return out

    def export_state(self) -> np.ndarray:
        out = np.zeros(self.n, FLAT_DTYPE)
        lib().po_batch_export(self._h, _ptr(out))
        return out

    def import_state(self, st: np.ndarray) -> None:
        st = np.ascontiguousarray(st, dtype=FLAT_DTYPE)
        assert st.shape == (self.n,)
        lib().po_batch_import(self._h, _ptr(st))

    def first_ai_done(self) -> np.ndarray:
        L = lib()
        return np.array(
            [bool(L.po_gym_first_ai_done(L.po_batch_env(self._h, i))) for i in range(self.n)]
        )

    def is_done(self) -> np.ndarray:
        L = lib()
        return np.array([bool(L.po_gym_is_done(L.po_batch_env(self._h, i))) for i in range(self.n)])


def flat_from_product(st: np.ndarray) -> np.ndarray:
    """Product `pb_env_state` records (pacbot_rl_b200.ENV_STATE_DTYPE, what `get_state` returns)
    -> oracle flat states for `OracleBatch.import_state`."""
    out = np.zeros(len(st), FLAT_DTYPE)
    for f in out.dtype.names:
        if f in st.dtype.names:
            out[f] = st[f]
    out["random_start"] = st["cfg_flags"] & 1
    out["random_ticks"] = (st["cfg_flags"] >> 1) & 1
    out["randomize_ghosts"] = (st["cfg_flags"] >> 2) & 1
    out["remove_super_pellets"] = (st["cfg_flags"] >> 3) & 1
    return out


def cfg_flags_of(flat: np.ndarray) -> np.ndarray:
    """cfg_flags byte of the product's state record from an oracle flat state."""
    return (flat["random_start"].astype(np.uint8) | (flat["random_ticks"].astype(np.uint8) << 1)
            | (flat["randomize_ghosts"].astype(np.uint8) << 2)
            | (flat["remove_super_pellets"].astype(np.uint8) << 3))


# fields shared by the oracle's flat state and the product's pb_env_state
COMMON_FIELDS = [
    "curr_ticks", "rng_ctr", "pellets", "level_steps", "curr_score", "num_pellets", "last_score",
    "update_period", "mode", "paused", "pause_on_update", "mode_steps", "curr_level",
    "curr_lives", "ghost_combo", "fruit_steps", "last_action", "ticks_per_step",
    "pac_row", "pac_col", "pac_dir", "g_row", "g_col", "g_dir", "g_next_row", "g_next_col",
    "g_next_dir", "g_fright", "g_trapped", "g_spawning", "g_eaten",
]


def states_differ(flat: np.ndarray, product: np.ndarray):
    """None if an oracle flat-state array and a product state array describe the same envs,
    else a one-line description of the first difference."""
    for f in COMMON_FIELDS:
        a, b = flat[f], product[f]
        if not np.array_equal(a, b):
            bad = np.flatnonzero(np.asarray(a != b).reshape(len(a), -1).any(axis=1))
            i = int(bad[0])
            return f"field {f!r} differs for {len(bad)} envs; first env {i}: oracle={a[i]!r} gpu={b[i]!r}"
    if not np.array_equal(cfg_flags_of(flat), product["cfg_flags"]):
        return "cfg_flags differ"
    return None


def rollout(n_envs: int, n_steps: int, warmup: int = 0, seed: int = 0, n_threads: int = 1):
    """CPU baseline: returns (seconds, env_steps, checksum, episodes)."""
    cs = C.c_uint64(0)
    ep = C.c_int64(0)
    secs = lib().po_rollout(
        n_envs, n_steps, warmup, seed, n_threads, C.byref(cs), C.byref(ep)
    )
    return secs, n_envs * n_steps, cs.value, ep.value
