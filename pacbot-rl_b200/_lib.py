"""ctypes loader of the C-ABI library (include/pacbot_b200.h).

There is deliberately no fallback: if the CUDA extension is missing or no CUDA device is
visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpacbot_b200.so")

PB_OK = 0
PB_ERR_INVALID_ARG = -1
PB_ERR_CUDA = -2
PB_ERR_INVALID_ACTION = -3
PB_ERR_NO_DEVICE = -4

OBS_SHAPE = (17, 28, 31)
OBS_FLOATS = 17 * 28 * 31
NUM_ACTIONS = 5

CFG_RANDOM_START = 1
CFG_RANDOM_TICKS = 2
CFG_RANDOMIZE_GHOSTS = 4
CFG_REMOVE_SUPER_PELLETS = 8

Q_SCORE, Q_LIVES, Q_IS_DONE, Q_FIRST_AI_DONE, Q_REMAINING_PELLETS = 0, 1, 2, 3, 4
Q_TICKS_PER_STEP, Q_CURR_TICKS, Q_LEVEL, Q_UPDATE_PERIOD = 5, 6, 7, 8

# numpy mirror of `pb_env_state` (include/pacbot_b200.h)
ENV_STATE_DTYPE = np.dtype(
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
        ("cfg_flags", "u1"),
        ("pac_row", "i1"),
        ("pac_col", "i1"),
        ("pac_dir", "u1"),
        ("reserved0", "u1"),
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

# every symbol include/pacbot_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _u64, _int = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
SIGNATURES = {
    "pb_version": (_int, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_sizeof_env_state": (C.c_size_t, []),
    "pb_state_bytes_per_env": (_int, []),
    "pb_draw_raw": (C.c_uint32, [_u64, _u64, C.c_uint32]),
    "pb_action_raw": (C.c_uint32, [_u64, _u64, _u64]),
    "pb_maze_wall_row": (C.c_uint32, [_int]),
    "pb_maze_pellet_row": (C.c_uint32, [_int]),
    "pb_maze_node": (_int, [_int, C.POINTER(_int), C.POINTER(_int)]),
    "pb_create": (_int, [C.POINTER(_vp), _i64, _int, _u64, _i64, _vp]),
    "pb_destroy": (_int, [_vp]),
    "pb_num_envs": (_i64, [_vp]),
    "pb_device": (_int, [_vp]),
    "pb_set_cfg_flags": (_int, [_vp, _i64, _i64, _vp]),
    "pb_set_draw_tape": (_int, [_vp, _vp, _i64]),
    "pb_reset": (_int, [_vp, _vp, _vp]),
    "pb_step": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_encode_obs": (_int, [_vp, _vp, _i64, _vp]),
    "pb_encode_obs_nhwc32": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "pb_encode_obs_range": (_int, [_vp, _i64, _i64, _vp, _i64, _vp]),
    "pb_encode_obs_ring": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "pb_encode_obs_ring_nhwc32": (_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "pb_step_encode": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_step_encode_small": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp, _i64, _vp, _vp]),
    "pb_action_mask": (_int, [_vp, _vp, _vp]),
    "pb_query": (_int, [_vp, _int, _vp, _vp]),
    "pb_sample_actions": (_int, [_vp, _vp, _u64, _vp]),
    "pb_step_random": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_step_random_range": (_int, [_vp, _i64, _i64, _u64, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_step_random_flip": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_step_flip": (_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_rollout_random_step": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_rollout_host_step": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_rollout_chunk_step": (_int, [_vp, _i64, _i64, _vp, _u64, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_rollout_chunk_host_step": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_rollout_host_sync": (_int, [_vp]),
    "pb_rollout_join": (_int, [_vp, _vp]),
    "pb_qnet_head_fwd": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _int, _vp]),
    "pb_qnet_head_fwd_train": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp,
                               _vp, _vp, _int, _vp]),
    "pb_qnet_head_greedy": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp,
                            _int, _vp]),
    "pb_qnet_head_policy_value": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_float, _vp, _vp,
                                  _vp, _int, _vp]),
    "pb_qnet_head_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _vp,
                         _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_qnet_head_bwd_scratch_floats": (_i64, [_i64]),
    "pb_dqn_target": (_int, [_vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _i64, _i64, _int, _vp]),
    "pb_dqn_loss": (_int, [_vp, _vp, _vp, _i64, _i64, C.c_double, _vp, _vp, _int, _vp]),
    "pb_clip_adam": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double,
                     C.c_double, C.c_double, _vp, _int, _vp]),
    "pb_replay_advance": (_int, [_vp, _vp, _i64, _int, _vp]),
    "pb_step_snapshot_host": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pb_snapshot_block": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "pb_rollout_streams": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "pb_check_actions": (_int, [_vp, C.POINTER(_i64), _vp]),
    "pb_get_state": (_int, [_vp, _i64, _i64, _vp]),
    "pb_set_state": (_int, [_vp, _i64, _i64, _vp]),
    "pb_get_state_on": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "pb_set_state_on": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "pb_clone_env": (_int, [_vp, _i64, _vp, _i64, _vp]),
    "pb_step_host": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp]),
    "pb_step_host_begin": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_step_host_end": (_int, [_vp, _vp]),
    "pb_obs_host": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "pb_select_actions": (_int, [_vp, _vp, _vp, C.c_float, _vp, _u64, _vp]),
    "pb_select_actions_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pb_replay_raw": (C.c_uint32, [_u64, _u64, C.c_uint32]),
    "pb_replay_sample_capacity": (_i64, []),
    "pb_replay_record": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "pb_replay_sample": (_int, [_vp, _i64, _i64, _u64, _vp, _vp, _vp, _int, _vp]),
    "pb_replay_collate": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp,
                                 _vp, _vp, _int, _vp]),
    "pb_replay_slot_obs": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _int, _vp]),
    "pb_eval_track": (_int, [_vp, _vp, _vp, _vp]),
    "pb_bias_pool_silu_fwd": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _vp, _int, _vp]),
    "pb_bias_pool_silu_bwd": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _int, _vp]),
    "pb_trace_begin": (_int, [_vp, _i64]),
    "pb_trace_mark": (_int, [_vp, _vp]),
    "pb_trace_end": (_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "pb_gather_obs": (_int, [_vp, _i64, _vp, _vp, _i64, _vp, _int, _vp]),
    "pb_copy_envs": (_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "pb_mcts_create": (_int, [C.POINTER(_vp), _vp, _vp, _int]),
    "pb_mcts_destroy": (_int, [_vp]),
    "pb_sim_seed": (_u64, [_u64, _u64]),
    "pb_mcts_max_nodes": (_int, [_vp]),
    "pb_mcts_reset_root": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pb_mcts_root_noise": (_int, [_vp, _vp, _vp, C.c_float, _vp]),
    "pb_mcts_root_noise_dirichlet": (_int, [_vp, _vp, C.c_float, C.c_float, _vp]),
    "pb_mcts_select": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pb_mcts_backprop": (_int, [_vp, _vp, _vp, _vp]),
    "pb_mcts_take_action": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pb_mcts_collect_act": (_int, [_vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pb_mcts_backprop_collect": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 C.c_float, C.c_float, _int, _vp]),
    "pb_mcts_step_word": (_int, [_vp, C.POINTER(_vp)]),
    "pb_mcts_query": (_int, [_vp, _int, _vp, _vp]),
    "pb_mcts_check": (_int, [_vp, C.POINTER(_u64), _vp]),
}

MCTS_Q_BEST_ACTION, MCTS_Q_ACTION_DISTRIBUTION, MCTS_Q_POLICY_PRIOR, MCTS_Q_ACTION_VALUES = 0, 1, 2, 3
MCTS_Q_VALUE, MCTS_Q_MAX_DEPTH, MCTS_Q_NODE_COUNT, MCTS_Q_ROOT_VISITS = 4, 5, 6, 7

_lib = None


class PacbotB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libpacbot_b200.so (built by pacbot-rl_b200/build.py).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PacbotB200Error(
            f"CUDA extension not built: {LIB_PATH} is missing. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "This engine has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.pb_sizeof_env_state() != ENV_STATE_DTYPE.itemsize:
        raise PacbotB200Error("pb_env_state layout mismatch between the .so and _lib.py")
    _lib = L
    return L


def raw_stream(device) -> C.c_void_p:
    """cudaStream_t of torch's current stream on `device` (a torch.device with an index) as the
    `void *stream` argument of the C ABI.  The private raw accessor is ~10x cheaper than building
    a torch.cuda.Stream object per call and returns the same handle."""
    import torch

    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(device.index))
    except (AttributeError, TypeError):
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(rc: int) -> None:
    if rc == PB_OK:
        return
    msg = load().pb_last_error().decode("utf-8", "replace")
    if rc == PB_ERR_INVALID_ACTION:
        raise ValueError("Invalid action")  # what pyo3 raises, env.rs:83-88
    if rc == PB_ERR_INVALID_ARG:
        raise ValueError(msg)
    raise PacbotB200Error(f"pacbot_b200 error {rc}: {msg}")
