"""pacbot-rl_b200: B200-native batched PacBot environment engine behind the reference's
`PacmanGym` API (qxzcode/pacbot-rl, pacbot_rs/src/env.rs).  CUDA only: importing works anywhere,
creating an engine needs the built extension and a CUDA device."""
from . import _lib, maze
from ._lib import (
    CFG_RANDOM_START,
    CFG_RANDOM_TICKS,
    CFG_RANDOMIZE_GHOSTS,
    CFG_REMOVE_SUPER_PELLETS,
    ENV_STATE_DTYPE,
    NUM_ACTIONS,
    OBS_FLOATS,
    OBS_SHAPE,
    PacbotB200Error,
)
from .engine import BatchedPacmanGym
from .gym import PacmanGym, PacmanGymView
from . import dist_utils, models, policies, replay_buffer
from . import mcts, alphazero

__all__ = [
    "BatchedPacmanGym",
    "PacmanGym",
    "PacmanGymView",
    "PacbotB200Error",
    "OBS_SHAPE",
    "OBS_FLOATS",
    "NUM_ACTIONS",
    "ENV_STATE_DTYPE",
    "CFG_RANDOM_START",
    "CFG_RANDOM_TICKS",
    "CFG_RANDOMIZE_GHOSTS",
    "CFG_REMOVE_SUPER_PELLETS",
    "maze",
]
