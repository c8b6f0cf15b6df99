"""`PacmanGym`: drop-in for the reference's pyo3 class of the same name
(pacbot_rs/src/env.rs:130-359; stub pacbot_rs/pacbot_rs.pyi:6-19), one env per object, so that
`src/replay_buffer.py`, `src/algorithms/train_dqn.py` and the duck-typed env protocol of
`src/debug_probe_envs.py` keep working unchanged.  Each object owns a 1-env slice of the CUDA
engine; `PacmanGymView` exposes the same read-only surface for env i of a `BatchedPacmanGym`."""
from __future__ import annotations

import itertools
import os

import numpy as np

from ._lib import CFG_RANDOM_START, OBS_SHAPE
from .engine import BatchedPacmanGym

_WALL_ROWS = None

# Defaults for objects built the reference's way, `PacmanGym(random_start, random_ticks)`: the
# reference gives every env its own `rand::thread_rng()` stream (env.rs:146), so two fresh objects
# must not replay the same random ticks / start node / frightened-ghost turns.  Unless the caller
# passes `seed` / `env_id` (reproducible tests), every object gets the next env id of a
# process-wide counter under a seed drawn once per process from the OS.
_PROCESS_SEED = int.from_bytes(os.urandom(8), "little")
_ENV_IDS = itertools.count()


def _wall_at(row: int, col: int) -> bool:
    global _WALL_ROWS
    if _WALL_ROWS is None:
        from .maze import WALL_ROWS

        _WALL_ROWS = WALL_ROWS
    if not (0 <= row < 31 and 0 <= col < 28):
        return True
    return bool((_WALL_ROWS[row] >> col) & 1)


class _SingleEnvSurface:
    """Read-only per-env API shared by PacmanGym and PacmanGymView."""

    _batch: BatchedPacmanGym
    _index: int

    def _state(self):
        return self._batch.get_state(self._index, 1)[0]

    def score(self) -> int:  # env.rs:269-271
        return int(self._state()["curr_score"])

    def lives(self) -> int:  # env.rs:273-275
        return int(self._state()["curr_lives"])

    def is_done(self) -> bool:  # env.rs:277-279
        s = self._state()
        return bool(s["curr_lives"] < 3 or s["curr_level"] != 1)

    def first_ai_done(self) -> bool:  # env.rs:281-285
        s = self._state()
        supers = any((int(s["pellets"][r]) >> c) & 1 for r, c in ((3, 1), (3, 26), (23, 1), (23, 26)))
        return (not supers) and not bool((s["g_fright"] > 0).any())

    def remaining_pellets(self) -> int:  # env.rs:287-289
        return int(self._state()["num_pellets"])

    def action_mask(self) -> list[bool]:  # env.rs:293-302
        m = self._batch.action_mask()[self._index]
        return [bool(x) for x in m.tolist()]

    def obs_numpy(self) -> np.ndarray:  # env.rs:305-307: fresh (17, 28, 31) float32 array
        return self._batch.obs_numpy(self._index, 1)[0]

    def print_game_state(self) -> None:  # env.rs:310-358
        s = self._state()
        line = f"Score: {int(s['curr_score'])}"
        if self.is_done():
            line += "  [DONE]"
        print(line)
        ghosts = [(int(s["g_row"][i]), int(s["g_col"][i]), bool(s["g_fright"][i] > 0)) for i in range(4)]
        out = []
        for row in range(31):
            cells = []
            for col in range(28):
                ch, style = " ", ""
                if (row, col) == (int(s["pac_row"]), int(s["pac_col"])):
                    ch, style = "@", "93"
                else:
                    for i, name in enumerate("RPBO"):
                        if (row, col) == ghosts[i][:2]:
                            ch, style = name, ("96" if ghosts[i][2] else "38;5;206")
                            break
                    else:
                        if _wall_at(row, col):
                            ch, style = "#", "90"
                        elif (row, col) == (17, 13) and s["fruit_steps"] > 0:
                            ch, style = "c", "31"
                        elif (int(s["pellets"][row]) >> col) & 1:
                            ch = "o" if row in (3, 23) and col in (1, 26) else "."
                cells.append(f"\x1b[{style}m{ch}\x1b[0m")
            out.append("".join(cells))
        print("\n".join(out))


class PacmanGym(_SingleEnvSurface):
    """pacbot_rs.PacmanGym(random_start, random_ticks) on the B200 engine."""

    def __init__(self, random_start: bool, random_ticks: bool, *, device=None, seed=None,
                 env_id=None) -> None:
        if seed is None:
            seed = _PROCESS_SEED
        if env_id is None:
            env_id = next(_ENV_IDS)
        self._batch = BatchedPacmanGym(
            1, bool(random_start), bool(random_ticks), device=device, seed=int(seed),
            env_id_base=int(env_id))
        self._index = 0
        self._snap = None

    @property
    def random_start(self) -> bool:  # #[pyo3(get, set)] env.rs:100-101
        return bool(self._batch._cfg[0] & CFG_RANDOM_START)

    @random_start.setter
    def random_start(self, value: bool) -> None:
        self._batch.random_start = bool(value)
        self._snap = None

    # ---- one device round trip per mutation --------------------------------------------------
    # The reference's loop asks an env several questions between two mutations (`is_done()`,
    # `obs_numpy()`, `action_mask()`, `score()` ...: src/replay_buffer.py:107-137,
    # src/algorithms/train_dqn.py:95-113).  `step()` brings everything back with the step's own
    # launch (pb_step_snapshot_host: k_gym_call steps, answers and encodes into pinned memory), and the
    # getters answer from that snapshot until the env changes again.
    def _snapshot(self) -> dict:
        if self._snap is None:
            self._snap = self._batch.step_snapshot(None, want_obs=True)
        return self._snap

    def invalidate(self) -> None:
        """Forget the host snapshot (call after changing the env through `_batch` directly)."""
        self._snap = None

    def reset(self) -> None:  # env.rs:140-212
        self._batch.reset()
        self._snap = None

    def step(self, action: int) -> tuple[int, bool]:  # env.rs:216-267
        if not isinstance(action, (int, np.integer)):
            raise TypeError(f"'{type(action).__name__}' object cannot be interpreted as an integer")
        if not 0 <= int(action) <= 255:
            raise OverflowError("out of range integral type conversion attempted")
        self._snap = None
        self._snap = snap = self._batch.step_snapshot((int(action),), want_obs=True)
        return int(snap["reward"][0]), bool(snap["done"][0])

    def score(self) -> int:  # env.rs:269-271
        return int(self._snapshot()["info"][0, 0])

    def lives(self) -> int:  # env.rs:273-275
        return int(self._snapshot()["info"][0, 1])

    def is_done(self) -> bool:  # env.rs:277-279
        return bool(self._snapshot()["info"][0, 2])

    def first_ai_done(self) -> bool:  # env.rs:281-285
        return bool(self._snapshot()["info"][0, 3])

    def remaining_pellets(self) -> int:  # env.rs:287-289
        return int(self._snapshot()["info"][0, 4])

    def action_mask(self) -> list[bool]:  # env.rs:293-302
        return self._snapshot()["mask"][0].tolist()

    def obs_numpy(self) -> np.ndarray:  # env.rs:305-307: fresh (17, 28, 31) float32 array
        return self._snapshot()["obs"][0].copy()

    def clone(self) -> "PacmanGym":  # #[derive(Clone)] env.rs:96
        other = PacmanGym(False, False, device=self._batch.device, seed=self._batch.seed,
                          env_id=self._batch.env_id_base)
        other._batch.clone_env_from(0, self._batch, 0)
        return other


class PacmanGymView(_SingleEnvSurface):
    """Read-only `PacmanGym` surface for env `index` of a BatchedPacmanGym."""

    def __init__(self, batch: BatchedPacmanGym, index: int) -> None:
        if not 0 <= index < batch.num_envs:
            raise IndexError(index)
        self._batch = batch
        self._index = int(index)

    @property
    def random_start(self) -> bool:
        return bool(self._batch._cfg[self._index] & CFG_RANDOM_START)


OBS_SHAPE = OBS_SHAPE
