"""Generate tests/golden/trajectory_v1.npz: a seeded 64-env x 240-step rollout of the CPU oracle
(actions, injected draw tape, reward, done, mask, digests of the full state and of the f32
observation every step).  It pins the ORACLE's behaviour over time (any edit to oracle/ that
changes a trajectory shows up in the CPU suite) and gives the GPU suite a committed fixture to
reproduce bit for bit.  It does NOT pin the oracle to the reference (nothing can, see DESIGN.md
section 3).  Usage: python tools/make_golden_trajectory.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

N, STEPS, TAPE = 64, 240, 127
DIGEST_FIELDS = ["curr_ticks", "rng_ctr", "pellets", "level_steps", "curr_score", "num_pellets",
                 "last_score", "update_period", "mode", "paused", "pause_on_update", "mode_steps",
                 "curr_level", "curr_lives", "ghost_combo", "fruit_steps", "last_action",
                 "ticks_per_step", "pac_row", "pac_col", "pac_dir", "g_row", "g_col", "g_dir",
                 "g_next_row", "g_next_col", "g_next_dir", "g_fright", "g_trapped", "g_spawning",
                 "g_eaten"]


def state_digest(st) -> np.ndarray:
    """8-byte digest per step over the listed fields of all envs (layout independent)."""
    h = hashlib.blake2b(digest_size=8)
    for f in DIGEST_FIELDS:
        h.update(np.ascontiguousarray(st[f]).astype(np.int64).tobytes())
    return np.frombuffer(h.digest(), np.uint64)[0]


def obs_digest(obs) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(obs).view(np.uint32).tobytes(),
                                         digest_size=8).digest(), np.uint64)[0]


def rollout(batch_factory):
    rng = np.random.default_rng(20261020)
    tape = rng.integers(0, 2**32, size=(N, TAPE), dtype=np.uint32)
    cfg = (np.arange(N) % 4).astype(np.uint8)
    b = batch_factory(N, cfg, tape)
    b.reset()
    actions = np.zeros((STEPS, N), np.uint8)
    reward = np.zeros((STEPS, N), np.int32)
    done = np.zeros((STEPS, N), bool)
    masks = np.zeros((STEPS, N, 5), bool)
    sdig = np.zeros(STEPS, np.uint64)
    odig = np.zeros(STEPS, np.uint64)
    for t in range(STEPS):
        m = b.action_mask()
        masks[t] = m
        u = rng.random((N, 5)) * m
        a = u.argmax(1).astype(np.uint8)
        wild = rng.random(N) < 0.1
        a[wild] = rng.integers(0, 5, int(wild.sum()), dtype=np.uint8)
        actions[t] = a
        reward[t], done[t] = b.step(a, auto_reset=True)
        sdig[t] = state_digest(b.export_state())
        odig[t] = obs_digest(b.obs())
    return dict(tape=tape, cfg=cfg, actions=actions, reward=reward, done=done, masks=masks,
                state_digest=sdig, obs_digest=odig)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "trajectory_v1.npz")
    data = rollout(pyoracle.OracleBatch)
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes; episodes finished:", int(data["done"].sum()),
          "reward sum:", int(data["reward"].sum()))
