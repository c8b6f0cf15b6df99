"""Test helpers for the experience-generation semantics of the reference
(/root/reference/src/replay_buffer.py:44-54 `reset_env`, :102-139 `generate_experience_step`).

  * `restated_env_loop`  -- the reference's per-env behaviour restated on the scalar oracle: the
    expectation the GPU test of the batched ReplayBuffer compares against
    (tests/test_dqn_gpu.py::test_two_stage_reset_matches_reference_loop).
  * `bfs_policy`         -- a deterministic policy that reads nothing but exactly representable
    observation cells (walls, Pac-Man, pellet / super-pellet maps), so CPU and GPU agree on it.

tests/test_reference_replay_live.py runs the reference's OWN replay_buffer.py (unmodified, in a
subprocess, over an oracle-backed `pacbot_rs.PacmanGym` shim) and checks `restated_env_loop`
against it: the restatement is pinned to the live reference code, the GPU path to the restatement."""
from collections import deque

import numpy as np

STAY, DOWN, UP, LEFT, RIGHT = 0, 1, 2, 3, 4
# observation coordinates: x = col, y = 30 - row (env.rs:124,425) => Down is y - 1, Up is y + 1
MOVES = ((DOWN, 0, -1), (UP, 0, 1), (LEFT, -1, 0), (RIGHT, 1, 0))


def bfs_action(obs: np.ndarray, mask: np.ndarray, target_channels) -> int:
    """First move of a shortest path from Pac-Man (ch 2) to the nearest cell that is non-zero in one
    of `target_channels` (tried in order); Stay if there is none.  Ties: BFS order over MOVES."""
    pac = obs[2]
    if not pac.any():
        return STAY
    px, py = np.unravel_index(int(pac.argmax()), pac.shape)
    walls = obs[0] > 0
    for ch in target_channels:
        tgt = obs[ch] > 0
        tgt[px, py] = False
        if not tgt.any():
            continue
        seen = np.zeros_like(walls)
        seen[px, py] = True
        q = deque()
        for a, dx, dy in MOVES:
            x, y = px + dx, py + dy
            if mask[a] and 0 <= x < 28 and 0 <= y < 31 and not walls[x, y]:
                seen[x, y] = True
                q.append((x, y, a))
        while q:
            x, y, first = q.popleft()
            if tgt[x, y]:
                return first
            for _, dx, dy in MOVES:
                u, v = x + dx, y + dy
                if 0 <= u < 28 and 0 <= v < 31 and not walls[u, v] and not seen[u, v]:
                    seen[u, v] = True
                    q.append((u, v, first))
    return STAY


def bfs_policy(target_channels):
    """(obs [B,17,28,31], masks [B,5]) numpy -> actions uint8 [B]"""

    def policy(obs, masks):
        obs, masks = np.asarray(obs), np.asarray(masks)
        return np.array([bfs_action(o, m, target_channels) for o, m in zip(obs, masks)], np.uint8)

    return policy


# the "old" policy hunts super pellets (ch 16) and falls back to ordinary pellets / fruit (ch 1);
# the main policy eats whatever is nearest
OLD_POLICY_CHANNELS = (16, 1)
MAIN_POLICY_CHANNELS = (1,)


def restated_env_loop(ob, main, old, max_env_steps, max_recorded=None, in_phase1=None):
    """One env (an OracleBatch of size 1, already in its start state) played the way the reference
    plays it: while `in_phase1` the OLD policy acts, unrecorded, until first_ai_done(), re-resetting
    on death (reset_env, replay_buffer.py:44-54); otherwise the main policy acts and the transition
    is recorded (generate_experience_step, :102-139), a terminal one being followed by reset_env.
    Returns a list of (obs_bits uint32 [1,17,28,31], action, reward, done, next_mask [5])."""
    rec = []
    if in_phase1 is None:
        in_phase1 = not ob.first_ai_done()[0]
    for _ in range(max_env_steps):
        obs, mask = ob.obs(), ob.action_mask()
        if in_phase1:
            a = np.asarray(old(obs, mask)).astype(np.uint8)
            ob.step(a, auto_reset=True)
            in_phase1 = not ob.first_ai_done()[0]
        else:
            a = np.asarray(main(obs, mask)).astype(np.uint8)
            r, d = ob.step(a, auto_reset=True)
            # next_action_mask is [False] * 5 for a finished episode (replay_buffer.py:115-117)
            nm = np.zeros(5, bool) if d[0] else ob.action_mask()[0]
            rec.append((obs.view(np.uint32).copy(), int(a[0]), int(r[0]), bool(d[0]), nm))
            if d[0]:
                in_phase1 = not ob.first_ai_done()[0]
            if max_recorded is not None and len(rec) >= max_recorded:
                break
    return rec
