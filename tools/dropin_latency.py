"""Where the microseconds of one per-env drop-in step go (pacbot_rl_b200.PacmanGym, one env per
object, driven as src/replay_buffer.py:107-137 does).  Run on a GPU box:
    python tools/dropin_latency.py
Prints wall-clock microseconds per call for the layers of the path: the raw C-ABI call
(pb_step_snapshot_host = one k_gym_call launch + the wait for its sequence word), the engine
method around it, the gym method around that, and the whole reference call pattern."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pacbot_rl_b200 as pb  # noqa: E402


def per_call(fn, k=3000, warm=200):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    return 1e6 * (time.perf_counter() - t0) / k


def main():
    gym = pb.PacmanGym(True, True, device="cuda:0", seed=1, env_id=5)
    gym.reset()
    b = gym._batch
    b.auto_reset = True
    b.step_snapshot((0,), want_obs=True)
    p, L, h = b._snap_ptrs, b._L, b._h
    stream = b._stream()
    out = {}
    out["raw_call_step_obs"] = per_call(lambda: L.pb_step_snapshot_host(
        h, p["actions"], 1, p["reward"], p["done"], p["info"], p["mask"], p["obs"], stream))
    out["raw_call_step_no_obs"] = per_call(lambda: L.pb_step_snapshot_host(
        h, p["actions"], 1, p["reward"], p["done"], p["info"], p["mask"], None, stream))
    out["raw_call_no_step_no_obs"] = per_call(lambda: L.pb_step_snapshot_host(
        h, None, 1, None, None, p["info"], p["mask"], None, stream))
    out["engine_step_snapshot"] = per_call(lambda: b.step_snapshot((0,), want_obs=True))
    out["stream_accessor"] = per_call(b._stream)
    b.auto_reset = False
    gym.reset()

    def gym_step():
        _r, done = gym.step(0)
        if done:
            gym.reset()

    out["gym_step"] = per_call(gym_step)
    out["gym_obs_numpy"] = per_call(gym.obs_numpy)
    out["gym_action_mask"] = per_call(gym.action_mask)
    out["gym_is_done"] = per_call(gym.is_done)
    rng = np.random.default_rng(0)

    def pattern():
        mask = gym.action_mask()
        valid = [a for a in range(5) if mask[a]]
        _r, done = gym.step(valid[int(rng.integers(len(valid)))])
        if gym.is_done() or done:
            gym.reset()
        gym.obs_numpy()
        gym.action_mask()

    out["reference_call_pattern"] = per_call(pattern)
    print(json.dumps({k: round(v, 2) for k, v in out.items()}))


if __name__ == "__main__":
    main()
