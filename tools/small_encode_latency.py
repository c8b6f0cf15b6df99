"""GPU time of small observation-encode launches (k_obs_small vs the persistent k_encode_obs):
    python tools/small_encode_latency.py
    PB_SMALL_ENC_MAX=0 python tools/small_encode_latency.py      # persistent kernel everywhere
    PB_SMALL_CARVEOUT=0 python tools/small_encode_latency.py     # driver-chosen carve-out
CUDA events around 300 launches: the encoder alone back to back, and alternating with k_step the way
the DQN loop's experience step issues them (k_step runs under the largest shared-memory carve-out)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pacbot_rl_b200 as pb  # noqa: E402


def timed(fn, reps=300, warm=30):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps


def main():
    out = {"PB_SMALL_ENC_MAX": os.environ.get("PB_SMALL_ENC_MAX"), "PB_SMALL_CARVEOUT": os.environ.get("PB_SMALL_CARVEOUT")}
    for n in (1, 32, 128, 512, 1024):
        env = pb.BatchedPacmanGym(n, True, True, device="cuda:0", seed=1, auto_reset=True)
        env.reset()
        obs = torch.empty((n,) + pb.OBS_SHAPE, device="cuda:0")
        ctr = [0]
        reward = torch.empty(n, dtype=torch.int32, device="cuda:0")
        done = torch.empty(n, dtype=torch.uint8, device="cuda:0")

        def step():
            ctr[0] += 1
            env.step_random(call_idx=ctr[0], reward_out=reward, done_out=done)

        t_enc = timed(lambda: env.obs(out=obs))
        t_step = timed(step)

        def both():
            step()
            env.obs(out=obs)

        t_both = timed(both)
        out[f"n{n}"] = {"encode_us": round(t_enc, 2), "step_us": round(t_step, 2), "step_then_encode_us": round(t_both, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
