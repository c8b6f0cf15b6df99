"""ncu target for the small-batch kernels: six 32-env k_obs_small launches behind k_step, six
k_step_obs_small launches (32 envs, ring slot + channels-last output) and six k_gym_call launches
of a one-env drop-in.  Run under
    ncu --set full --clock-control none --import-source on \
        -k regex:"k_obs_small|k_step_obs_small|k_gym_call" -c 18 -o out python tools/ncu_small_kernels.py"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import pacbot_rl_b200 as pb
env = pb.BatchedPacmanGym(32, True, True, device="cuda:0", seed=1, auto_reset=True)
env.reset()
obs = torch.empty((32,) + pb.OBS_SHAPE, device="cuda:0")
for i in range(6):
    env.step_random(call_idx=i)
    env.obs(out=obs)
reward = torch.zeros(32, dtype=torch.int32, device="cuda:0")
done = torch.zeros(32, dtype=torch.bool, device="cuda:0")
mask = torch.zeros((32, 5), dtype=torch.bool, device="cuda:0")
ring = torch.zeros((3, 32) + pb.OBS_SHAPE, device="cuda:0")
nhwc = torch.empty((32, 32, 28, 31), device="cuda:0").contiguous(memory_format=torch.channels_last)
slot = torch.zeros(1, dtype=torch.int64, device="cuda:0")
for i in range(6):
    acts = env.sample_actions()
    env.step_encode_small(acts, reward_out=reward, done_out=done, mask_out=mask, ring=ring, slot_index=slot,
                          nhwc32_out=nhwc)
torch.cuda.synchronize()
g = pb.PacmanGym(True, True, device="cuda:0", seed=1, env_id=3)
g.reset()
for i in range(6):
    g.step(0)
