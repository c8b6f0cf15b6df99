"""ncu target for the small-batch kernels: six 32-env k_obs_small launches behind k_step and six
k_gym_call launches of a one-env drop-in.  Run under
    ncu --set full --clock-control none --import-source on -k regex:"k_obs_small|k_gym_call" -c 12 -o out python tools/ncu_small_kernels.py"""
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
torch.cuda.synchronize()
g = pb.PacmanGym(True, True, device="cuda:0", seed=1, env_id=3)
g.reset()
for i in range(6):
    g.step(0)
