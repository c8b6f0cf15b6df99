"""Where does a collector step (one MCTS search iteration of every tree + due env steps) spend its
GPU time?  torch.profiler kernel table for the eager step, wall time per step eager vs graph.
Run on the GPU box: python tools/profile_mcts.py [num_trees]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pacbot_rl_b200 import alphazero, models

dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = models.NetV2((17, 28, 31), 6).to(dev).use_fused_blocks().eval()

@torch.no_grad()
def evaluator(obs, mask):
    return model.policy_value(obs, mask, 100.0)

def make():
    return alphazero.BatchedExperienceCollector(evaluator, alphazero.AlphaZeroConfig(100, 1000, 0.99, P), device=dev,
                                                leaf_layout="nhwc32")

def wall(col, steps=300):
    fin = []
    for _ in range(120):
        col.generate_experience_step(fin)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        col.generate_experience_step(fin)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

if "--graph-only" in sys.argv:   # A/B runs (PB_MCTS_TREES_PER_WARP, PB_MCTS_POOL_FACTOR)
    colg = make(); colg.enable_cuda_graph()
    print(f"{P} trees, tpw={os.environ.get('PB_MCTS_TREES_PER_WARP', 'auto')}: graph step {wall(colg):.3f} ms")
    sys.exit(0)
col = make()
print(f"{P} trees, eager step: {wall(col):.3f} ms")
fin = []
if "--plain" in sys.argv:   # for ncu launch lists: no CUPTI client inside the process
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(50):
        col.generate_experience_step(fin)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
rows = [e for e in prof.key_averages() if e.self_device_time_total > 0 and e.device_type == DeviceType.CUDA]
rows.sort(key=lambda e: -e.self_device_time_total)
total = sum(e.self_device_time_total for e in rows)
print(f"GPU time per collector step: {total / 50:.1f} us ({len(rows)} kernel rows)")
for e in rows[:40]:
    print(f"{e.self_device_time_total / 50:10.1f} us {e.count / 50:6.1f} calls  {e.key[:120]}")
colg = make(); colg.enable_cuda_graph()
print(f"{P} trees, graph step: {wall(colg):.3f} ms")
# the search iteration alone (no env steps, no host read)
m = colg.mcts
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(300):
    m._graph.replay() if m._graph is not None else m.search_iteration()
torch.cuda.synchronize()
print(f"search iteration only (eager launches): {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms")
