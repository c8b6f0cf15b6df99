"""Where does a DQN iteration spend its GPU time?  (torch.profiler kernel table + QNetV2
fwd/bwd micro-timings under a few cuDNN settings).  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pacbot_rl_b200 as pb
from pacbot_rl_b200 import train_dqn
from pacbot_rl_b200.models import QNetV2

dev = torch.device("cuda:0")

def time_fwd_bwd(net, x, reps=30):
    for _ in range(5):
        net.zero_grad(); net(x).sum().backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        net.zero_grad(); net(x).sum().backward()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

def time_fwd(net, x, reps=30):
    with torch.no_grad():
        for _ in range(5): net(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): net(x)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

x = torch.rand(512, 17, 28, 31, device=dev)
for name, bench, cl in [("default", False, False), ("cudnn.benchmark", True, False), ("benchmark+channels_last", True, True)]:
    torch.backends.cudnn.benchmark = bench
    net = QNetV2(pb.OBS_SHAPE, 5).to(dev)
    xx = x
    if cl:
        net = net.to(memory_format=torch.channels_last); xx = x.contiguous(memory_format=torch.channels_last)
    print(f"{name:28s} fwd {time_fwd(net, xx):.3f} ms  fwd+bwd {time_fwd_bwd(net, xx):.3f} ms", flush=True)

torch.backends.cudnn.benchmark = False
cfg = train_dqn.build_parser().parse_args(["--device", "cuda:0", "--no-checkpoints", "--log_every", "100000"])
tr = train_dqn.DQNTrainer(cfg, dev)
tr.replay.fill(); tr.epsilon = 0.1
for _ in range(5): tr.train_iteration()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): tr.train_iteration()
    torch.cuda.synchronize()
from torch.autograd import DeviceType
rows = [e for e in prof.key_averages() if e.self_device_time_total > 0 and e.device_type == DeviceType.CUDA]
rows.sort(key=lambda e: -e.self_device_time_total)
total = sum(e.self_device_time_total for e in rows)
print(f"GPU time per iteration: {total / 10 / 1e3:.3f} ms ({len(rows)} kernel rows)")
print(f"{'self device us / iter':>22s} {'%':>6s} {'calls / iter':>13s}  name")
for e in rows[:90]:
    print(f"{e.self_device_time_total / 10:22.1f} {100 * e.self_device_time_total / total:6.2f} {e.count / 10:13.1f}  {e.key[:110]}")
