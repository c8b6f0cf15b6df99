import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import pacbot_rl_b200 as pb
from pacbot_rl_b200.rollout import HostActionRollout
dev = torch.device("cuda:0"); n = 65536
env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=0, auto_reset=True); env.reset()
ring = torch.empty((3, n) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
acts = [torch.randint(0, 5, (n,), dtype=torch.uint8).pin_memory() for _ in range(8)]
def run(variant, steps=1000):
    r = HostActionRollout(env)
    def step(i):
        if not r._started:
            cur = torch.cuda.current_stream(dev)
            for s in (r.step_stream, r.enc_stream, r.copy_stream): s.wait_stream(cur)
            r._started = True
        with torch.cuda.stream(r.step_stream):
            if "noh2d" not in variant: r._actions.copy_(acts[i % 8], non_blocking=True)
            r.step_stream.wait_event(r._enc_done[r._t & 1])
            env.step(r._actions, check_actions=False, mask_out=r._mask, reward_out=r._reward, done_out=r._done, flip=True)
            r._step_done.record(r.step_stream)
        with torch.cuda.stream(r.enc_stream):
            r.enc_stream.wait_event(r._step_done)
            if "tev" in variant: torch.cuda.Event(enable_timing=True).record(r.enc_stream)
            if "pev" in variant: torch.cuda.Event().record(r.enc_stream)
            env.obs(out=ring[i % 3])
            r._enc_done[r._t & 1].record(r.enc_stream)
        if "nod2h" not in variant:
            with torch.cuda.stream(r.copy_stream):
                r.copy_stream.wait_event(r._step_done)
                r.reward.copy_(r._reward, non_blocking=True); r.done.copy_(r._done, non_blocking=True); r.mask.copy_(r._mask, non_blocking=True)
        if "nosync" not in variant:
            st = (r.copy_stream if "nod2h" not in variant else r.step_stream)
            if "blockev" in variant:
                ev = torch.cuda.Event(blocking=True); ev.record(st); ev.synchronize()
            elif "sleep" in variant:
                ev = torch.cuda.Event(); ev.record(st)
                while not ev.query(): time.sleep(20e-6)
            else:
                st.synchronize()
            if "delay" in variant:
                t_end = time.perf_counter() + 150e-6
                while time.perf_counter() < t_end: pass
        r._t += 1
    for i in range(5): step(i)
    r.join(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps): step(i)
    r.join(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{variant:24s} {dt/steps*1e3:.4f} ms/step  {n*steps/dt/1e6:.2f} M/s", flush=True)
for v in ["nosync", "full", "full+delay", "nod2h+noh2d", "nod2h+noh2d+delay", "full"]:
    run(v)
