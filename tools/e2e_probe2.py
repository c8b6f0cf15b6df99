"""Where does the host-synchronised (e2e) pipelined step lose 17 us against the free-running one?
Timing events around every step / encode kernel on their streams."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import pacbot_rl_b200 as pb
dev = torch.device("cuda:0"); n = 65536
env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=0, auto_reset=True); env.reset()
ring = torch.empty((3, n) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
ss, es = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
rew = torch.empty(n, dtype=torch.int32, device=dev); done = torch.empty(n, dtype=torch.bool, device=dev)
mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
def run(sync, steps=300):
    enc_done = [torch.cuda.Event(), torch.cuda.Event()]; step_done = torch.cuda.Event()
    T = lambda: torch.cuda.Event(enable_timing=True)
    base = T(); torch.cuda.synchronize(); base.record(); ss.wait_stream(torch.cuda.current_stream()); es.wait_stream(torch.cuda.current_stream())
    ev = []
    for t in range(steps):
        s0, s1, e0, e1 = T(), T(), T(), T()
        with torch.cuda.stream(ss):
            ss.wait_event(enc_done[t & 1]); s0.record(ss)
            env.step_random(call_idx=t, mask_out=mask, reward_out=rew, done_out=done, flip=True)
            s1.record(ss); step_done.record(ss)
        with torch.cuda.stream(es):
            es.wait_event(step_done); e0.record(es)
            env.obs(out=ring[t % 3])
            e1.record(es); enc_done[t & 1].record(es)
        if sync: ss.synchronize()
        ev.append((s0, s1, e0, e1))
    torch.cuda.synchronize()
    tt = np.array([[base.elapsed_time(x) * 1e3 for x in row] for row in ev])  # us
    s0, s1, e0, e1 = tt.T
    k = slice(50, steps)
    print(f"sync={sync}: period {np.diff(e0)[k].mean():7.1f} us | enc {(e1-e0)[k].mean():7.1f} | step {(s1-s0)[k].mean():6.1f} | "
          f"gap enc->next enc {(e0[1:]-e1[:-1])[k].mean():6.1f} | step start after enc(t-2) end {(s0[2:]-e1[:-2])[k].mean():7.1f} | "
          f"step end before enc(t-1) end {(e1[1:-1]-s1[2:])[k].mean():7.1f}")
for s in (False, True, False, True):
    run(s)
