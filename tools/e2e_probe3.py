"""Kernel-side timeline (globaltimer stamps written by an instrumented build, -DPB_TRACE style
patch) of the pipelined step with and without a host sync per step."""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import pacbot_rl_b200 as pb
from pacbot_rl_b200 import _lib
L = _lib.load()
L.pb_debug_set_trace.argtypes = [C.c_void_p]; L.pb_debug_trace_count.restype = C.c_uint
dev = torch.device("cuda:0"); n = 65536
env = pb.BatchedPacmanGym(n, np.arange(n) < n // 2, True, device=dev, seed=0, auto_reset=True); env.reset()
ring = torch.empty((3, n) + pb.OBS_SHAPE, dtype=torch.float32, device=dev)
rew = torch.empty(n, dtype=torch.int32, device=dev); done = torch.empty(n, dtype=torch.bool, device=dev)
mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
acts = torch.randint(0, 5, (n,), dtype=torch.uint8).pin_memory().numpy()
outs = (np.empty(n, np.int32), np.empty(n, np.bool_), np.empty((n, 5), np.bool_))
buf = torch.zeros(60000, dtype=torch.int64, device=dev)
names = {1: "step start", 2: "step end", 3: "enc entry", 4: "enc past wait", 5: "enc 1st ticket", 6: "enc end"}
def run(sync, steps=40):
    for t in range(10):   # warm
        env.rollout_host_step(acts, ring[t % 3], outs) if sync else env.rollout_random_step(ring[t % 3], call_idx=t, mask_out=mask, reward_out=rew, done_out=done)
    env.rollout_join(); torch.cuda.synchronize()
    L.pb_debug_set_trace(C.c_void_p(buf.data_ptr()))
    for t in range(steps):
        env.rollout_host_step(acts, ring[t % 3], outs) if sync else env.rollout_random_step(ring[t % 3], call_idx=100 + t, mask_out=mask, reward_out=rew, done_out=done)
    env.rollout_join(); torch.cuda.synchronize()
    cnt = L.pb_debug_trace_count(); L.pb_debug_set_trace(None)
    raw = buf[:cnt].cpu().numpy().astype(np.uint64)
    tag = (raw >> np.uint64(56)).astype(int); ts = (raw & np.uint64((1 << 56) - 1)).astype(np.int64)
    order = np.argsort(ts); tag, ts = tag[order], ts[order]
    t0 = ts[0]
    print(f"--- sync={sync}: {cnt} stamps")
    ends = ts[tag == 6]
    rows = []
    for e0, e1 in zip(ends[8:28], ends[9:29]):
        m = (ts > e0) & (ts <= e1) & (tag >= 16)
        marks = dict(zip(tag[m] - 16, (ts[m] - e0) / 1e3))
        rows.append([marks.get(q, np.nan) for q in range(16)] + [(e1 - e0) / 1e3])
    rows = np.array(rows)
    print("mean us after previous enc end at which ticket q*4096 was taken (q=0..15), then enc end:")
    print(np.round(np.nanmean(rows, 0), 1))
    print("increments:", np.round(np.diff(np.nanmean(rows, 0)), 1))
for s in (False, True, False, True):
    run(s)
