"""Write-only / copy HBM bandwidth of this GPU with stock torch ops (context for the roofline:
the encoder is a pure-write stream, MEASURED_PEAKS.json's figure is a copy)."""
import json
import torch

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

n = 65536 * 14756  # floats in one 65 536-env observation slot (3.87 GB)
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
out = {}
out["fill_gbs"] = 4 * n / timeit(lambda: x.fill_(1.0)) / 1e6
out["zero_gbs"] = 4 * n / timeit(lambda: x.zero_()) / 1e6
out["copy_gbs_rw"] = 8 * n / timeit(lambda: y.copy_(x)) / 1e6
print(json.dumps(out))
