"""Does padding the observation's 17 channels to a multiple of 8 (zeros, NHWC) get the first
convolution of QNetV2 onto a tensor-core kernel?  fwd / fwd+bwd at batch 512 and 32, B200."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
import pacbot_rl_b200 as pb
from pacbot_rl_b200.models import QNetV2

dev = torch.device("cuda:0")


def bench(fn, reps=50):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for b in (512, 32, 1):
    x = torch.rand(b, 17, 28, 31, device=dev).contiguous(memory_format=torch.channels_last)
    net = QNetV2(pb.OBS_SHAPE, 5).to(dev).use_channels_last()
    conv = net.backbone[0]
    for cpad in (17, 20, 24, 32):
        xp = F.pad(x, (0, 0, 0, 0, 0, cpad - 17)).contiguous(memory_format=torch.channels_last)

        def first_fwd():
            w = F.pad(conv.weight, (0, 0, 0, 0, 0, cpad - 17)).contiguous(memory_format=torch.channels_last)
            return F.conv2d(xp, w, conv.bias, padding="same")

        def full(train):
            w = F.pad(conv.weight, (0, 0, 0, 0, 0, cpad - 17)).contiguous(memory_format=torch.channels_last)
            h = F.conv2d(xp, w, conv.bias, padding="same")
            out = net.backbone[1:](h)
            q = net.value_head(out) + net.advantage_head(out)
            if train:
                net.zero_grad()
                q.sum().backward()
            return q

        with torch.no_grad():
            t1 = bench(first_fwd)
            t2 = bench(lambda: full(False))
        t3 = bench(lambda: full(True))
        with torch.no_grad():
            ref = net(x)
            got = full(False)
        # the reference forward adds the heads differently; compare the first conv only
        with torch.no_grad():
            d = (first_fwd() - F.conv2d(x, conv.weight, conv.bias, padding="same")).abs().max().item()
        print(f"batch {b:4d} C={cpad:2d}: conv1 fwd {t1*1e3:7.1f} us   net fwd {t2:6.3f} ms   net fwd+bwd {t3:6.3f} ms   "
              f"max |conv1 - conv1(C=17)| {d:.2e}", flush=True)
