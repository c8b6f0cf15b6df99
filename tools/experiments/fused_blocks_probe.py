import copy, sys, os
sys.path.insert(0, "/root/repo")
import torch
import pacbot_rl_b200 as pb
from pacbot_rl_b200.models import QNetV2
dev = torch.device("cuda:0")
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.manual_seed(2)
    plain = QNetV2(pb.OBS_SHAPE, 5).to(dev).use_channels_last()
    with torch.no_grad():
        for p in plain.parameters():
            if p.dim() == 1: p.normal_(0, 0.3)
        for head in (plain.value_head, plain.advantage_head): head[0].weight.normal_(0, 0.3)
    fused = copy.deepcopy(plain).use_fused_blocks()
    plain2 = copy.deepcopy(plain)
    g = torch.Generator(device="cpu").manual_seed(3)
    obs = ((torch.rand((48,) + pb.OBS_SHAPE, generator=g) < 0.15).float() * torch.rand((48,) + pb.OBS_SHAPE, generator=g)).to(dev)
    for x in (obs, torch.nn.functional.pad(obs, (0, 0, 0, 0, 0, 15))):
        for net in (plain, fused, plain2):
            net.zero_grad(); (net(x) ** 2).sum().backward()
        print("tf32", tf32, "C", x.shape[1])
        for (name, pa), pf, p2 in zip(plain.named_parameters(), fused.parameters(), plain2.parameters()):
            s = float(pa.grad.abs().max())
            print(f"  {name:28s} scale {s:10.4f} fused-plain {float((pf.grad-pa.grad).abs().max())/s:.2e}  plain2-plain {float((p2.grad-pa.grad).abs().max())/s:.2e}")
