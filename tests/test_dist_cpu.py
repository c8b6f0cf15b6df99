"""CPU tests of the N>1 host logic with the gloo backend, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pacbot_rl_b200 as pb
    from pacbot_rl_b200 import dist_utils, models

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert dist_utils.rank_world() == (rank, rank, world)
        # gradient averaging: every rank ends with the mean of the per-rank gradients
        torch.manual_seed(0)
        net = models.QNetV2(pb.OBS_SHAPE, 5)
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        dist_utils.allreduce_mean_grads(list(net.parameters()), world)
        for i, p in enumerate(net.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))
        # max-over-ranks timing reduction
        assert dist_utils.max_over_ranks(10.0 + rank, torch.device("cpu"), world) == 11.0
        # env shards tile the global id range; per-env RNG streams are keyed by global id
        first, count = dist_utils.shard_env_ids(1001, rank, world)
        ranges = [None] * world
        dist.all_gather_object(ranges, (first, count))
        assert ranges[0] == (0, 501) and ranges[1] == (501, 500)
        L = pb._lib.load()
        mine = [L.pb_draw_raw(7, first + i, 3) for i in range(4)]
        single = [L.pb_draw_raw(7, g, 3) for g in range(first, first + 4)]
        assert mine == single
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gloo_world2_host_logic():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_env_ids_properties():
    import pacbot_rl_b200  # noqa: F401
    from pacbot_rl_b200.dist_utils import shard_env_ids

    for total in (1, 7, 64, 65536, 16 * 2**20 + 5):
        for world in (1, 2, 4, 8):
            spans = [shard_env_ids(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_env_ids(10, 2, 2)


def test_qnet_v2_known_answers():
    """K8: 156 934 parameters, candle-compatible key names, Q == 0 at init."""
    import pacbot_rl_b200 as pb
    from pacbot_rl_b200.models import QNetV2

    net = QNetV2(pb.OBS_SHAPE, 5)
    assert sum(p.numel() for p in net.parameters()) == 156_934
    keys = set(net.state_dict())
    expect = {f"backbone.{i}.{w}" for i in (0, 2, 5, 8, 11, 15) for w in ("weight", "bias")}
    expect |= {f"{h}.0.{w}" for h in ("value_head", "advantage_head") for w in ("weight", "bias")}
    assert keys == expect
    x = torch.randn(3, *pb.OBS_SHAPE)
    q = net(x)
    assert q.shape == (3, 5) and torch.count_nonzero(q) == 0
    # spatial pyramid 28x31 -> 14x16 -> 7x8 -> 4x4 (ceil-mode pooling)
    h = x
    sizes = []
    for layer in net.backbone[:11]:
        h = layer(h)
        if isinstance(layer, torch.nn.MaxPool2d):
            sizes.append(tuple(h.shape[-2:]))
    assert sizes == [(14, 16), (7, 8), (4, 4)]
