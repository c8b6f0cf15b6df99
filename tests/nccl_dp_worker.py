"""torchrun worker of tests/test_dp_nccl_gpu.py (one process per GPU, NCCL): the data-parallel DQN
path of BASELINE.json configs[4].

  1. every rank builds `DQNTrainer(world=N)`: identical initial weights, its own env shard;
  2. one batch per rank, backward WITHOUT the collective -> local flat gradients, all-gathered;
  3. `allreduce_mean_grads` (the ONE NCCL all-reduce of the flat gradient, dist_utils.py) -> every
     rank must hold the mean of the per-rank gradients (bit-exact for 2 ranks: a + b commutes);
  4. 30 `train_iteration`s with the collective INSIDE the captured CUDA graph -> the weights must
     stay identical on every rank (they drift apart immediately if a replay skipped the collective)
     and differ from the initial ones."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler

    faulthandler.dump_traceback_later(240, exit=True)   # a hung collective must not hold the box
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from pacbot_rl_b200 import train_dqn

    cfg = train_dqn.build_parser().parse_args(
        ["--device", str(dev), "--seed", "5", "--no-checkpoints", "--batch_size", "128",
         "--replay_buffer_size", "2048", "--num_iters", "1000", "--no-eval-episodes"])
    tr = train_dqn.DQNTrainer(cfg, dev, rank, world)
    tr.replay.fill(cfg.batch_size)
    tr.epsilon = cfg.initial_epsilon

    def flat(ts):
        return torch.cat([t.detach().reshape(-1) for t in ts])

    def gathered(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    w0 = flat(tr.params).clone()
    for w in gathered(w0):
        assert torch.equal(w, w0), "initial weights differ between ranks"
    # the env shards differ: every rank owns its own global env-id range
    assert tr.replay.envs.env_id_base == rank * cfg.num_parallel_envs
    # on the trainer's side stream, like its own eager warm-up iterations: a backward on the legacy
    # default stream would tie the parameters' gradient accumulation to a stream that a later
    # graph capture may not depend on
    side = tr._side
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        batch = tr.replay.sample_batch(cfg.batch_size)
        target = train_dqn.dqn_targets(batch, tr.q_net, tr.target_q_net, cfg.discount_factor, cfg.reward_scale)
        tr.optimizer.zero_grad(set_to_none=False)
        loss, _ = train_dqn.dqn_loss(batch, tr.q_net, target)
        loss.backward()
        local_g = flat([p.grad for p in tr.params]).clone()
        every = gathered(local_g)
        if world > 1:
            assert not torch.equal(every[0], every[1]), "ranks produced identical gradients: envs not sharded"
        mean = torch.stack(every).sum(0) / world
        tr._allreduce_grads()
        got = flat([p.grad for p in tr.params])
        if world == 2:
            assert torch.equal(got, mean), float((got - mean).abs().max())
        else:
            torch.testing.assert_close(got, mean, rtol=1e-5, atol=1e-7)
    torch.cuda.current_stream(dev).wait_stream(side)
    # the collective inside the captured graph
    tr.capture_graph()
    for _ in range(30):
        tr.train_iteration()
    assert tr._graph is not None
    torch.cuda.synchronize(dev)
    w1 = flat(tr.params).clone()
    assert not torch.equal(w1, w0)
    for r, w in enumerate(gathered(w1)):
        assert torch.equal(w, w1), f"weights of rank {r} drifted from rank {rank}'s"
    m = tr.pop_metrics()
    assert m["loss"] == m["loss"]   # finite (pop_metrics raises on a non-finite loss)
    dist.barrier()
    torch.cuda.synchronize(dev)
    if rank == 0:
        print(f"DP_NCCL_OK world={world} loss={m['loss']:.6f}", flush=True)
    # leave without tearing the communicator down: destroy_process_group() waits forever while a
    # live CUDA graph still holds the captured NCCL kernels (observed: the run sat in it until the
    # caller's timeout)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
