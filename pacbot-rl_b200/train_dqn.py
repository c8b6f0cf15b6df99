"""Double-DQN trainer on GPU-resident environments: the device-side mirror of
/root/reference/src/algorithms/train_dqn.py (hyper-parameters and flag names :26-57, loop
:116-265, evaluation :88-113, checkpoint cadence :239-253).

What changed relative to the reference loop (and nothing else):
  * envs, replay ring, epsilon-greedy and the collate live on the GPU (no per-step host trip);
  * metrics are reduced on the device and read back every `--log_every` iterations instead of a
    `loss.item()` sync per iteration;
  * the phase-1 "old policy" of `reset_env` is optional (its checkpoint is not in the reference
    repo): `--old-policy PATH` enables it;
  * wandb is optional and disabled by default (no network on the GPU box); metrics also go to a
    JSONL file;
  * under torchrun the envs shard per rank and gradients are averaged with ONE NCCL all-reduce
    of the flat 156 934-float gradient per iteration.

Run:  python -m pacbot_rl_b200.train_dqn --num_iters 2000 [--device cuda:0]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import shutil
import time
from pathlib import Path
from typing import Optional

import torch
import torch.nn.functional as F

from . import models
from .engine import BatchedPacmanGym
from .policies import EpsilonGreedy, MaxQPolicy
from .replay_buffer import ReplayBuffer, reset_env

# names and defaults exactly as the reference (train_dqn.py:26-42)
hyperparam_defaults = {
    "learning_rate": 0.0001,
    "batch_size": 512,
    "num_iters": 2_000_000,
    "replay_buffer_size": 10_000,
    "num_parallel_envs": 32,
    "random_start_proportion": 0.5,
    "experience_steps": 4,
    "target_network_update_steps": 500,
    "evaluate_steps": 10,
    "initial_epsilon": 0.1,
    "final_epsilon": 0.1 * 0.05,
    "discount_factor": 0.99,
    "reward_scale": 1 / 50,
    "grad_clip_norm": 10_000,
    "model": "QNetV2",
}


def lerp(a: float, b: float, t: float) -> float:  # reference src/utils.py
    return (1 - t) * a + t * b


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--eval", metavar="CHECKPOINT", default=None)
    p.add_argument("--finetune", metavar="CHECKPOINT", default=None)
    p.add_argument("--no-wandb", action="store_true", default=True)
    p.add_argument("--wandb", dest="no_wandb", action="store_false")
    p.add_argument("--checkpoint-dir", default="checkpoints")
    p.add_argument("--device", default=None)
    p.add_argument("--old-policy", default=None, help="q_net-old.safetensors for the two-stage reset")
    p.add_argument("--log_every", type=int, default=50)
    p.add_argument("--metrics-file", default=None)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--eval_envs", type=int, default=1,
                   help="greedy evaluation episodes run in parallel (the reference runs 1)")
    p.add_argument("--no-checkpoints", action="store_true")
    p.add_argument("--no-channels-last", dest="channels_last", action="store_false", default=True,
                   help="keep the Q-network in NCHW (default: NHWC weights/activations for cuDNN)")
    p.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", default=True,
                   help="run the iteration eagerly instead of replaying one captured CUDA graph")
    for name, default in hyperparam_defaults.items():
        p.add_argument(f"--{name}", type=type(default), default=default, help="Default: %(default)s")
    return p


class DQNTrainer:
    """Owns the Q-networks, optimiser, replay buffer and the batched envs."""

    def __init__(self, cfg: argparse.Namespace, device: torch.device, rank: int = 0, world: int = 1):
        self.cfg = cfg
        self.device = device
        self.rank, self.world = rank, world
        torch.manual_seed(cfg.seed)  # identical initial weights on every rank
        model_cls = getattr(models, cfg.model)
        self.q_net = model_cls((17, 28, 31), 5).to(device)
        if cfg.finetune:
            self.q_net = torch.load(cfg.finetune, map_location=device, weights_only=False)
        if cfg.channels_last and hasattr(self.q_net, "use_channels_last"):
            self.q_net.use_channels_last()
        torch.manual_seed(cfg.seed * 1000003 + rank)
        self.target_q_net = copy.deepcopy(self.q_net).eval()
        phase1 = None
        if cfg.old_policy:
            import safetensors.torch

            old = model_cls((17, 28, 31), 5).to(device)
            old.load_state_dict(safetensors.torch.load_file(cfg.old_policy))
            phase1 = MaxQPolicy(old.eval())
        # epsilon lives in a 0-dim device tensor so that changing it needs no re-capture
        self._eps_t = torch.tensor(cfg.initial_epsilon if cfg.finetune else 1.0, device=device)
        self._eps_host = float(self._eps_t)
        self._graph = None
        self.policy = EpsilonGreedy(MaxQPolicy(self.q_net), 5, self._eps_t)
        self.replay = ReplayBuffer(
            maxlen=cfg.replay_buffer_size, policy=self.policy, num_parallel_envs=cfg.num_parallel_envs,
            random_start_proportion=cfg.random_start_proportion, device=device, seed=cfg.seed,
            env_id_base=rank * cfg.num_parallel_envs, phase1_policy=phase1)
        self.optimizer = torch.optim.Adam(self.q_net.parameters(), lr=cfg.learning_rate,
                                          capturable=True)
        self.params = [p for p in self.q_net.parameters()]
        self.iter_num = 0
        # streaming metrics (device side)
        self._acc = torch.zeros(5, device=device)
        self._acc_n = 0

    # ------------------------------------------------------------------ one iteration
    @property
    def epsilon(self) -> float:
        return self._eps_host

    @epsilon.setter
    def epsilon(self, value: float) -> None:
        self._eps_host = float(value)
        self._eps_t.fill_(float(value))

    @torch.no_grad()
    def update_target(self) -> None:
        torch._foreach_copy_(list(self.target_q_net.parameters()), list(self.q_net.parameters()))

    def _iteration_body(self) -> None:
        """Everything of one iteration that runs on the device: sample + collate, double-DQN
        target, loss, backward, (all-reduce,) clip, Adam, metrics, experience collection.  No
        host synchronisation and no host-dependent indexing, so it can be CUDA-graph captured."""
        cfg = self.cfg
        batch = self.replay.sample_batch(cfg.batch_size)
        with torch.no_grad():  # double DQN target (train_dqn.py:167-183)
            masks = batch.next_action_masks
            next_q = self.target_q_net(batch.next_obs).masked_fill(~masks, -torch.inf)
            online_next_q = self.q_net(batch.next_obs).masked_fill(~masks, -torch.inf)
            next_actions = online_next_q.argmax(dim=1)
            returns = next_q.gather(1, next_actions.unsqueeze(1)).squeeze(1)
            discounted = torch.where(batch.done, torch.zeros_like(returns), cfg.discount_factor * returns)
            target_q = batch.rewards * cfg.reward_scale + discounted
        self.optimizer.zero_grad(set_to_none=False)
        all_q = self.q_net(batch.obs)
        predicted = all_q.gather(1, batch.actions.unsqueeze(1)).squeeze(1)
        loss = F.mse_loss(predicted, target_q)
        loss.backward()
        if self.world > 1:
            self._allreduce_grads()
        grad_norm = torch.nn.utils.clip_grad_norm_(self.params, max_norm=cfg.grad_clip_norm)
        self.optimizer.step()
        with torch.no_grad():
            self._acc += torch.stack([
                loss.detach(), grad_norm.detach(), all_q.detach().amax(dim=1).mean() / cfg.reward_scale,
                target_q.mean() / cfg.reward_scale, torch.isfinite(loss.detach()).float()])
        # collect experience (train_dqn.py:263-265); epsilon was set by the host wrapper
        for _ in range(cfg.experience_steps):
            self.replay.generate_experience_step(count_on_host=False)

    def capture_graph(self) -> None:
        """Capture `_iteration_body` into ONE CUDA graph (the reference iteration is hundreds of
        tiny launches + Python; here a replay is a single launch).  Needs >= 3 eager warm-up
        iterations on a side stream first (optimizer state, cuDNN autotune, allocator)."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._iteration_body()
                self.replay._t += self.cfg.experience_steps
                self._acc_n += 1
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._iteration_body()
        # the capture pass does not execute anything; nothing to account for on the host

    def train_iteration(self) -> None:
        cfg = self.cfg
        if self.iter_num % cfg.target_network_update_steps == 0:
            self.update_target()
        # anneal epsilon (train_dqn.py:256-260).  The reference assigns it after the update and
        # before the experience steps of the same iteration; the update does not read epsilon, so
        # assigning it before the (captured) body is equivalent.
        self.epsilon = lerp(cfg.initial_epsilon, cfg.final_epsilon,
                            self.iter_num / max(1, cfg.num_iters - 1))
        if self._graph is not None:
            self._graph.replay()
        else:
            self._iteration_body()
        self.replay._t += cfg.experience_steps
        self._acc_n += 1
        self.iter_num += 1

    def _allreduce_grads(self) -> None:
        from .dist_utils import allreduce_mean_grads

        allreduce_mean_grads(self.params, self.world)

    def pop_metrics(self) -> dict:
        vals = (self._acc / max(1, self._acc_n)).tolist()   # the only host sync of the loop
        n = self._acc_n
        self._acc.zero_()
        self._acc_n = 0
        if vals[4] < 1.0:
            raise FloatingPointError("non-finite loss")  # error_if_nonfinite tripwire (:200-204)
        return {"loss": vals[0], "grad_norm": vals[1], "avg_predicted_value": vals[2],
                "avg_target_q_value": vals[3], "exploration_epsilon": self.epsilon,
                "iters_averaged": n}

    # ------------------------------------------------------------------ evaluation (:88-113)
    @torch.no_grad()
    def evaluate_episode(self, max_steps: int = 1000, num_envs: int = 1) -> dict:
        """Greedy episode(s) from the fixed start (random_start=False, random_ticks=False), after
        the old policy's phase when there is one (`reset_env(gym)`, train_dqn.py:95-96)."""
        env = BatchedPacmanGym(num_envs, False, False, device=self.device, seed=self.cfg.seed + 17,
                               env_id_base=10_000_000 + self.iter_num * num_envs)
        reset_env(env, self.replay.phase1_policy)
        pellets_start = env.remaining_pellets()
        self.q_net.eval()
        policy = MaxQPolicy(self.q_net)
        obs = env.obs()
        alive = torch.ones(num_envs, dtype=torch.bool, device=self.device)
        steps = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        score = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        lives = torch.full((num_envs,), 3, dtype=torch.int32, device=self.device)
        pellets_end = pellets_start.clone()
        for step_num in range(1, max_steps + 1):
            actions = policy(obs, env.action_mask())
            _, done = env.step_obs(actions, obs)
            steps += alive.int()
            # freeze the statistics of finished envs (they keep stepping: harmless, Q12)
            now = alive & done
            score = torch.where(alive, env.score(), score)
            lives = torch.where(alive, env.lives(), lives)
            pellets_end = torch.where(alive, env.remaining_pellets(), pellets_end)
            alive = alive & ~now
            if step_num % 32 == 0 and not bool(alive.any()):
                break
        cleared = (~alive) & (lives == 3)
        pellets_end = torch.where(cleared, torch.zeros_like(pellets_end), pellets_end)
        return {
            "eval_episode_score": score.float().mean().item(),
            "eval_episode_steps": steps.float().mean().item(),
            "cleared_board": cleared.float().mean().item(),
            "eval_pellets_start": pellets_start.float().mean().item(),
            "eval_pellets_end": pellets_end.float().mean().item(),
        }

    # ------------------------------------------------------------------ checkpoints (:239-253)
    def save_checkpoint(self, checkpoint_dir: str, iter_num: int) -> None:
        d = Path(checkpoint_dir)
        d.mkdir(exist_ok=True, parents=True)
        pt = d / "q_net-latest.pt"
        # plain row-major tensors (what the reference writes and its to_safetensors.py / candle
        # loader expect), whatever memory format the live network uses
        net = models.portable_copy(self.q_net)
        torch.save(net, pt)
        shutil.copyfile(pt, d / f"q_net-iter{iter_num:07}.pt")
        if iter_num % 1_000 == 0:
            try:
                import safetensors.torch

                # the reference's key names (candle_qnet.rs:18-44)
                safetensors.torch.save_file(net.state_dict(), d / "q_net-latest.safetensors")
            except ImportError:
                pass


def train(cfg: argparse.Namespace, max_seconds: Optional[float] = None) -> DQNTrainer:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(cfg.device or f"cuda:{local_rank}")
    if device.type != "cuda":
        raise SystemExit("the B200 trainer needs a CUDA device (the env engine has no CPU path)")
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    wandb = None
    if not cfg.no_wandb and rank == 0:
        import wandb as _wandb

        wandb = _wandb
        wandb.init(project="pacbot-ind-study", tags=["DQN", "b200"],
                   config={k: getattr(cfg, k) for k in hyperparam_defaults})
    trainer = DQNTrainer(cfg, device, rank, world)
    print(f"q_net has {sum(p.numel() for p in trainer.q_net.parameters())} parameters", flush=True)
    trainer.replay.fill()
    trainer.epsilon = cfg.initial_epsilon
    if cfg.cuda_graph:
        trainer.capture_graph()
    metrics_f = open(cfg.metrics_file, "a") if (cfg.metrics_file and rank == 0) else None
    t0 = time.time()
    last_t, last_it = t0, 0
    for it in range(cfg.num_iters):
        trainer.train_iteration()
        if (it + 1) % cfg.log_every == 0 or it == cfg.num_iters - 1:
            m = trainer.pop_metrics()
            if (it + 1) % (cfg.log_every * cfg.evaluate_steps) == 0:
                m.update(trainer.evaluate_episode(num_envs=cfg.eval_envs))
            now = time.time()
            m["iter"] = it + 1
            m["iters_per_sec"] = (it + 1 - last_it) / max(1e-9, now - last_t)
            m["env_steps_per_sec"] = m["iters_per_sec"] * cfg.experience_steps * cfg.num_parallel_envs * world
            last_t, last_it = now, it + 1
            if rank == 0:
                print(json.dumps(m), flush=True)
                if metrics_f:
                    metrics_f.write(json.dumps(m) + "\n")
                    metrics_f.flush()
                if wandb:
                    wandb.log(m)
        if not cfg.no_checkpoints and rank == 0 and it % 500 == 0:
            trainer.save_checkpoint(cfg.checkpoint_dir, it)
        if max_seconds is not None and time.time() - t0 > max_seconds:
            break
    if metrics_f:
        metrics_f.close()
    return trainer


def main(argv=None) -> None:
    cfg = build_parser().parse_args(argv)
    if not hasattr(models, cfg.model):
        raise SystemExit(f"Invalid --model: {cfg.model!r} (must be a class in models.py)")
    if cfg.eval:
        device = torch.device(cfg.device or "cuda")
        trainer = DQNTrainer(cfg, device)
        trainer.q_net = torch.load(cfg.eval, map_location=device, weights_only=False)
        print(json.dumps(trainer.evaluate_episode(num_envs=cfg.eval_envs)))
        return
    train(cfg)


if __name__ == "__main__":
    main()
