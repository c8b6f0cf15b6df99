"""AlphaZero-style trainer on GPU-resident environments and search trees: the device-side mirror
of /root/reference/src/algorithms/train_alphazero.py (hyper-parameters and flags :22-52,
model_evaluator :81-101, evaluate_episode :104-134, train :137-263).

What changed relative to the reference loop (and nothing else):
  * the `num_parallel_envs` MCTS trees, their envs and the leaf observations live on the GPU
    (BatchedMCTS / BatchedExperienceCollector); the evaluator is called with device tensors;
  * the experience buffer keeps 176-byte env STATES (plus mask / value / visit distribution) and
    the observation batch is produced by the encoder kernel when a batch is sampled: the
    reference keeps a deque of 59 KB NumPy observations and stacks + uploads them per batch
    (:189-197);
  * greedy evaluation runs `--eval_envs` episodes in parallel on the device;
  * wandb is optional and off by default; metrics go to a JSONL file; no interactive visualiser;
  * under torchrun every rank runs its own trees / envs / experience buffer and the gradients are
    averaged with one NCCL all-reduce per batch (the reference is single process).

Run:  python -m pacbot_rl_b200.train_alphazero --num_iters 20 [--device cuda:0]
"""
from __future__ import annotations

import argparse
import json
import shutil
import time
from pathlib import Path
from typing import Optional

import torch
import torch.nn.functional as F

from . import models
from ._lib import NUM_ACTIONS, OBS_SHAPE
from .alphazero import AlphaZeroConfig, BatchedExperienceCollector
from .dist_utils import allreduce_mean_grads, rank_world
from .engine import BatchedPacmanGym

# names and defaults exactly as the reference (train_alphazero.py:22-38)
hyperparam_defaults = {
    "learning_rate": 0.0001,
    "policy_loss_weight": 10.0,
    "num_iters": 10_000,
    "batch_size": 2048,
    "train_instances_per_iter": 2048 * 5,
    "num_parallel_envs": 128,
    "experience_steps": 500,
    "experience_buffer_size": 20_000,
    "tree_size": 100,
    "max_episode_length": 1000,
    "discount_factor": 0.99,
    "evaluate_iters": 1,
    "reward_scale": 1 / 100,
    "grad_clip_norm": 0.1,
    "model": "NetV2",
}


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--eval", metavar="CHECKPOINT", default=None)
    p.add_argument("--no-wandb", action="store_true", default=True)
    p.add_argument("--wandb", dest="no_wandb", action="store_false")
    p.add_argument("--checkpoint-dir", default="checkpoints")
    p.add_argument("--device", default=None)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--eval_envs", type=int, default=1)
    p.add_argument("--metrics-file", default=None)
    p.add_argument("--no-checkpoints", action="store_true")
    p.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", default=True,
                   help="run the collector step eagerly instead of replaying one captured CUDA graph")
    for name, default in hyperparam_defaults.items():
        p.add_argument(f"--{name}", type=type(default), default=default, help="Default: %(default)s")
    return p


class ExperienceBuffer:
    """`deque[ExperienceItem](maxlen=experience_buffer_size)` (train_alphazero.py:148) as a ring
    of env states on the device; observations are encoded per sampled batch."""

    def __init__(self, capacity: int, batch_size: int, device, seed: int = 0) -> None:
        self.capacity = int(capacity)
        self.device = device
        self.states = BatchedPacmanGym(self.capacity, False, False, device=device, seed=seed)
        self.stage = BatchedPacmanGym(int(batch_size), False, False, device=device, seed=seed)
        self.action_mask = torch.zeros((self.capacity, NUM_ACTIONS), dtype=torch.bool, device=device)
        self.value = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.action_distribution = torch.zeros((self.capacity, NUM_ACTIONS), dtype=torch.float32, device=device)
        self._obs = torch.empty((int(batch_size),) + OBS_SHAPE, dtype=torch.float32, device=device)
        self.size = 0
        self.head = 0  # next slot to write

    def __len__(self) -> int:
        return self.size

    def extend(self, batch) -> None:
        k = len(batch)
        src_slots, mask, value, dist = batch.state_slots, batch.action_mask, batch.value, batch.action_distribution
        if k > self.capacity:  # a deque(maxlen) keeps the newest maxlen items
            keep = slice(k - self.capacity, k)
            src_slots, mask, value, dist = src_slots[keep], mask[keep], value[keep], dist[keep]
            k = self.capacity
        slots = (self.head + torch.arange(k, device=self.device)) % self.capacity
        self.states.copy_envs_from(batch.state_engine, src_index=src_slots, dst_index=slots)
        self.action_mask[slots] = mask
        self.value[slots] = value
        self.action_distribution[slots] = dist
        self.head = (self.head + k) % self.capacity
        self.size = min(self.size + k, self.capacity)

    def sample(self, k: int):
        """random.sample(buffer, k) (:187) -> (obs [k,17,28,31], mask, value, distribution)."""
        idx = torch.randperm(self.size, device=self.device)[:k]
        self.stage.copy_envs_from(self.states, src_index=idx, count=k)
        obs = self.stage.obs_range(0, k, self._obs[:k])
        return obs, self.action_mask[idx], self.value[idx], self.action_distribution[idx]


class AlphaZeroTrainer:
    def __init__(self, args) -> None:
        self.args = args
        # under torchrun: one process per GPU, every rank collects with its own trees (global env
        # ids rank * num_parallel_envs ...), gradients are averaged with one all-reduce per batch
        self.rank, local_rank, self.world = rank_world()
        self.device = torch.device(args.device or f"cuda:{local_rank}")
        torch.cuda.set_device(self.device)
        if self.world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=self.device)
        torch.manual_seed(args.seed)  # identical initial weights on every rank
        model_class = getattr(models, args.model)
        self.model = model_class(OBS_SHAPE, NUM_ACTIONS + 1).to(self.device)
        if hasattr(self.model, "use_fused_blocks"):   # NHWC weights, fused block tails / inference head
            self.model.use_fused_blocks()
        torch.manual_seed(args.seed + 7919 * (self.rank + 1))  # per-rank batch sampling / root noise
        self.has_model_updated = False
        self.collector = BatchedExperienceCollector(
            self.model_evaluator,
            AlphaZeroConfig(args.tree_size, args.max_episode_length, args.discount_factor,
                            args.num_parallel_envs),
            device=self.device, seed=args.seed, env_id_base=self.rank * args.num_parallel_envs,
            # NetV2.policy_value takes the padded channels-last batch as the encoder writes it
            leaf_layout="nhwc32" if getattr(self.model, "_fused_blocks", False) else "nchw")
        self.buffer = ExperienceBuffer(args.experience_buffer_size, args.batch_size, self.device, args.seed)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=args.learning_rate)
        self.num_exp_steps_needed = self.buffer.capacity  # start by filling the buffer (:149)
        self.num_train_instances_needed = 0
        self.eval_env: Optional[BatchedPacmanGym] = None

    # train_alphazero.py:81-101
    @torch.no_grad()
    def model_evaluator(self, obs: torch.Tensor, action_mask: torch.Tensor):
        if self.has_model_updated:
            if hasattr(self.model, "policy_value"):   # value scaling + masked softmax inside the head's launch
                return self.model.policy_value(obs, action_mask, 1.0 / self.args.reward_scale)
            predictions = self.model(obs)
            values = predictions[:, -1] / self.args.reward_scale
            policy_logits = predictions[:, :-1].masked_fill(~action_mask, -torch.inf)
            return values, torch.softmax(policy_logits, dim=-1)
        # freshly initialised model: uniform predictions (:97-101)
        value = torch.zeros(obs.shape[0], dtype=torch.float32, device=obs.device)
        m = action_mask.to(torch.float32)
        return value, m / m.sum(dim=-1, keepdim=True)

    # train_alphazero.py:104-123 (greedy branch), `eval_envs` episodes in parallel
    @torch.no_grad()
    def evaluate_episode(self, max_steps: int = 1000):
        self.model.eval()
        k = max(1, int(self.args.eval_envs))
        if self.eval_env is None:
            self.eval_env = BatchedPacmanGym(k, True, False, device=self.device, seed=self.args.seed + 99991)
        env = self.eval_env
        env.reset()
        alive = torch.ones(k, dtype=torch.bool, device=self.device)
        steps = torch.zeros(k, dtype=torch.int32, device=self.device)
        score = torch.zeros(k, dtype=torch.int32, device=self.device)
        for _ in range(max_steps):
            _, policy = self.model_evaluator(env.obs(), env.action_mask())
            actions = policy.argmax(dim=-1).to(torch.uint8)
            _, done = env.step(actions, check_actions=False)
            steps += alive.to(torch.int32)
            score = torch.where(alive, env.score(), score)
            alive &= ~done
            if not bool(alive.any()):
                break
        return float(score.float().mean()), float(steps.float().mean())

    def collect(self) -> int:
        a = self.args
        self.model.eval()
        if a.cuda_graph and self.has_model_updated and self.collector._graph is None:
            # from the first update on the evaluator is the network: capture the collector step
            # (search iteration + env steps + root noise) once; Adam updates the weights in place
            self.collector.enable_cuda_graph()
        self.num_exp_steps_needed = min(self.num_exp_steps_needed + a.experience_steps, self.buffer.capacity)
        collected = 0
        while self.num_exp_steps_needed > 0:
            batch = self.collector.generate_experience(materialize_obs=False)
            self.num_exp_steps_needed -= len(batch)
            self.buffer.extend(batch)
            collected += len(batch)
        return collected

    def train_batches(self) -> dict:
        a = self.args
        stats = {k: [] for k in ("value_loss", "policy_loss", "entropy", "loss", "grad_norm")}
        self.num_train_instances_needed += a.train_instances_per_iter
        predicted_values = value_target = None
        while self.num_train_instances_needed > 0:
            obs, mask, value_target, policy_target = self.buffer.sample(a.batch_size)
            self.num_train_instances_needed -= a.batch_size
            value_target = value_target * a.reward_scale
            self.model.train()
            predictions = self.model(obs)
            predicted_values = predictions[:, -1]
            policy_logits = torch.where(mask, predictions[:, :-1], -torch.inf)
            value_loss = F.mse_loss(predicted_values, value_target)
            log_denominator = torch.logsumexp(policy_logits, dim=-1, keepdim=True)
            unmasked = policy_target * (policy_logits - log_denominator)
            policy_loss = -torch.where(mask, unmasked, 0.0).sum(-1).mean()
            loss = value_loss + a.policy_loss_weight * policy_loss
            self.optimizer.zero_grad()
            loss.backward()
            allreduce_mean_grads(self.model.parameters(), self.world)
            grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=a.grad_clip_norm,
                                                       error_if_nonfinite=True)
            self.optimizer.step()
            with torch.no_grad():
                logp = policy_logits - log_denominator
                entropy = -torch.where(mask, logp.exp() * logp, 0.0).sum(-1).mean()
            for k, v in (("value_loss", value_loss), ("policy_loss", policy_loss), ("entropy", entropy),
                         ("loss", loss), ("grad_norm", grad_norm)):
                stats[k].append(v.detach())
            self.has_model_updated = True
        out = {f"avg_{k}": float(torch.stack(v).mean()) for k, v in stats.items() if k != "grad_norm"}
        out["max_grad_norm"] = float(torch.stack(stats["grad_norm"]).max())
        out["avg_predicted_value"] = float(predicted_values.detach().mean()) / a.reward_scale
        out["avg_target_value"] = float(value_target.mean()) / a.reward_scale
        return out

    def save_checkpoint(self, iter_num: int) -> None:  # train_alphazero.py:252-262
        d = Path(self.args.checkpoint_dir)
        d.mkdir(exist_ok=True)
        torch.save(models.portable_copy(self.model), d / "model-latest.pt")   # row-major weights
        shutil.copyfile(d / "model-latest.pt", d / f"model-iter{iter_num:07}.pt")

    def train(self) -> None:
        a = self.args
        log = open(a.metrics_file, "a") if (a.metrics_file and self.rank == 0) else None
        for iter_num in range(a.num_iters):
            t0 = time.perf_counter()
            with torch.no_grad():
                collected = self.collect()
            t1 = time.perf_counter()
            metrics = self.train_batches()
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            metrics.update(iter=iter_num, collected_steps=collected, collect_s=t1 - t0, train_s=t2 - t1,
                           search_iterations=self.collector.search_iterations, buffer=len(self.buffer))
            if self.rank != 0:
                continue
            metrics["world_size"] = self.world
            if iter_num % a.evaluate_iters == 0:
                score, steps = self.evaluate_episode()
                metrics.update(eval_episode_score=score, eval_episode_steps=steps)
            print(json.dumps(metrics), flush=True)
            if log:
                log.write(json.dumps(metrics) + "\n")
                log.flush()
            if iter_num % 10 == 0 and not a.no_checkpoints:
                self.save_checkpoint(iter_num)


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    if not hasattr(models, args.model):
        raise SystemExit(f"Invalid --model: {args.model!r} (must be a class in models.py)")
    trainer = AlphaZeroTrainer(args)
    print(f"model has {sum(p.numel() for p in trainer.model.parameters())} parameters")
    if args.eval:
        trainer.model = torch.load(args.eval, map_location=trainer.device, weights_only=False)
        trainer.has_model_updated = True
        print(json.dumps(dict(zip(("eval_episode_score", "eval_episode_steps"), trainer.evaluate_episode()))))
        return
    trainer.train()
    if trainer.world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
