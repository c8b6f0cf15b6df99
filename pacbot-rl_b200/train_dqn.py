"""Double-DQN trainer on GPU-resident environments: the device-side mirror of
/root/reference/src/algorithms/train_dqn.py (hyper-parameters and flag names :26-57, loop
:116-265, evaluation :88-113, checkpoint cadence :239-253).

What changed relative to the reference loop (and nothing else):
  * envs, replay ring, epsilon-greedy, sampling and the collate are the repo's CUDA kernels (no
    per-step host trip); the whole iteration is ONE CUDA graph.  The first three iterations run
    eagerly (cuDNN autotune, optimizer state) and are ordinary, counted iterations; the graph is
    captured before the fourth;
  * the evaluation episode runs at the reference's cadence (every `evaluate_steps` = 10
    iterations, train_dqn.py:220-229) but on a SIDE STREAM with a snapshot of the weights, in
    chunks of 25 graph-captured env steps, while training goes on; the host only polls it.  Its
    result is logged with the iteration it was started at;
  * metrics are reduced on the device and read back every `--log_every` iterations instead of a
    `loss.item()` sync per iteration;
  * the phase-1 "old policy" of `reset_env` is optional (its checkpoint is not in the reference
    repo): `--old-policy PATH` enables it;
  * wandb is optional and disabled by default (no network on the GPU box); metrics also go to a
    JSONL file;
  * under torchrun the envs shard per rank and gradients are averaged with ONE NCCL all-reduce
    of the flat 156 934-float gradient per iteration; a `max_seconds` stop is decided by rank 0
    and broadcast, so every rank leaves the loop at the same iteration.

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

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib, models
from ._lib import check
from .engine import BatchedPacmanGym
from .optim import ClipAdam
from .policies import EpsilonGreedy, MaxQPolicy
from .replay_buffer import ReplayBatch, ReplayBuffer, reset_env

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
    p.add_argument("--no-eval-episodes", dest="eval_episodes", action="store_false", default=True,
                   help="skip the evaluation episode every evaluate_steps iterations")
    p.add_argument("--no-checkpoints", action="store_true")
    p.add_argument("--no-channels-last", dest="channels_last", action="store_false", default=True,
                   help="keep the Q-network in NCHW (default: NHWC weights/activations for cuDNN)")
    p.add_argument("--no-fused-blocks", dest="fused_blocks", action="store_false", default=True,
                   help="run bias / max-pool / SiLU of QNetV2's convolution blocks as separate torch "
                        "kernels instead of the fused pb_bias_pool_silu kernels")
    p.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", default=True,
                   help="run the iteration eagerly instead of replaying one captured CUDA graph")
    for name, default in hyperparam_defaults.items():
        p.add_argument(f"--{name}", type=type(default), default=default, help="Default: %(default)s")
    return p


def _fusable(*tensors: torch.Tensor) -> bool:
    """CUDA, contiguous, and float32 where floating: what pb_dqn_target / pb_dqn_loss take."""
    return all(t.is_cuda and t.is_contiguous() and (not t.is_floating_point() or t.dtype == torch.float32)
               for t in tensors)


def dqn_targets(batch: ReplayBatch, q_net, target_q_net, discount_factor: float,
                reward_scale: float) -> torch.Tensor:
    """Double-DQN target (train_dqn.py:167-183): r * reward_scale + gamma * Q_target(s')[argmax_a
    masked Q_online(s')], 0 where the episode ended.  A finished transition has an all-False next
    mask: every Q is -inf there, exactly as in the reference, and the done mask zeroes it.
    On the GPU the arithmetic after the two forwards is ONE launch (pb_dqn_target: the same float32
    operations in the same order, bit-identical to the expressions below)."""
    with torch.no_grad():
        masks = batch.next_action_masks
        q_t, q_o = target_q_net(batch.next_obs), q_net(batch.next_obs)
        if (_fusable(q_t, q_o, masks, batch.rewards, batch.done) and masks.dtype == torch.bool
                and batch.done.dtype == torch.bool and q_t.shape[1] <= 8):
            out = torch.empty(q_t.shape[0], dtype=torch.float32, device=q_t.device)
            p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
            check(_lib.load().pb_dqn_target(
                p(q_t), p(q_o), p(masks), p(batch.rewards), p(batch.done), float(discount_factor),
                float(reward_scale), p(out), q_t.shape[0], q_t.shape[1], q_t.device.index,
                _lib.raw_stream(q_t.device)))
            return out
        next_q = q_t.masked_fill(~masks, -torch.inf)
        online_next_q = q_o.masked_fill(~masks, -torch.inf)
        next_actions = online_next_q.argmax(dim=1)
        returns = next_q.gather(1, next_actions.unsqueeze(1)).squeeze(1)
        discounted = torch.where(batch.done, torch.zeros_like(returns), discount_factor * returns)
        return batch.rewards * reward_scale + discounted


class _QLoss(torch.autograd.Function):
    """mse_loss(Q[range(B), actions], target) of train_dqn.py:193-197 with its gradient and the
    iteration's logged statistics from one launch (pb_dqn_loss)."""

    @staticmethod
    def forward(ctx, all_q, actions, target, stat_scale):
        out = torch.empty(4, dtype=torch.float32, device=all_q.device)
        dq = torch.empty_like(all_q)
        p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
        check(_lib.load().pb_dqn_loss(
            p(all_q), p(actions), p(target), all_q.shape[0], all_q.shape[1], float(stat_scale), p(out),
            p(dq), all_q.device.index, _lib.raw_stream(all_q.device)))
        ctx.save_for_backward(dq)
        loss, stats = out[0], out[1:]
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        (dq,) = ctx.saved_tensors
        return dq * grad_loss, None, None, None


def dqn_loss_stats(batch: ReplayBatch, q_net, target_q: torch.Tensor, stat_scale: float = 1.0):
    """MSE between Q(s)[a] and the target (train_dqn.py:193-197) -> (loss, all Q-values, stats)
    with stats = [stat_scale * mean_b max_a Q, stat_scale * mean(target), isfinite(loss)], the
    quantities train_dqn.py:208-216 logs."""
    all_q = q_net(batch.obs)
    if (_fusable(all_q, batch.actions, target_q) and batch.actions.dtype == torch.int64
            and all_q.shape[1] <= 8):
        loss, stats = _QLoss.apply(all_q, batch.actions, target_q, stat_scale)
        return loss, all_q, stats
    predicted = all_q.gather(1, batch.actions.unsqueeze(1)).squeeze(1)
    loss = F.mse_loss(predicted, target_q)
    with torch.no_grad():
        stats = torch.stack([all_q.detach().amax(dim=1).mean() * stat_scale, target_q.mean() * stat_scale,
                             torch.isfinite(loss.detach()).to(all_q.dtype)])
    return loss, all_q, stats


def dqn_loss(batch: ReplayBatch, q_net, target_q: torch.Tensor):
    """MSE between Q(s)[a] and the target (train_dqn.py:193-197) -> (loss, all Q-values)."""
    loss, all_q, _ = dqn_loss_stats(batch, q_net, target_q)
    return loss, all_q


EVAL_FIELDS = ("score", "steps", "cleared", "pellets_start", "pellets_end", "finished")


class EpisodeEvaluator:
    """`evaluate_episode` (train_dqn.py:88-113) for `num_envs` envs at once, resumable in chunks.

    Per env, exactly the reference's loop: fixed start (random_start=False, random_ticks=False),
    `reset_env`, then up to `max_steps` greedy steps until done; score / lives / pellets are read
    after the step that ended the episode.  Finished envs keep being stepped (harmless, Q12) but
    their statistics are frozen.  A chunk of `chunk` steps is captured in a CUDA graph after its
    first eager run; `alive.any()` and the statistics travel to pinned host memory at the end of
    every chunk, so the host decides between chunks without touching the device."""

    def __init__(self, policy, device, num_envs: int = 1, max_steps: int = 1000, chunk: int = 25,
                 seed: int = 17, env_id_base: int = 10_000_000, phase1_policy=None,
                 use_graph: bool = True) -> None:
        self.policy, self.device, self.n = policy, device, int(num_envs)
        self.max_steps, self.chunk = int(max_steps), max(1, int(chunk))
        self.phase1_policy, self.use_graph = phase1_policy, use_graph
        self.env = BatchedPacmanGym(self.n, False, False, device=device, seed=seed, env_id_base=env_id_base)
        n = self.n
        # a fused QNetV2 takes the observation channels-last and padded to 32 channels: the encoder
        # writes it that way itself (pb_encode_obs_nhwc32), one launch less per evaluation step
        self._nhwc = bool(getattr(getattr(policy, "q_net", None), "_fused_blocks", False))
        if self._nhwc:
            self.obs = torch.empty((n, 32, 28, 31), dtype=torch.float32, device=device,
                                   memory_format=torch.channels_last)
        else:
            self.obs = torch.empty((n, 17, 28, 31), dtype=torch.float32, device=device)
        self.mask = torch.zeros((n, 5), dtype=torch.bool, device=device)
        self._reward = torch.zeros(n, dtype=torch.int32, device=device)
        self._done = torch.zeros(n, dtype=torch.bool, device=device)
        # per-env bookkeeping, updated by ONE kernel per step (pb_eval_track): rows 0 score,
        # 1 steps, 2 lives, 3 pellets_end, 4 finished (done within max_steps), 5 alive, 6 steps left
        self.track = torch.zeros((7, n), dtype=torch.int32, device=device)
        self.pellets_start = torch.zeros(n, dtype=torch.int32, device=device)
        self.stats_dev = torch.zeros((6, n), dtype=torch.int32, device=device)
        self.stats_host = torch.zeros((6, n), dtype=torch.int32).pin_memory()
        self._graph = None
        self._eager_chunks = 0
        self.steps_run = 0

    @torch.no_grad()
    def begin(self) -> None:
        """reset_env(gym) + pellets_start (train_dqn.py:95-97)."""
        reset_env(self.env, self.phase1_policy)
        self.pellets_start.copy_(self.env.remaining_pellets())
        if self._nhwc:
            self.env.obs_nhwc32(out=self.obs)
        else:
            self.env.obs(out=self.obs)
        self.env.action_mask(out=self.mask)
        t = self.track
        t.zero_()
        t[2].fill_(3)
        t[3].copy_(self.pellets_start)
        t[5].fill_(1)
        t[6].fill_(self.max_steps)
        self.steps_run = 0

    @torch.no_grad()
    def _step(self) -> None:
        """One greedy step of every env (train_dqn.py:99-105) + its bookkeeping: forward, step,
        observation encode (one C call) and pb_eval_track."""
        env = self.env
        pol = self.policy
        actions = pol.actions_u8(self.obs, self.mask) if hasattr(pol, "actions_u8") else pol(self.obs, self.mask)
        if self.n <= BatchedPacmanGym.SMALL_BATCH:     # step + encode: one launch, one CTA per env
            done = self._done
            env.step_encode_small(actions, reward_out=self._reward, done_out=done, mask_out=self.mask,
                                  obs_out=None if self._nhwc else self.obs,
                                  nhwc32_out=self.obs if self._nhwc else None)
        elif self._nhwc:
            _, done = env.step(actions, check_actions=False, mask_out=self.mask)
            env.obs_nhwc32(out=self.obs)
        else:
            _, done = env.step_obs(actions, self.obs, mask_out=self.mask)
        check(env._L.pb_eval_track(env._h, C.c_void_p(done.data_ptr()), C.c_void_p(self.track.data_ptr()),
                                   env._stream()))

    @torch.no_grad()
    def _chunk_body(self) -> None:
        for _ in range(self.chunk):
            self._step()
        t = self.track
        finished = t[4] != 0             # `done` of the reference's loop; False if max_steps ran out
        cleared = finished & (t[2] == 3)                                    # train_dqn.py:107
        self.stats_dev.copy_(torch.stack([
            t[0], t[1], cleared.int(), self.pellets_start,
            torch.where(cleared, torch.zeros_like(t[3]), t[3]),             # train_dqn.py:108
            t[4]]))
        self.stats_host.copy_(self.stats_dev, non_blocking=True)

    def run_chunk(self) -> None:
        """Enqueue `chunk` more steps on the current stream (graph replay after the second chunk)."""
        if self._graph is not None:
            self._graph.replay()
        elif self.use_graph and self._eager_chunks >= 2:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream(self.device),
                                  capture_error_mode="thread_local"):
                self._chunk_body()
            self._graph = g
            g.replay()     # the capture pass executed nothing
        else:
            self._chunk_body()
            self._eager_chunks += 1
        self.steps_run += self.chunk

    def all_finished(self) -> bool:
        """Valid once the last enqueued chunk has completed (host reads pinned memory only)."""
        return bool(self.stats_host[5].all()) or self.steps_run >= self.max_steps

    def per_env(self) -> dict:
        return {k: self.stats_host[i].clone() for i, k in enumerate(EVAL_FIELDS)}

    def result(self) -> dict:
        s = self.stats_host.float()
        return {
            "eval_episode_score": s[0].mean().item(),
            "eval_episode_steps": s[1].mean().item(),
            "cleared_board": s[2].mean().item(),
            "eval_pellets_start": s[3].mean().item(),
            "eval_pellets_end": s[4].mean().item(),
        }


def run_eval_episodes(env_policy, device, num_envs: int = 1, max_steps: int = 1000, seed: int = 17,
                      env_id_base: int = 10_000_000, phase1_policy=None, chunk: int = 25) -> dict:
    """Synchronous `evaluate_episode` for `num_envs` envs: per-env int32 tensors (host) for
    score / steps / cleared / pellets_start / pellets_end / finished."""
    ev = EpisodeEvaluator(env_policy, device, num_envs, max_steps, chunk, seed, env_id_base,
                          phase1_policy, use_graph=False)
    ev.begin()
    while True:
        ev.run_chunk()
        torch.cuda.current_stream(device).synchronize()
        if ev.all_finished():
            return ev.per_env()


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
        nhwc = bool(cfg.channels_last and hasattr(self.q_net, "use_channels_last"))
        if nhwc:
            self.q_net.use_channels_last()
            if cfg.fused_blocks and hasattr(self.q_net, "use_fused_blocks"):
                self.q_net.use_fused_blocks()
        torch.manual_seed(cfg.seed * 1000003 + rank)
        self.target_q_net = copy.deepcopy(self.q_net).eval()
        phase1 = None
        if cfg.old_policy:
            import safetensors.torch

            old = model_cls((17, 28, 31), 5).to(device)
            old.load_state_dict(safetensors.torch.load_file(cfg.old_policy))
            phase1 = MaxQPolicy(old.eval())
        # epsilon lives in a device tensor so that changing it needs no re-capture
        self._eps_t = torch.full((1,), cfg.initial_epsilon if cfg.finetune else 1.0, device=device)
        self._eps_host = float(self._eps_t)
        self._graph = None
        self._want_graph = False
        self._eager_iters = 0
        self._side = torch.cuda.Stream(device)
        self.policy = EpsilonGreedy(MaxQPolicy(self.q_net), 5, self._eps_t)
        self.replay = ReplayBuffer(
            maxlen=cfg.replay_buffer_size, policy=self.policy, num_parallel_envs=cfg.num_parallel_envs,
            random_start_proportion=cfg.random_start_proportion, device=device, seed=cfg.seed,
            env_id_base=rank * cfg.num_parallel_envs, phase1_policy=phase1, channels_last=nhwc,
            pad_channels=nhwc and isinstance(self.q_net, models.QNetV2) and phase1 is None)
        # Adam as in the reference (train_dqn.py:133) together with the gradient clipping in front
        # of it (:200-206), as two launches of the repo's kernels (optim.ClipAdam -> pb_clip_adam);
        # torch's pair is ~8 launches even with the fused multi-tensor Adam (0.1 ms of a 2.1 ms
        # iteration), its default foreach Adam a dozen more
        self.params = [p for p in self.q_net.parameters()]
        if len(self.params) <= 32 and all(p.dtype == torch.float32 for p in self.params):
            self.optimizer = ClipAdam(self.params, lr=cfg.learning_rate)
        else:
            self.optimizer = torch.optim.Adam(self.q_net.parameters(), lr=cfg.learning_rate,
                                              capturable=True, fused=True)
        self.iter_num = 0
        # streaming metrics (device side)
        self._acc = torch.zeros(5, device=device)
        self._acc_n = 0
        # evaluation on a side stream (created on first use)
        self._eval: Optional[EpisodeEvaluator] = None
        self._eval_net = None
        self._eval_stream = None
        self._eval_busy = False
        self._eval_iter = -1
        self._eval_event = None
        self.eval_log: list = []

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
        target_q = dqn_targets(batch, self.q_net, self.target_q_net, cfg.discount_factor, cfg.reward_scale)
        # set_to_none: backward then WRITES each gradient instead of adding it to a zeroed one (16
        # accumulation launches and the zero fill per iteration); inside the captured graph the
        # gradients come from the graph's private pool, i.e. the same addresses at every replay
        self.optimizer.zero_grad(set_to_none=True)
        loss, _all_q, stats = dqn_loss_stats(batch, self.q_net, target_q, 1.0 / cfg.reward_scale)
        loss.backward()
        if self.world > 1:
            self._allreduce_grads()
        if isinstance(self.optimizer, ClipAdam):
            grad_norm = self.optimizer.clip_and_step(cfg.grad_clip_norm)
        else:
            grad_norm = torch.nn.utils.clip_grad_norm_(self.params, max_norm=cfg.grad_clip_norm)
            self.optimizer.step()
        with torch.no_grad():   # loss, grad norm, avg predicted value, avg target, finite flag
            self._acc += torch.cat([torch.stack([loss.detach(), grad_norm.detach()]), stats])
        # collect experience (train_dqn.py:263-265); epsilon was set by the host wrapper
        for _ in range(cfg.experience_steps):
            self.replay.generate_experience_step(count_on_host=False)

    def capture_graph(self) -> None:
        """Switch to CUDA-graph mode: the next three iterations still run eagerly on a side stream
        (cuDNN autotune, optimizer state, allocator) -- they are ordinary, counted iterations --
        and `_iteration_body` is captured into ONE graph before the fourth (the reference
        iteration is hundreds of tiny launches + Python; a replay is a single launch)."""
        if self.replay.num_envs * self.replay._ring and len(self.replay) < self.cfg.batch_size:
            raise RuntimeError(f"replay buffer holds {len(self.replay)} recorded transitions, fewer than "
                               f"one batch ({self.cfg.batch_size}): call replay.fill(batch_size) first")
        self._want_graph = True

    def _run_body(self) -> None:
        if self._graph is not None:
            self._graph.replay()
            return
        if not self._want_graph:
            self._iteration_body()
            return
        cur = torch.cuda.current_stream(self.device)
        if self._eager_iters >= 3:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._iteration_body()
            self._graph.replay()    # the capture pass executed nothing
            return
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            self._iteration_body()
        cur.wait_stream(self._side)
        self._eager_iters += 1

    def train_iteration(self) -> None:
        cfg = self.cfg
        if self.iter_num % cfg.target_network_update_steps == 0:
            self.update_target()
        # anneal epsilon (train_dqn.py:256-260).  The reference assigns it after the update and
        # before the experience steps of the same iteration; the update does not read epsilon, so
        # assigning it before the (captured) body is equivalent.
        self.epsilon = lerp(cfg.initial_epsilon, cfg.final_epsilon,
                            self.iter_num / max(1, cfg.num_iters - 1))
        self._run_body()
        self.replay._t += cfg.experience_steps
        self._acc_n += 1
        self.iter_num += 1

    def _allreduce_grads(self) -> None:
        from .dist_utils import allreduce_mean_grads

        allreduce_mean_grads(self.params, self.world, assign_views=True)

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
    def _make_evaluator(self, num_envs: int, use_graph: bool) -> EpisodeEvaluator:
        self._eval_net = copy.deepcopy(self.q_net).eval()
        return EpisodeEvaluator(MaxQPolicy(self._eval_net), self.device, num_envs, 1000, 25,
                                seed=self.cfg.seed + 17, env_id_base=10_000_000,
                                phase1_policy=self.replay.phase1_policy, use_graph=use_graph)

    @torch.no_grad()
    def evaluate_episode(self, max_steps: int = 1000, num_envs: int = 1) -> dict:
        """Synchronous greedy episode(s) from the fixed start, after the old policy's phase when
        there is one (`reset_env(gym)`, train_dqn.py:95-96), with the CURRENT weights."""
        net = copy.deepcopy(self.q_net).eval()
        per = run_eval_episodes(MaxQPolicy(net), self.device, num_envs, max_steps,
                                seed=self.cfg.seed + 17, env_id_base=10_000_000 + self.iter_num * num_envs,
                                phase1_policy=self.replay.phase1_policy)
        f = {k: v.float().mean().item() for k, v in per.items()}
        return {"eval_episode_score": f["score"], "eval_episode_steps": f["steps"],
                "cleared_board": f["cleared"], "eval_pellets_start": f["pellets_start"],
                "eval_pellets_end": f["pellets_end"]}

    def maybe_evaluate(self, force: bool = False) -> None:
        """Call after `train_iteration()`.  Starts an evaluation episode when the iteration just
        finished is a multiple of `evaluate_steps` (train_dqn.py:220), on the evaluation stream
        with a snapshot of the weights, and otherwise just advances a running one by a chunk when
        its last chunk has completed (`event.query()`, no blocking)."""
        if not self.cfg.eval_episodes:
            return
        due = force or (self.iter_num - 1) % self.cfg.evaluate_steps == 0
        if self._eval_busy:
            self._poll_evaluation(block=due)   # one evaluation at a time, like the reference
        if due and not self._eval_busy:
            self._start_evaluation()

    def _start_evaluation(self) -> None:
        if self._eval is None:
            self._eval_stream = torch.cuda.Stream(self.device)
            self._eval = self._make_evaluator(self.cfg.eval_envs, self.cfg.cuda_graph)
            self._eval_event = torch.cuda.Event()
        cur = torch.cuda.current_stream(self.device)
        es = self._eval_stream
        es.wait_stream(cur)                       # weights of the iteration just enqueued
        with torch.cuda.stream(es), torch.no_grad():
            torch._foreach_copy_(list(self._eval_net.parameters()), list(self.q_net.parameters()))
            snap = torch.cuda.Event()
            snap.record(es)
            self._eval.begin()
            self._eval.run_chunk()
            self._eval_event.record(es)
        cur.wait_event(snap)                      # the next optimizer step waits for the snapshot only
        self._eval_busy, self._eval_iter = True, self.iter_num - 1

    def _poll_evaluation(self, block: bool = False) -> None:
        while self._eval_busy:
            if block:
                self._eval_event.synchronize()
            elif not self._eval_event.query():
                return
            if self._eval.all_finished():
                res = self._eval.result()
                res["eval_iter"] = self._eval_iter
                self.eval_log.append(res)
                self._eval_busy = False
                return
            with torch.cuda.stream(self._eval_stream):
                self._eval.run_chunk()
                self._eval_event.record(self._eval_stream)
            if not block:
                return

    def finish_evaluation(self) -> None:
        self._poll_evaluation(block=True)

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
    dist = None
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
    trainer.replay.fill(cfg.batch_size)
    trainer.epsilon = cfg.initial_epsilon
    if cfg.cuda_graph:
        trainer.capture_graph()
    metrics_f = open(cfg.metrics_file, "a") if (cfg.metrics_file and rank == 0) else None
    t0 = time.time()
    last_t, last_it, evals_logged = t0, 0, 0
    for it in range(cfg.num_iters):
        trainer.train_iteration()
        if rank == 0:
            trainer.maybe_evaluate()    # the evaluation episode is rank 0's, as the log is
        last = it == cfg.num_iters - 1
        if (it + 1) % cfg.log_every == 0 or last:
            if last and rank == 0:
                trainer.finish_evaluation()
            m = trainer.pop_metrics()
            if len(trainer.eval_log) > evals_logged:   # the most recent finished evaluation
                m.update(trainer.eval_log[-1])
                m["evals_since_last_log"] = len(trainer.eval_log) - evals_logged
                evals_logged = len(trainer.eval_log)
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
            if max_seconds is not None:
                # rank 0 decides, every rank obeys: leaving the loop at different iterations
                # would hang the others in the in-graph all-reduce
                stop = torch.tensor([1 if time.time() - t0 > max_seconds else 0], device=device)
                if dist is not None:
                    dist.broadcast(stop, src=0)
                if bool(stop.item()):
                    break
        if not cfg.no_checkpoints and rank == 0 and it % 500 == 0:
            trainer.save_checkpoint(cfg.checkpoint_dir, it)
    if rank == 0:
        trainer.finish_evaluation()
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
