"""AlphaZero experience collection on the GPU engine: host-side mirror of
pacbot_rs/src/alphazero.rs (`AlphaZeroConfig` :13-41, `ExperienceItem` :43-49,
`ExperienceCollector` :58-197; stub pacbot_rs/pacbot_rs.pyi:43-70).

`BatchedExperienceCollector` runs `num_parallel_envs` MCTS trees in lock step (BatchedMCTS) and
keeps the unfinished episodes as 176-byte env STATES in a snapshot engine rather than as 59 KB
observations; the observations are produced by the encoder kernel when an episode is handed out
(or, with `materialize_obs=False`, never: the trainer can keep states and encode per batch).
`ExperienceCollector` is the drop-in with the reference's NumPy evaluator and `ExperienceItem`
list."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from ._lib import NUM_ACTIONS, OBS_SHAPE
from .engine import BatchedPacmanGym
from .mcts import BatchedMCTS, numpy_evaluator_adapter


class AlphaZeroConfig:
    """alphazero.rs:13-41 (`#[pyclass(get_all, set_all)]`)."""

    def __init__(self, tree_size: int, max_episode_length: int, discount_factor: float,
                 num_parallel_envs: int) -> None:
        assert tree_size >= 2
        assert max_episode_length >= 1
        assert 0.0 <= discount_factor <= 1.0
        assert num_parallel_envs >= 1
        self.tree_size = int(tree_size)
        self.max_episode_length = int(max_episode_length)
        self.discount_factor = float(discount_factor)
        self.num_parallel_envs = int(num_parallel_envs)

    def __repr__(self) -> str:  # alphazero.rs:38-40 ({self:?})
        return (f"AlphaZeroConfig {{ tree_size: {self.tree_size}, max_episode_length: "
                f"{self.max_episode_length}, discount_factor: {self.discount_factor}, "
                f"num_parallel_envs: {self.num_parallel_envs} }}")


@dataclass
class ExperienceItem:
    """alphazero.rs:43-49."""

    obs: np.ndarray                    # float32 (17, 28, 31)
    action_mask: list                  # 5 bools
    value: float
    action_distribution: list          # 5 floats


@dataclass
class ExperienceBatch:
    """Finished episodes of one generate_experience() call, on the device, episode after episode
    in the order the reference appends them (env order within a step, time order within an
    episode)."""

    obs: Optional[torch.Tensor]            # float32 [K,17,28,31] (None if not materialised)
    action_mask: torch.Tensor              # bool [K,5]
    value: torch.Tensor                    # float32 [K]
    action_distribution: torch.Tensor      # float32 [K,5]
    state_slots: torch.Tensor              # int64 [K]: slots of the snapshot engine holding the states

    def __len__(self) -> int:
        return int(self.value.shape[0])


class BatchedExperienceCollector:
    """ExperienceCollector (alphazero.rs:58-197) with device-resident trees, envs and episodes."""

    def __init__(self, evaluator: Callable, config: AlphaZeroConfig, *, device=None, seed: int = 0,
                 env_id_base: int = 0, noise_fn: Optional[Callable] = None) -> None:
        self.config = config
        P, L = config.num_parallel_envs, config.max_episode_length
        # PacmanGym::new(true, false) + reset (alphazero.rs:72-73)
        self.env = BatchedPacmanGym(P, True, False, device=device, seed=seed, env_id_base=env_id_base)
        self.env.reset()
        self.device = self.env.device
        self.mcts = BatchedMCTS(self.env, evaluator, config.tree_size, noise_fn=noise_fn)
        self.mcts.reset_root()  # MCTSContext::new -> reset_root
        dev = self.device
        self.remaining = torch.full((P,), config.tree_size - 1, dtype=torch.int32, device=dev)
        # unfinished episodes: states in a snapshot engine (slot = env * L + t) + small tensors
        self.snap = BatchedPacmanGym(P * L, False, False, device=dev, seed=seed, env_id_base=env_id_base)
        self.ep_mask = torch.zeros((P, L, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self.ep_dist = torch.zeros((P, L, NUM_ACTIONS), dtype=torch.float32, device=dev)
        self.ep_value = torch.zeros((P, L), dtype=torch.float32, device=dev)
        self.ep_len = np.zeros(P, np.int64)  # host mirror (the step loop reads flags anyway)
        self.search_iterations = 0

    # one parallel search iteration + the env steps that became due (alphazero.rs:97-196)
    def generate_experience_step(self, finished: list) -> None:
        cfg, P, L = self.config, self.config.num_parallel_envs, self.config.max_episode_length
        self.mcts.search_iteration()
        self.search_iterations += 1
        self.remaining -= 1
        act_host = (self.remaining == 0).cpu().numpy()
        if not act_host.any():
            return
        dev = self.device
        idx_host = np.nonzero(act_host)[0]
        idx = torch.from_numpy(idx_host).to(dev)
        act = torch.from_numpy(act_host).to(dev)
        t = torch.from_numpy(self.ep_len[idx_host]).to(dev)
        # experience of the acting envs: state snapshot, mask, visit distribution (:147-161)
        self.snap.copy_envs_from(self.env, src_index=idx, dst_index=idx * L + t)
        self.ep_mask[idx, t] = self.env.action_mask()[idx]
        self.ep_dist[idx, t] = self.mcts.action_distribution()[idx]
        best = self.mcts.best_action()
        reward, done, _ = self.mcts.take_action(best.to(torch.uint8), mask=act)
        self.ep_value[idx, t] = reward[idx].to(torch.float32)
        self.ep_len[idx_host] += 1
        done_host = done.cpu().numpy()
        ended = [int(i) for i in idx_host if done_host[i] or self.ep_len[i] >= L]
        if ended:
            truncated = [i for i in ended if not done_host[i]]
            if truncated:  # bootstrap from the new root's value (:180-184): one mul, one add
                ti = torch.tensor(truncated, device=dev)
                tl = torch.from_numpy(self.ep_len[truncated] - 1).to(dev)
                boot = self.mcts.value()[ti] * cfg.discount_factor
                self.ep_value[ti, tl] = self.ep_value[ti, tl] + boot
            for i in ended:  # end_episode (:163-176), env order
                n = int(self.ep_len[i])
                values = self.ep_value[i, :n].cpu().numpy()
                nxt = np.float32(0.0)
                disc = np.float32(cfg.discount_factor)
                for k in range(n - 1, -1, -1):
                    values[k] = values[k] + disc * nxt
                    nxt = values[k]
                finished.append((i, n, torch.from_numpy(values).to(dev), self.ep_mask[i, :n].clone(),
                                 self.ep_dist[i, :n].clone()))
                self.ep_len[i] = 0
            em = torch.zeros(P, dtype=torch.bool, device=dev)
            ei = torch.tensor(ended, device=dev)
            em[ei] = True
            self.mcts.reset(em, ei)  # mcts_context.reset(): env.reset + fresh root
        counts = self.mcts.node_count()
        self.remaining[idx] = (cfg.tree_size - counts[idx]).clamp(min=0)

    def generate_experience(self, materialize_obs: bool = True) -> ExperienceBatch:
        """generate_experience (alphazero.rs:86-92): at least one finished episode."""
        L = self.config.max_episode_length
        finished: list = []
        pending_obs = []
        while not finished:
            self.generate_experience_step(finished)
            # the snapshot slots of an ended episode are reused by that env's next episode, so
            # encode its observations before searching on
            if materialize_obs:
                for (i, n, *_rest) in finished[len(pending_obs):]:
                    out = torch.empty((n,) + OBS_SHAPE, dtype=torch.float32, device=self.device)
                    pending_obs.append(self.snap.obs_range(i * L, n, out))
        dev = self.device
        slots = torch.cat([torch.arange(i * L, i * L + n, device=dev) for (i, n, *_r) in finished])
        return ExperienceBatch(
            obs=torch.cat(pending_obs) if materialize_obs else None,
            action_mask=torch.cat([f[3] for f in finished]),
            value=torch.cat([f[2] for f in finished]),
            action_distribution=torch.cat([f[4] for f in finished]),
            state_slots=slots,
        )


class ExperienceCollector:
    """pacbot_rs.ExperienceCollector(evaluator, config): NumPy evaluator in, list of
    ExperienceItem out (alphazero.rs:66-92)."""

    def __init__(self, evaluator: Callable, config: AlphaZeroConfig, *, device=None, seed: int = 0) -> None:
        self._evaluator = evaluator
        dev = torch.device("cuda" if device is None else device)
        self._inner = BatchedExperienceCollector(
            lambda obs, mask: numpy_evaluator_adapter(self._evaluator, obs.device)(obs, mask),
            config, device=dev, seed=seed)

    @property
    def config(self) -> AlphaZeroConfig:
        return self._inner.config

    def generate_experience(self) -> list:
        b = self._inner.generate_experience()
        obs = b.obs.cpu().numpy()
        mask = b.action_mask.cpu().numpy()
        value = b.value.cpu().numpy()
        dist = b.action_distribution.cpu().numpy()
        return [ExperienceItem(obs[k].copy(), [bool(x) for x in mask[k]], float(value[k]),
                               [float(x) for x in dist[k]]) for k in range(len(b))]
