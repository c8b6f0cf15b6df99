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

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from ._lib import NUM_ACTIONS, OBS_SHAPE, check
from .engine import BatchedPacmanGym
from .mcts import ST_DONE, ST_NEEDS_ROOT, ST_REROOTED, ST_WAS_DONE, BatchedMCTS, numpy_evaluator_adapter


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
    state_engine: BatchedPacmanGym         # engine holding the K env states ...
    state_slots: torch.Tensor              # ... at these slots (int64 [K]); valid until the next call

    def __len__(self) -> int:
        return int(self.value.shape[0])


class BatchedExperienceCollector:
    """ExperienceCollector (alphazero.rs:58-197) with device-resident trees, envs and episodes.

    One generate_experience_step = one search iteration of every tree (BatchedMCTS) + ONE kernel
    for the env steps that became due (pb_mcts_collect_act: experience record, take_action, next
    search budget) + root noise for the re-rooted trees; the host reads one flag word per step
    and only gets involved when an episode ended."""

    def __init__(self, evaluator: Callable, config: AlphaZeroConfig, *, device=None, seed: int = 0,
                 env_id_base: int = 0, noise_fn: Optional[Callable] = None,
                 leaf_layout: str = "nchw") -> None:
        self.config = config
        P, L = config.num_parallel_envs, config.max_episode_length
        # PacmanGym::new(true, false) + reset (alphazero.rs:72-73)
        self.env = BatchedPacmanGym(P, True, False, device=device, seed=seed, env_id_base=env_id_base)
        self.env.reset()
        self.device = self.env.device
        self.mcts = BatchedMCTS(self.env, evaluator, config.tree_size, noise_fn=noise_fn,
                                leaf_layout=leaf_layout)
        self.mcts.reset_root()  # MCTSContext::new -> reset_root
        dev = self.device
        self.remaining = torch.full((P,), config.tree_size - 1, dtype=torch.int32, device=dev)
        # unfinished episodes: states in a snapshot engine (slot = env * L + t) + small tensors;
        # finished ones move to the outbox engine until they are handed out
        self.snap = BatchedPacmanGym(P * L, False, False, device=dev, seed=seed, env_id_base=env_id_base)
        self.outbox = BatchedPacmanGym(P * L, False, False, device=dev, seed=seed, env_id_base=env_id_base)
        self.ep_mask = torch.zeros((P, L, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self.ep_dist = torch.zeros((P, L, NUM_ACTIONS), dtype=torch.float32, device=dev)
        self.ep_value = torch.zeros((P, L), dtype=torch.float32, device=dev)
        self.ep_len = torch.zeros(P, dtype=torch.int32, device=dev)
        self._status = torch.zeros(P, dtype=torch.uint8, device=dev)
        self._flags = torch.zeros(1, dtype=torch.int32, device=dev)
        self._graph = None
        # backprop + env steps + built-in root noise as one launch (pb_mcts_backprop_collect);
        # False: the three separate launches (what runs with an injected noise_fn anyway)
        self.fused_tail = True
        # the fused tail reports to the host through a pinned word its last CTA publishes
        # ((launch count << 32) | flags): the per-step flag read is a poll of host memory, not a
        # copy + stream synchronisation.  None: the flags are read from the device (`.item()`).
        self._word = None
        self._word_expected = 0
        step_word = getattr(self.mcts._L, "pb_mcts_step_word", None)
        if step_word is not None and noise_fn is None:
            addr = C.c_void_p()
            check(step_word(self.mcts._h, C.byref(addr)))
            self._word = C.c_uint64.from_address(addr.value)
        self._pending: list = []   # finished episodes not handed out yet
        self._out_count = 0
        self.search_iterations = 0

    # ------------------------------------------------------------------ device part of a step
    def _device_step(self) -> None:
        cfg = self.config
        p = BatchedMCTS._p
        m = self.mcts
        if self.fused_tail and m.noise_fn is None and m._graph is None:
            # built-in root noise: everything after the evaluator -- backprop, the env steps that
            # became due, the noise on the re-rooted trees -- is ONE launch
            value, policy = m.select_and_evaluate()
            check(m._L.pb_mcts_backprop_collect(
                m._h, p(value), p(policy), self.snap._h, cfg.max_episode_length, cfg.tree_size,
                p(self.remaining), p(self.ep_len), p(self.ep_mask), p(self.ep_dist), p(self.ep_value),
                p(self._status), p(self._flags), m.dirichlet_alpha, m.noise_mix,
                1 if self._word is not None else 0, m._stream()))
            return
        self.mcts.search_iteration()
        check(self.mcts._L.pb_mcts_collect_act(
            self.mcts._h, self.snap._h, cfg.max_episode_length, cfg.tree_size, p(self.remaining),
            p(self.ep_len), p(self.ep_mask), p(self.ep_dist), p(self.ep_value), p(self._status),
            p(self._flags), self.mcts._stream()))
        # take_action re-rooted these trees: root noise (mcts.rs:160-163)
        self.mcts.add_root_noise((self._status & 7) == ST_REROOTED)

    def enable_cuda_graph(self) -> None:
        """Capture the device part of a step (search iteration, env steps, root noise) in ONE CUDA
        graph.  Needs a capturable evaluator and the built-in Dirichlet noise.  The warm-up
        consists of three ordinary steps; episodes they finish are handed out by the next
        generate_experience()."""
        if self.mcts.noise_fn is not None:
            raise ValueError("a host noise_fn cannot be captured in a CUDA graph")
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.generate_experience_step(self._pending)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._device_step()
        self._graph = graph

    def _ensure_outbox(self, need: int) -> None:
        """The outbox holds the states of finished episodes until `generate_experience` hands them
        out.  P * L slots are enough for one hand-out of ordinary episodes; callers that keep
        stepping without handing out (the graph warm-up with very short episodes, a benchmark loop
        over `generate_experience_step`) make it grow instead of losing states."""
        have = self.outbox.num_envs
        if need <= have:
            return
        bigger = BatchedPacmanGym(max(need, 2 * have), False, False, device=self.device,
                                  seed=self.outbox.seed, env_id_base=self.outbox.env_id_base)
        if self._out_count:
            bigger.copy_envs_from(self.outbox, count=self._out_count)
        self.outbox = bigger

    # ------------------------------------------------------------------ host part of a step
    def _finish(self, finished: list) -> None:
        """Episodes that ended in this step (alphazero.rs:163-184) and trees that need a fresh
        root (mcts.rs:164-166)."""
        cfg, P, L = self.config, self.config.num_parallel_envs, self.config.max_episode_length
        dev = self.device
        status = self._status.cpu().numpy()
        st = status & 7
        if (st == ST_WAS_DONE).any():
            raise RuntimeError("collector stepped a finished env (assert!(!is_done) in the reference)")
        needs_root = st == ST_NEEDS_ROOT
        if needs_root.any():
            idx = torch.from_numpy(np.nonzero(needs_root)[0]).to(dev)
            self.mcts.reset_root(torch.from_numpy(needs_root).to(dev), idx)
            self.remaining[idx] = cfg.tree_size - 1
        ended = np.nonzero((st == ST_DONE) | ((status & 8) != 0))[0]
        if ended.size == 0:
            return
        ep_len = self.ep_len.cpu().numpy()
        truncated = [int(i) for i in ended if status[i] & 8]
        if truncated:  # bootstrap from the new root's value (:180-184): one mul, one add
            ti = torch.tensor(truncated, device=dev)
            tl = torch.from_numpy(ep_len[truncated].astype(np.int64) - 1).to(dev)
            boot = self.mcts.value()[ti] * cfg.discount_factor
            self.ep_value[ti, tl] = self.ep_value[ti, tl] + boot
        for i in (int(i) for i in ended):  # end_episode (:163-176), env order
            n = int(ep_len[i])
            values = self.ep_value[i, :n].cpu().numpy()
            nxt = np.float32(0.0)
            disc = np.float32(cfg.discount_factor)
            for k in range(n - 1, -1, -1):
                values[k] = values[k] + disc * nxt
                nxt = values[k]
            # the snapshot slots are reused by this env's next episode: move the states out
            self._ensure_outbox(self._out_count + n)
            src = torch.arange(i * L, i * L + n, device=dev)
            dst = torch.arange(self._out_count, self._out_count + n, device=dev)
            self.outbox.copy_envs_from(self.snap, src_index=src, dst_index=dst)
            finished.append((dst, torch.from_numpy(values).to(dev), self.ep_mask[i, :n].clone(),
                             self.ep_dist[i, :n].clone()))
            self._out_count += n
        ei = torch.from_numpy(ended).to(dev)
        em = torch.zeros(P, dtype=torch.bool, device=dev)
        em[ei] = True
        self.ep_len[ei] = 0
        self.mcts.reset(em, ei)  # mcts_context.reset(): env.reset + fresh root
        self.remaining[ei] = cfg.tree_size - 1   # tree_size - node_count() of a fresh tree (:192-193)

    # one parallel search iteration + the env steps that became due (alphazero.rs:97-196)
    def generate_experience_step(self, finished: list) -> None:
        if self._graph is not None:
            self._graph.replay()
        else:
            self._device_step()
        self.search_iterations += 1
        if self._publishing():
            flags = self._wait_word()             # the one host read per step: a pinned word
        else:
            flags = int(self._flags.item())       # ... or a device read
            if flags:
                self._flags.zero_()
        if flags:
            self._finish(finished)

    def _publishing(self) -> bool:
        m = self.mcts
        return self._word is not None and self.fused_tail and m.noise_fn is None and m._graph is None

    def _wait_word(self) -> int:
        """Wait for the step just enqueued to publish its word -> the step's flags."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("generate_experience_step waits for the device: it cannot run inside a CUDA "
                               "graph capture (use enable_cuda_graph(), which captures the device part only)")
        self._word_expected = (self._word_expected + 1) & 0xFFFFFFFF    # the word carries 32 bits of it
        want, word = self._word_expected, self._word
        spins = 0
        while (word.value >> 32) != want:
            spins += 1
            if spins % 200_000 == 0:
                # nothing for a long time: drain the device (surfaces a kernel error) and trust the
                # count the device reports from here on
                torch.cuda.synchronize(self.device)
                got = word.value >> 32
                if got != want:
                    self._word_expected = got
                    break
        return int(word.value & 0xFFFFFFFF)

    def generate_experience(self, materialize_obs: bool = True) -> ExperienceBatch:
        """generate_experience (alphazero.rs:86-92): at least one finished episode."""
        finished, self._pending = self._pending, []
        if not finished:
            self._out_count = 0
        while not finished:
            self.generate_experience_step(finished)
        self._pending = []
        slots = torch.cat([f[0] for f in finished])
        obs = None
        if materialize_obs:  # the outbox is filled front to back: one encode launch
            k = self._out_count
            obs = torch.empty((k,) + OBS_SHAPE, dtype=torch.float32, device=self.device)
            self.outbox.obs_range(0, k, obs)
        batch = ExperienceBatch(
            obs=obs,
            action_mask=torch.cat([f[2] for f in finished]),
            value=torch.cat([f[1] for f in finished]),
            action_distribution=torch.cat([f[3] for f in finished]),
            state_engine=self.outbox,
            state_slots=slots,
        )
        self._out_count = 0  # the next call overwrites the outbox
        return batch


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
