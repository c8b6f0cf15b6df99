"""Batched MCTS on the GPU engine: host-side mirror of the reference's `MCTSContext`
(pacbot_rs/src/mcts.rs:113-382; stub pacbot_rs/pacbot_rs.pyi:26-41).

`BatchedMCTS` owns one search tree per env of a `BatchedPacmanGym`; all trees advance in lock
step: select (one kernel) -> leaf observations (k_encode_obs on the simulation slots) ->
evaluator (any torch callable, on the device) -> backprop (one kernel).  Nothing crosses PCIe per
search iteration.  `MCTSContext` is the per-env drop-in with the pyo3 class's surface and a NumPy
evaluator, as `src/algorithms/train_alphazero.py:119-131,269-305` uses it.

Evaluator contracts
  batched : evaluator(obs f32 [n,17,28,31] cuda, mask bool [n,5] cuda) -> (value f32 [n], policy f32 [n,5])
            rows must be independent; rows of finished / inactive trees are ignored
  drop-in : evaluator(obs np.float32 [B,17,28,31], mask np.bool_ [B,5]) -> (value [B], policy [B,5])
            (`EvaluatorFunc`, pacbot_rs.pyi:21-24)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import NUM_ACTIONS, OBS_SHAPE, check, raw_stream
from .engine import BatchedPacmanGym
from .gym import PacmanGym

NOISE_MIX = 0.25         # mcts.rs:262
DIRICHLET_ALPHA = 11.0   # mcts.rs:263

ST_UNTOUCHED, ST_DONE, ST_REROOTED, ST_NEEDS_ROOT, ST_WAS_DONE, ST_BAD_ACTION = range(6)


def dirichlet_root_noise(num_valid: torch.Tensor, alpha: float = DIRICHLET_ALPHA) -> torch.Tensor:
    """Dirichlet(alpha / num_valid) samples of size num_valid per row (mcts.rs:266-271), laid out
    compactly: the k-th valid action takes column k; columns >= num_valid are zero.  A torch
    formulation usable as `noise_fn=lambda nv, mask: dirichlet_root_noise(nv)`; the built-in noise
    (`noise_fn=None`) is drawn by the device kernel `pb_mcts_root_noise_dirichlet` instead."""
    n = num_valid.shape[0]
    cols = torch.arange(NUM_ACTIONS, device=num_valid.device)[None, :] < num_valid[:, None]
    conc = (alpha / num_valid.clamp(min=1).to(torch.float32))[:, None].expand(n, NUM_ACTIONS)
    g = torch._standard_gamma(conc.contiguous()) * cols
    return g / g.sum(dim=1, keepdim=True).clamp(min=torch.finfo(torch.float32).tiny)


class BatchedMCTS:
    """One search tree per env of `env`, advanced in lock step on the device."""

    def __init__(self, env: BatchedPacmanGym, evaluator: Callable, max_nodes: int, *,
                 noise_fn: Optional[Callable] = None, noise_mix: float = NOISE_MIX,
                 dirichlet_alpha: float = DIRICHLET_ALPHA, leaf_layout: str = "nchw") -> None:
        if leaf_layout not in ("nchw", "nhwc32"):
            raise ValueError("leaf_layout: 'nchw' or 'nhwc32'")
        self.env = env
        self.evaluator = evaluator
        # "nchw": the evaluator gets leaf observations as the reference's evaluator does, float32
        # [n, 17, 28, 31].  "nhwc32": channels-last and zero-padded to 32 channels ([n, 32, 28, 31]
        # in channels_last memory format), written that way by the encoder itself -- what
        # NetV2.policy_value / QNetV2 feed their first convolution, without the layout-conversion
        # launch in between.  The observations of fresh roots (reset_root) stay "nchw".
        self.leaf_layout = leaf_layout
        self.max_nodes = int(max_nodes)
        self.noise_mix = float(noise_mix)
        self.dirichlet_alpha = float(dirichlet_alpha)
        # noise_fn(num_valid int64 [n], mask uint8 [n] or None) -> float32 [n,5] (compact layout);
        # tests inject deterministic noise here
        self.noise_fn = noise_fn
        n = env.num_envs
        self.num_envs = n
        self.device = env.device
        self._L = env._L
        # slot i of `sim` receives the clone of env i that a simulation steps (mcts.rs:293)
        self.sim = BatchedPacmanGym(n, False, False, device=env.device, seed=env.seed,
                                    env_id_base=env.env_id_base)
        self._h = C.c_void_p()
        check(self._L.pb_mcts_create(C.byref(self._h), env._h, self.sim._h, self.max_nodes))
        dev = self.device
        self._obs = torch.empty((n,) + OBS_SHAPE, dtype=torch.float32, device=dev)
        if leaf_layout == "nhwc32":
            self._leaf_obs = torch.empty((n, 32) + OBS_SHAPE[1:], dtype=torch.float32, device=dev,
                                         memory_format=torch.channels_last)
        else:
            self._leaf_obs = torch.empty((n,) + OBS_SHAPE, dtype=torch.float32, device=dev)
        self._graph = None
        self._leaf_kind = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._leaf_mask = torch.zeros((n, NUM_ACTIONS), dtype=torch.uint8, device=dev)
        self._value = torch.zeros(n, dtype=torch.float32, device=dev)
        self._policy = torch.zeros((n, NUM_ACTIONS), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.pb_mcts_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> C.c_void_p:
        return raw_stream(self.device)

    def _mask_u8(self, mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if mask is None:
            return None
        m = mask.to(self.device)
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        if m.dtype != torch.uint8 or tuple(m.shape) != (self.num_envs,):
            raise ValueError(f"mask: expected bool/uint8 [{self.num_envs}]")
        return m.contiguous()

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _evaluate(self, obs: torch.Tensor, mask: torch.Tensor):
        value, policy = self.evaluator(obs, mask)
        value = value.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        policy = policy.to(device=self.device, dtype=torch.float32).reshape(-1, NUM_ACTIONS).contiguous()
        return value, policy

    def check(self) -> int:
        """Synchronise and raise if a kernel reported an error; returns the epoch counter."""
        epoch = C.c_uint64(0)
        check(self._L.pb_mcts_check(self._h, C.byref(epoch), self._stream()))
        return int(epoch.value)

    # ------------------------------------------------------------------ roots
    def add_root_noise(self, mask: Optional[torch.Tensor] = None) -> None:
        """add_root_policy_noise (mcts.rs:259-281) for the masked trees."""
        m = self._mask_u8(mask)
        if self.noise_fn is None:   # built-in: drawn on the device, one launch
            check(self._L.pb_mcts_root_noise_dirichlet(self._h, self._p(m), self.dirichlet_alpha,
                                                       self.noise_mix, self._stream()))
            return
        num_valid = self.env.action_mask().sum(dim=1)
        noise = self.noise_fn(num_valid, m)
        noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        check(self._L.pb_mcts_root_noise(self._h, self._p(m), self._p(noise), self.noise_mix,
                                         self._stream()))

    def reset_root(self, mask: Optional[torch.Tensor] = None, index: Optional[torch.Tensor] = None) -> None:
        """reset_root (mcts.rs:252-257): evaluate the real env's state, root := new_visited, noise.

        `index` (int64 device tensor of the env ids selected by `mask`) lets the evaluator run on
        just those rows; without it every row is evaluated and the masked ones are used."""
        m = self._mask_u8(mask)
        obs = self.env.obs(out=self._obs)
        amask = self.env.action_mask()
        if index is not None and index.numel() < self.num_envs:
            v, p = self._evaluate(obs.index_select(0, index), amask.index_select(0, index))
            self._value.index_copy_(0, index, v)
            self._policy.index_copy_(0, index, p)
            value, policy = self._value, self._policy
        else:
            value, policy = self._evaluate(obs, amask)
        check(self._L.pb_mcts_reset_root(self._h, self._p(m), self._p(value), self._p(policy),
                                         self._stream()))
        self.add_root_noise(m)

    def reset(self, mask: Optional[torch.Tensor] = None, index: Optional[torch.Tensor] = None) -> None:
        """MCTSContext::reset (mcts.rs:138-145): reset the env(s) and start fresh trees."""
        self.env.reset(mask)
        self.reset_root(mask, index)

    # ------------------------------------------------------------------ search
    def search_iteration(self, active: Optional[torch.Tensor] = None) -> None:
        """One select -> evaluate -> backprop round for every active tree (the loop body of
        ponder_and_choose, mcts.rs:213-221; one generate_experience_step, alphazero.rs:106-145)."""
        if active is None and self._graph is not None:
            self._graph.replay()
            return
        value, policy = self.select_and_evaluate(active)
        check(self._L.pb_mcts_backprop(self._h, self._p(value), self._p(policy), self._stream()))

    def select_and_evaluate(self, active: Optional[torch.Tensor] = None):
        """The first half of a search iteration: select a trajectory per active tree, encode the
        leaf observations, run the evaluator -> (value [n], policy [n, 5]) for pb_mcts_backprop
        (or the collector's fused tail, pb_mcts_backprop_collect)."""
        a = self._mask_u8(active)
        check(self._L.pb_mcts_select(self._h, self._p(a), self._p(self._leaf_kind),
                                     self._p(self._leaf_mask), self._stream()))
        if self.leaf_layout == "nhwc32":
            leaf_obs = self.sim.obs_nhwc32(out=self._leaf_obs)
        else:
            leaf_obs = self.sim.obs(out=self._leaf_obs)
        return self._evaluate(leaf_obs, self._leaf_mask.view(torch.bool))

    def enable_cuda_graph(self) -> None:
        """Capture select -> encode -> evaluator -> backprop (all trees active) in ONE CUDA graph;
        search_iteration() without a mask then replays it.  The evaluator must be capturable (no
        host sync, static shapes) and must read its weights in place.  The warm-up runs with every
        tree inactive: trees are untouched, only the epoch counter (simulation RNG stream)
        advances."""
        idle = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.search_iteration(idle)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.search_iteration(None)
        self._graph = graph

    def ponder(self, max_tree_size: int) -> torch.Tensor:
        """ponder_and_choose (mcts.rs:208-224) for all trees: grow every tree to `max_tree_size`
        nodes (a tree already that large does nothing), then return best_action() int32 [n]."""
        if max_tree_size > self.max_nodes:
            raise ValueError(f"max_tree_size {max_tree_size} exceeds max_nodes {self.max_nodes}")
        remaining = (max_tree_size - self.node_count()).clamp(min=0)
        for it in range(int(remaining.max().item())):
            self.search_iteration(remaining > it)
        return self.best_action()

    def take_action(self, actions: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """take_action (mcts.rs:153-170) for the masked trees -> (reward int32 [n], done bool [n],
        status uint8 [n]); re-rooted trees get root noise, missing children a fresh root."""
        m = self._mask_u8(mask)
        a = actions.to(device=self.device, dtype=torch.uint8).contiguous()
        n = self.num_envs
        reward = torch.zeros(n, dtype=torch.int32, device=self.device)
        done = torch.zeros(n, dtype=torch.uint8, device=self.device)
        status = torch.zeros(n, dtype=torch.uint8, device=self.device)
        check(self._L.pb_mcts_take_action(self._h, self._p(m), self._p(a), self._p(reward),
                                          self._p(done), self._p(status), self._stream()))
        st = status.cpu()  # tiny D2H: which trees need a fresh root / noise
        if bool((st == ST_WAS_DONE).any()):
            raise RuntimeError("take_action on a finished env (the reference panics: assert!(!is_done))")
        if bool((st == ST_BAD_ACTION).any()):
            raise ValueError("Invalid action")
        needs_root = st == ST_NEEDS_ROOT
        if bool(needs_root.any()):
            idx = needs_root.nonzero().reshape(-1).to(self.device)
            self.reset_root(needs_root.to(self.device), idx)
        rerooted = st == ST_REROOTED
        if bool(rerooted.any()):
            self.add_root_noise(rerooted.to(self.device))
        return reward, done.view(torch.bool), status

    # ------------------------------------------------------------------ queries
    def _query(self, what: int, shape, dtype) -> torch.Tensor:
        out = torch.empty(shape, dtype=dtype, device=self.device)
        check(self._L.pb_mcts_query(self._h, what, self._p(out), self._stream()))
        return out

    def best_action(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_BEST_ACTION, (self.num_envs,), torch.int32)

    def action_distribution(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_ACTION_DISTRIBUTION, (self.num_envs, NUM_ACTIONS), torch.float32)

    def policy_prior(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_POLICY_PRIOR, (self.num_envs, NUM_ACTIONS), torch.float32)

    def action_values(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_ACTION_VALUES, (self.num_envs, NUM_ACTIONS), torch.float32)

    def value(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_VALUE, (self.num_envs,), torch.float32)

    def max_depth(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_MAX_DEPTH, (self.num_envs,), torch.int32)

    def node_count(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_NODE_COUNT, (self.num_envs,), torch.int32)

    def root_visits(self) -> torch.Tensor:
        return self._query(_lib.MCTS_Q_ROOT_VISITS, (self.num_envs,), torch.int32)


def numpy_evaluator_adapter(evaluator: Callable, device) -> Callable:
    """Wrap a reference-style NumPy evaluator (pacbot_rs.pyi:21-24) for BatchedMCTS."""

    def run(obs: torch.Tensor, mask: torch.Tensor):
        value, policy = evaluator(obs.cpu().numpy(), mask.cpu().numpy())
        return (torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)).to(device),
                torch.from_numpy(np.ascontiguousarray(policy, dtype=np.float32)).to(device))

    return run


class MCTSContext:
    """pacbot_rs.MCTSContext(env, evaluator) on the B200 engine (one tree; for many trees use
    BatchedMCTS, which this class wraps with n = 1)."""

    def __init__(self, env: PacmanGym, evaluator: Callable, *, max_nodes: int = 4096) -> None:
        # pyo3 extracts `env: PacmanGym` by value: the context owns a copy (mcts.rs:131-135)
        self._gym = env.clone()
        self.evaluator = evaluator
        self._mcts = BatchedMCTS(self._gym._batch, self._call_evaluator, max_nodes)
        self._mcts.reset_root()

    def _call_evaluator(self, obs: torch.Tensor, mask: torch.Tensor):
        return numpy_evaluator_adapter(self.evaluator, self._mcts.device)(obs, mask)

    @property
    def env(self) -> PacmanGym:  # #[pyo3(get)] on a Clone field returns a copy (mcts.rs:115-116)
        return self._gym.clone()

    def reset(self) -> None:  # mcts.rs:138-145
        self._mcts.reset()
        self._gym.invalidate()   # the tree engine changed the env behind the drop-in's snapshot

    def take_action(self, action: int) -> tuple[int, bool]:  # mcts.rs:153-170
        if self._gym.is_done():
            raise RuntimeError("assertion failed: !self.env.is_done()")
        if not 0 <= int(action) <= 4:
            raise ValueError("Invalid action")
        reward, done, _ = self._mcts.take_action(torch.tensor([int(action)], dtype=torch.uint8))
        self._gym.invalidate()
        return int(reward[0]), bool(done[0])

    def best_action(self) -> int:  # mcts.rs:173-175
        return int(self._mcts.best_action()[0])

    def action_distribution(self) -> list[float]:  # mcts.rs:180-186
        if int(self._mcts.root_visits()[0]) <= 1:
            raise RuntimeError("assertion failed: self.root.visit_count > 1")
        return [float(x) for x in self._mcts.action_distribution()[0].tolist()]

    def policy_prior(self) -> list[float]:  # mcts.rs:189-192
        return [float(x) for x in self._mcts.policy_prior()[0].tolist()]

    def action_values(self) -> list[float]:  # mcts.rs:195-198
        return [float(x) for x in self._mcts.action_values()[0].tolist()]

    def value(self) -> float:  # mcts.rs:201-204
        return float(self._mcts.value()[0])

    def ponder_and_choose(self, max_tree_size: int) -> int:  # mcts.rs:208-224
        if self._gym.is_done():
            raise RuntimeError("assertion failed: !self.env.is_done()")
        action = int(self._mcts.ponder(int(max_tree_size))[0])
        self._mcts.check()
        return action

    def max_depth(self) -> int:  # mcts.rs:227-230
        return int(self._mcts.max_depth()[0])

    def node_count(self) -> int:  # mcts.rs:233-236
        return int(self._mcts.node_count()[0])

    def root_obs_numpy(self) -> np.ndarray:  # mcts.rs:239-242
        return self._gym.obs_numpy()
