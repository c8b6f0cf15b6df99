"""GPU-resident replay buffer + experience generation: the batched mirror of the reference's
`ReplayBuffer` (/root/reference/src/replay_buffer.py:57-143) and `reset_env` (:44-54).

Layout: a time ring of R = ceil(maxlen / N) + 1 slots; slot t % R holds, for all N envs, the
observation BEFORE action t plus (action, reward, done, next_action_mask, valid) of transition t.
The observation encoder writes straight into the ring (`pb_encode_obs_ring`), so `next_obs` of
transition t is simply slot (t+1) % R and is never copied; a finished episode's `next_obs` is
"None" in the reference (zeros at collate time, train_dqn.py:154-159) and the slot then already
holds the first observation of the next episode (the env was reset in the same kernel launch).
One slot is always "current observation only", which is why R has the +1.

Every per-step and per-batch operation is one of the repo's own kernels behind the C ABI
(include/pacbot_b200.h, csrc/pb_replay_abi.inl):

    generate_experience_step   pb_replay_slot_obs (last_obs, NCHW or channels-last)
                               -> Q-network forward (PyTorch/cuDNN)
                               -> pb_select_actions_dev (epsilon-greedy over the Q-values)
                               -> pb_step (+ auto reset) -> pb_encode_obs_ring -> pb_replay_record
    sample_batch               pb_replay_sample (uniform, without replacement: random.sample,
                               replay_buffer.py:141-143) -> pb_replay_collate (obs / next_obs /
                               action / reward / done / next mask of the batch in one launch)

Everything that depends on the ring position reads a DEVICE-side step counter (`_t_dev`), never a
Python int, so `generate_experience_step` + `sample_batch` can be captured in a CUDA graph and
replayed while the ring advances."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from ._lib import NUM_ACTIONS, OBS_SHAPE, check, raw_stream
from .engine import BatchedPacmanGym
from .policies import EpsilonGreedy, MaxQPolicy, Policy


class ReplayBatch(NamedTuple):
    """Collated transitions (what train_dqn.py:153-165 builds from a list of ReplayItems)."""

    obs: torch.Tensor                # float32 [B, 17, 28, 31] (channels-last strides, and 32
                                     # channels of which 17..31 are zero, if asked)
    actions: torch.Tensor            # int64   [B]
    rewards: torch.Tensor            # float32 [B]   (raw env reward, unscaled)
    next_obs: torch.Tensor           # float32 [B, 17, 28, 31], zeros where done
    done: torch.Tensor               # bool    [B]
    next_action_masks: torch.Tensor  # bool    [B, 5], all False where done


@torch.no_grad()
def reset_env(envs: BatchedPacmanGym, policy_old: Optional[Policy], max_steps: int = 100_000) -> None:
    """`reset_env` (replay_buffer.py:44-54) for every env of `envs`, synchronous form: reset, then
    `policy_old` plays until `first_ai_done()`, an env that dies on the way being reset again.
    Envs that are through wait -- parked in a second engine and restored after every step -- until
    the last one is.  `ReplayBuffer` does the same thing interleaved with recording (per-env phase
    flag); this form is for callers that need every env through phase 1 before they go on
    (`evaluate_episode`, train_dqn.py:95-96).  `policy_old=None` is a plain `reset()` (the
    reference cannot start without `q_net-old.safetensors`).  The reference would loop
    forever on an old policy that never finishes; here the loop stops after `max_steps`."""
    envs.reset()
    if policy_old is None:
        return
    n, dev = envs.num_envs, envs.device
    waiting = ~envs.first_ai_done()
    if not bool(waiting.any()):
        return
    parked = BatchedPacmanGym(n, False, False, device=dev)
    ids = torch.arange(n, dtype=torch.int64, device=dev)
    obs = envs.obs()
    auto = envs.auto_reset
    envs.auto_reset = True                      # "the first ai died :( try again" (:52-54)
    try:
        for _ in range(max_steps):
            through = torch.where(waiting, -1, ids).contiguous()   # negative index = skip
            parked.copy_envs_from(envs, through, through)
            envs.step_obs(policy_old(obs, envs.action_mask()), obs)
            envs.copy_envs_from(parked, through, through)
            waiting = waiting & ~envs.first_ai_done()
            if not bool(waiting.any()):
                break
    finally:
        envs.auto_reset = auto


def _channels_last_empty(shape, device) -> torch.Tensor:
    """float32 [B, C, H, W] tensor whose memory order is [B, H, W, C] (torch.channels_last)."""
    b, c, h, w = shape
    return torch.empty((b, h, w, c), dtype=torch.float32, device=device).permute(0, 3, 1, 2)


# channels of a zero-padded channels-last observation batch: cuDNN's tensor-core kernels for the
# first convolution want input channels that are a multiple of 8 (tools/experiments/
# pad_channels_probe.py: 17 -> 32 channels, 230 -> 140 us forward at batch 512)
PADDED_CHANNELS = 32


class ReplayBuffer:
    def __init__(
        self,
        maxlen: int,
        policy: Policy,
        num_parallel_envs: int,
        random_start_proportion: float,
        device: torch.device = torch.device("cuda"),
        *,
        seed: int = 0,
        env_id_base: int = 0,
        phase1_policy: Optional[Policy] = None,
        channels_last: bool = False,
        pad_channels: bool = False,
    ) -> None:
        """`channels_last`: observation batches (`sample_batch`, the policy's `last_obs`) come out
        in channels-last memory order; `pad_channels` (with channels_last): with `PADDED_CHANNELS`
        channels, the ones past the 17th all zero -- for networks that accept that
        (`QNetV2.forward` pads its first weight to match)."""
        self.policy = policy
        self.phase1_policy = phase1_policy
        self.channels_last = bool(channels_last)
        self.obs_channels = PADDED_CHANNELS if (channels_last and pad_channels) else OBS_SHAPE[0]
        # `channels_last` argument of pb_replay_collate / pb_replay_slot_obs
        self._cl_arg = self.obs_channels if self.channels_last else 0
        self._batch_shape = (self.obs_channels,) + tuple(OBS_SHAPE[1:])
        n = int(num_parallel_envs)
        self.num_envs = n
        self.maxlen = int(maxlen)
        self.seed = int(seed)
        self._slots = -(-self.maxlen // n)          # transitions kept: slots * n >= maxlen
        self._ring = self._slots + 1
        # replay_buffer.py:75-80: random_start for the first N * proportion envs, random_ticks all
        self.envs = BatchedPacmanGym(
            n, np.arange(n) < n * random_start_proportion, True, device=torch.device(device),
            seed=seed, env_id_base=env_id_base, auto_reset=True)
        dev = self.envs.device
        self.device = dev
        self._L = _lib.load()
        r = self._ring
        self._obs = torch.zeros((r, n) + OBS_SHAPE, dtype=torch.float32, device=dev)
        self._actions = torch.zeros((r, n), dtype=torch.uint8, device=dev)
        self._rewards = torch.zeros((r, n), dtype=torch.int32, device=dev)
        self._done = torch.zeros((r, n), dtype=torch.bool, device=dev)
        self._next_mask = torch.zeros((r, n, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self._valid = torch.zeros((r, n), dtype=torch.bool, device=dev)   # True = sampleable
        self._mask = torch.zeros((n, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self._in_phase1 = torch.zeros(n, dtype=torch.bool, device=dev)
        # per-step scratch (fixed addresses: CUDA-graph capture)
        self._last_obs = (_channels_last_empty((n,) + self._batch_shape, dev) if self.channels_last
                          else torch.empty((n,) + OBS_SHAPE, dtype=torch.float32, device=dev))
        # True once `_last_obs` holds the observation of ring slot t (written by the encoder of the
        # step before); anything that rewrites the current slot behind the buffer's back has to
        # call invalidate_last_obs()
        self._last_obs_fresh = False
        self._act = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._reward = torch.zeros(n, dtype=torch.int32, device=dev)
        self._done_step = torch.zeros(n, dtype=torch.bool, device=dev)
        self._eps_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._t = 0                                                  # host mirror of the step counter
        self._t_dev = torch.zeros(1, dtype=torch.int64, device=dev)  # the counter the kernels read
        self._nxt_dev = torch.ones(1, dtype=torch.int64, device=dev)  # (t + 1) % R
        self._sample_calls = torch.zeros(1, dtype=torch.int64, device=dev)
        self._num_valid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._batch_bufs: dict = {}
        self.envs.reset()
        self.envs.obs(out=self._obs[0])
        self.envs.action_mask(out=self._mask)
        if phase1_policy is not None:
            # reset_env (replay_buffer.py:44-54): a freshly reset env is first played by the old
            # policy until first_ai_done(); those steps are not recorded
            self._in_phase1 = ~self.envs.first_ai_done()

    # ------------------------------------------------------------------ reference surface
    @property
    def obs_shape(self) -> torch.Size:
        return torch.Size(OBS_SHAPE)

    @property
    def num_actions(self) -> int:
        return NUM_ACTIONS

    def __len__(self) -> int:
        """Number of RECORDED transitions that can be sampled (len(deque) in the reference).
        Reads the device (one synchronisation)."""
        return int(self._valid.sum().item())

    @property
    def capacity(self) -> int:
        return self._slots * self.num_envs

    def invalidate_last_obs(self) -> None:
        """The current ring slot was rewritten from outside (tests that plant env states): the
        next step re-derives the policy's input from the slot."""
        self._last_obs_fresh = False

    def fill(self, min_valid: int = 1) -> None:
        """Generate experience until the ring has wrapped once (the reference's `fill()` loops until
        the deque holds `maxlen` items, replay_buffer.py:97-100) AND at least `min_valid` recorded
        transitions can be sampled.  With the old policy of the two-stage reset, phase-1 steps
        occupy ring slots without being recorded, so a freshly wrapped ring can hold fewer than
        `capacity` transitions -- possibly fewer than a batch: callers pass their batch size."""
        limit = 16 * self._ring + 64
        while self._t < self._slots:
            self.generate_experience_step()
        while len(self) < max(1, min_valid):
            if self._t >= limit:
                raise RuntimeError(
                    f"replay buffer: only {len(self)} recorded transitions after {self._t} steps of "
                    f"{self.num_envs} envs (need {min_valid}); the old policy of the two-stage reset "
                    "does not get through phase 1")
            for _ in range(8):
                self.generate_experience_step()

    def _slot_ptrs(self):
        return (C.c_void_p(self._actions.data_ptr()), C.c_void_p(self._rewards.data_ptr()),
                C.c_void_p(self._done.data_ptr()), C.c_void_p(self._next_mask.data_ptr()))

    def _stream(self) -> C.c_void_p:
        return raw_stream(self.device)

    @torch.no_grad()
    def _select(self, last_obs: torch.Tensor) -> torch.Tensor:
        """actions uint8 [N] of `self.policy` for the current observations."""
        pol = self.policy
        if isinstance(pol, EpsilonGreedy) and isinstance(pol.original_policy, MaxQPolicy):
            # EpsilonGreedy(MaxQPolicy(q_net)) (policies.py:16-59) as ONE launch after the forward:
            # greedy iff rand > epsilon, else a uniformly random valid action; epsilon and the
            # call index are read from device memory at kernel time
            q = pol.original_policy.q_values(last_obs)
            eps = pol.epsilon
            if isinstance(eps, torch.Tensor):
                eps_dev = eps.reshape(1) if eps.dtype == torch.float32 and eps.device == self.device \
                    else self._eps_dev.copy_(eps.reshape(1))
            else:
                eps_dev = self._eps_dev.fill_(float(eps))
            return self.envs.select_actions_dev(q.contiguous(), eps_dev, self._t_dev, mask=self._mask,
                                                out=self._act)
        return self._act.copy_(pol(last_obs, self._mask))

    @torch.no_grad()
    def generate_experience_step(self, count_on_host: bool = True) -> None:
        """One env-step for each parallel env, appended to the ring (replay_buffer.py:102-139).
        `count_on_host=False` is for CUDA-graph replay: the caller advances `_t` itself."""
        r, n = self._ring, self.num_envs
        L, stream, dev = self._L, self._stream(), self.device.index
        # the policy's input: ring slot t in the network's layout.  With the padded channels-last
        # layout the encoder of the PREVIOUS step has already written it next to the ring slot
        # (one launch, pb_encode_obs_ring_nhwc32); otherwise, and for the first step, it is a
        # transposing copy of the slot.
        dual = self._cl_arg == PADDED_CHANNELS
        if not (dual and self._last_obs_fresh):
            check(L.pb_replay_slot_obs(C.c_void_p(self._obs.data_ptr()), n, r, C.c_void_p(self._t_dev.data_ptr()),
                                       C.c_void_p(self._last_obs.data_ptr()), self._cl_arg, dev, stream))
        actions = self._select(self._last_obs)
        if self.phase1_policy is not None:
            old = self.phase1_policy(self._last_obs, self._mask).to(torch.uint8)
            actions = self._act.copy_(torch.where(self._in_phase1, old, actions))
        # step + auto reset, then the observation of the resulting state straight into slot t+1
        if n <= BatchedPacmanGym.SMALL_BATCH:
            # ... in ONE launch for a small batch (every env steps on a warp of its own, its CTA
            # encodes), which with the padded channels-last layout also writes the policy's next input
            self.envs.step_encode_small(actions, reward_out=self._reward, done_out=self._done_step,
                                        mask_out=self._mask, ring=self._obs, slot_index=self._nxt_dev,
                                        nhwc32_out=self._last_obs if dual else None)
            self._last_obs_fresh = dual
        else:
            self.envs.step(actions, check_actions=False, mask_out=self._mask, reward_out=self._reward,
                           done_out=self._done_step)
            if dual:
                self.envs.obs_ring(self._obs, self._nxt_dev, nhwc32_out=self._last_obs)
                self._last_obs_fresh = True
            else:
                self.envs.obs_ring(self._obs, self._nxt_dev)
        a_p, r_p, d_p, m_p = self._slot_ptrs()
        phase = C.c_void_p(self._in_phase1.data_ptr()) if self.phase1_policy is not None else None
        check(L.pb_replay_record(n, r, C.c_void_p(self._t_dev.data_ptr()), C.c_void_p(actions.data_ptr()),
                                 C.c_void_p(self._reward.data_ptr()), C.c_void_p(self._done_step.data_ptr()),
                                 C.c_void_p(self._mask.data_ptr()), phase, a_p, r_p, d_p, m_p,
                                 C.c_void_p(self._valid.data_ptr()), dev, stream))
        if self.phase1_policy is not None:
            self._in_phase1.copy_((self._in_phase1 | self._done_step) & ~self.envs.first_ai_done())
        check(L.pb_replay_advance(C.c_void_p(self._t_dev.data_ptr()), C.c_void_p(self._nxt_dev.data_ptr()),
                                  r, dev, stream))
        if count_on_host:
            self._t += 1

    def _batch_buffers(self, b: int):
        bufs = self._batch_bufs.get(b)
        if bufs is None:
            dev = self.device
            mk = (lambda: _channels_last_empty((b,) + self._batch_shape, dev)) if self.channels_last else (
                lambda: torch.empty((b,) + OBS_SHAPE, dtype=torch.float32, device=dev))
            bufs = dict(idx=torch.zeros(b, dtype=torch.int64, device=dev), obs=mk(), next_obs=mk(),
                        actions=torch.zeros(b, dtype=torch.int64, device=dev),
                        rewards=torch.zeros(b, dtype=torch.float32, device=dev),
                        done=torch.zeros(b, dtype=torch.bool, device=dev),
                        next_mask=torch.zeros((b, NUM_ACTIONS), dtype=torch.bool, device=dev))
            self._batch_bufs[b] = bufs
        return bufs

    @torch.no_grad()
    def sample_indices(self, batch_size: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`batch_size` distinct flat ring indices (slot * N + env) of recorded transitions,
        uniform (random.sample).  `num_valid` afterwards holds the number of candidates."""
        total = self._ring * self.num_envs
        idx = self._batch_buffers(batch_size)["idx"] if out is None else out
        if total <= self._L.pb_replay_sample_capacity() and batch_size <= 4096:
            check(self._L.pb_replay_sample(C.c_void_p(self._valid.data_ptr()), total, batch_size,
                                           self.seed & 0xFFFFFFFFFFFFFFFF,
                                           C.c_void_p(self._sample_calls.data_ptr()),
                                           C.c_void_p(idx.data_ptr()), C.c_void_p(self._num_valid.data_ptr()),
                                           self.device.index, self._stream()))
            self._sample_calls += 1
        else:
            # ring too large for the shared-memory list of pb_replay_sample: exponential-race
            # top-k on the device (torch.multinomial without replacement)
            idx.copy_(torch.multinomial(self._valid.view(-1).to(torch.float32), batch_size, replacement=False))
            self._num_valid.copy_(self._valid.sum().to(torch.int32).reshape(1))
        return idx

    @property
    def num_valid(self) -> int:
        """Candidates the last `sample_indices` / `sample_batch` drew from (device read)."""
        return int(self._num_valid.item())

    @torch.no_grad()
    def sample_batch(self, batch_size: int) -> ReplayBatch:
        """Uniform sample without replacement + collate, two launches, all on device.  The
        returned tensors are per-batch-size buffers that the next call overwrites."""
        b = self._batch_buffers(batch_size)
        idx = self.sample_indices(batch_size)
        a_p, r_p, d_p, m_p = self._slot_ptrs()
        check(self._L.pb_replay_collate(
            C.c_void_p(self._obs.data_ptr()), self.num_envs, self._ring, C.c_void_p(idx.data_ptr()),
            batch_size, a_p, r_p, d_p, m_p, C.c_void_p(b["obs"].data_ptr()),
            C.c_void_p(b["next_obs"].data_ptr()), self._cl_arg,
            C.c_void_p(b["actions"].data_ptr()), C.c_void_p(b["rewards"].data_ptr()),
            C.c_void_p(b["done"].data_ptr()), C.c_void_p(b["next_mask"].data_ptr()),
            self.device.index, self._stream()))
        return ReplayBatch(obs=b["obs"], actions=b["actions"], rewards=b["rewards"],
                           next_obs=b["next_obs"], done=b["done"], next_action_masks=b["next_mask"])

    @property
    def current_obs(self) -> torch.Tensor:
        return self._obs[self._t % self._ring]

    @property
    def current_mask(self) -> torch.Tensor:
        return self._mask
