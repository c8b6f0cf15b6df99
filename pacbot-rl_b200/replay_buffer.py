"""GPU-resident replay buffer + experience generation: the batched mirror of the reference's
`ReplayBuffer` (/root/reference/src/replay_buffer.py:57-143) and `reset_env` (:44-54).

Layout: a time ring of R = ceil(maxlen / N) + 1 slots; slot t % R holds, for all N envs, the
observation BEFORE action t plus (action, reward, done, next_action_mask) of transition t.  The
observation encoder writes straight into the ring (`pb_encode_obs_ring`), so `next_obs` of
transition t is simply slot (t+1) % R and is never copied; a finished episode's `next_obs` is
"None" in the reference (zeros at collate time, train_dqn.py:154-159) and the slot then already
holds the first observation of the next episode (the env was reset in the same kernel launch).
One slot is always "current observation only", which is why R has the +1.  Sampling is uniform
without replacement over the valid transitions (`random.sample`, replay_buffer.py:141-143) and
the collate is two `pb_gather_obs` launches.

Everything that depends on the ring position uses a DEVICE-side step counter (`_t_dev`), never
a Python int, so `generate_experience_step` + `sample_batch` can be captured in a CUDA graph
and replayed while the ring advances."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from ._lib import NUM_ACTIONS, OBS_FLOATS, OBS_SHAPE, check
from .engine import BatchedPacmanGym
from .policies import Policy


class ReplayBatch(NamedTuple):
    """Collated transitions (what train_dqn.py:153-165 builds from a list of ReplayItems)."""

    obs: torch.Tensor                # float32 [B, 17, 28, 31]
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
    ) -> None:
        self.policy = policy
        self.phase1_policy = phase1_policy
        n = int(num_parallel_envs)
        self.num_envs = n
        self.maxlen = int(maxlen)
        self._slots = -(-self.maxlen // n)          # transitions kept: slots * n >= maxlen
        self._ring = self._slots + 1
        # replay_buffer.py:75-80: random_start for the first N * proportion envs, random_ticks all
        self.envs = BatchedPacmanGym(
            n, np.arange(n) < n * random_start_proportion, True, device=torch.device(device),
            seed=seed, env_id_base=env_id_base, auto_reset=True)
        dev = self.envs.device
        self.device = dev
        r = self._ring
        self._obs = torch.zeros((r, n) + OBS_SHAPE, dtype=torch.float32, device=dev)
        self._actions = torch.zeros((r, n), dtype=torch.int64, device=dev)
        self._rewards = torch.zeros((r, n), dtype=torch.int32, device=dev)
        self._done = torch.zeros((r, n), dtype=torch.bool, device=dev)
        self._next_mask = torch.zeros((r, n, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self._weight = torch.zeros((r, n), dtype=torch.float32, device=dev)  # 1 = sampleable
        self._mask = torch.zeros((n, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self._in_phase1 = torch.zeros(n, dtype=torch.bool, device=dev)
        self._t = 0                                                  # host mirror (len, fill)
        self._t_dev = torch.zeros(1, dtype=torch.int64, device=dev)  # the counter kernels use
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
        return min(self._t, self._slots) * self.num_envs

    @property
    def capacity(self) -> int:
        return self._slots * self.num_envs

    def fill(self) -> None:
        """Generate experience until the buffer holds `capacity` transitions."""
        while self._t < self._slots:
            self.generate_experience_step()

    @torch.no_grad()
    def generate_experience_step(self, count_on_host: bool = True) -> None:
        """One env-step for each parallel env, appended to the ring (replay_buffer.py:102-139).
        `count_on_host=False` is for CUDA-graph replay: the caller advances `_t` itself."""
        r, n = self._ring, self.num_envs
        slot = self._t_dev % r
        nxt = (self._t_dev + 1) % r
        last_obs = self._obs.view(r, -1).index_select(0, slot).view((n,) + OBS_SHAPE)
        actions = self.policy(last_obs, self._mask)
        if self.phase1_policy is not None:
            actions = torch.where(self._in_phase1, self.phase1_policy(last_obs, self._mask), actions)
        # step + auto reset, then the observation of the resulting state straight into slot t+1
        reward, done = self.envs.step(actions, check_actions=False, mask_out=self._mask)
        self.envs.obs_ring(self._obs, nxt)
        self._actions.index_copy_(0, slot, actions.to(torch.int64).unsqueeze(0))
        self._rewards.index_copy_(0, slot, reward.unsqueeze(0))
        self._done.index_copy_(0, slot, done.unsqueeze(0))
        # next_action_mask is [False]*5 when done (replay_buffer.py:115-117)
        self._next_mask.index_copy_(0, slot, (self._mask & ~done.unsqueeze(1)).unsqueeze(0))
        self._weight.index_copy_(0, slot, (~self._in_phase1).to(torch.float32).unsqueeze(0))
        self._weight.index_fill_(0, nxt, 0.0)                     # slot t+1 is "current obs" only
        if self.phase1_policy is not None:
            self._in_phase1 = (self._in_phase1 | done) & ~self.envs.first_ai_done()
        self._t_dev += 1
        if count_on_host:
            self._t += 1

    @torch.no_grad()
    def sample_batch(self, batch_size: int) -> ReplayBatch:
        """Uniform sample without replacement + collate, all on device."""
        r, n = self._ring, self.num_envs
        flat = torch.multinomial(self._weight.view(-1), batch_size, replacement=False)
        slot = flat // n
        env = flat - slot * n
        nxt_flat = ((slot + 1) % r) * n + env
        done = self._done.view(-1)[flat]
        obs = torch.empty((batch_size,) + OBS_SHAPE, dtype=torch.float32, device=self.device)
        next_obs = torch.empty_like(obs)
        L = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        ring_ptr = C.c_void_p(self._obs.data_ptr())
        check(L.pb_gather_obs(ring_ptr, OBS_FLOATS, C.c_void_p(flat.data_ptr()), None, batch_size,
                              C.c_void_p(obs.data_ptr()), self.device.index, stream))
        check(L.pb_gather_obs(ring_ptr, OBS_FLOATS, C.c_void_p(nxt_flat.data_ptr()),
                              C.c_void_p(done.view(torch.uint8).data_ptr()), batch_size,
                              C.c_void_p(next_obs.data_ptr()), self.device.index, stream))
        return ReplayBatch(
            obs=obs,
            actions=self._actions.view(-1)[flat],
            rewards=self._rewards.view(-1)[flat].to(torch.float32),
            next_obs=next_obs,
            done=done,
            next_action_masks=self._next_mask.view(-1, NUM_ACTIONS)[flat],
        )

    @property
    def current_obs(self) -> torch.Tensor:
        return self._obs[self._t % self._ring]

    @property
    def current_mask(self) -> torch.Tensor:
        return self._mask
