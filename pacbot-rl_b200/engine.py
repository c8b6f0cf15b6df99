"""Batched, GPU-resident PacBot environments: the host-side mirror of the reference's
`PacmanGym` (pacbot_rs/src/env.rs:130-359, pacbot_rs/pacbot_rs.pyi:6-19) over N envs at once.

All tensors live on the engine's CUDA device and every call is enqueued on the caller's current
torch stream.  PyTorch is plumbing here (memory + streams); the work is done by the sm_100a
kernels behind the C ABI in include/pacbot_b200.h."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import (
    CFG_RANDOM_START,
    CFG_RANDOM_TICKS,
    CFG_RANDOMIZE_GHOSTS,
    CFG_REMOVE_SUPER_PELLETS,
    ENV_STATE_DTYPE,
    NUM_ACTIONS,
    OBS_FLOATS,
    OBS_SHAPE,
    check,
)

BoolOrSeq = Union[bool, Sequence[bool], np.ndarray]


def _per_env_flags(n: int, random_start, random_ticks, randomize_ghosts, remove_super_pellets):
    def expand(v):
        a = np.asarray(v, dtype=bool)
        if a.ndim == 0:
            a = np.full(n, bool(a))
        if a.shape != (n,):
            raise ValueError(f"per-env flag must be a bool or have shape ({n},)")
        return a

    f = np.zeros(n, np.uint8)
    f |= expand(random_start).astype(np.uint8) * CFG_RANDOM_START
    f |= expand(random_ticks).astype(np.uint8) * CFG_RANDOM_TICKS
    f |= expand(randomize_ghosts).astype(np.uint8) * CFG_RANDOMIZE_GHOSTS
    f |= expand(remove_super_pellets).astype(np.uint8) * CFG_REMOVE_SUPER_PELLETS
    return f


class BatchedPacmanGym:
    """N independent `PacmanGym`s stepped by one kernel launch.

    Same semantics per env as the reference class; batched signatures:
      reset(mask=None), step(actions) -> (reward int32[N], done bool[N]),
      action_mask() -> bool[N,5], obs(out=None) -> float32[N,17,28,31], scalar getters -> [N].
    """

    def __init__(
        self,
        num_envs: int,
        random_start: BoolOrSeq = False,
        random_ticks: BoolOrSeq = False,
        *,
        device: Union[None, int, str, torch.device] = None,
        seed: int = 0,
        env_id_base: int = 0,
        randomize_ghosts: BoolOrSeq = False,
        remove_super_pellets: BoolOrSeq = False,
        auto_reset: bool = False,
    ) -> None:
        L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PacbotB200Error(
                "no CUDA device: the pacbot_b200 engine has no CPU path (use a B200)"
            )
        dev = torch.device("cuda" if device is None else device)
        if dev.type != "cuda":
            raise ValueError("BatchedPacmanGym needs a CUDA device")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._dev_index = int(dev.index)
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self._cfg = _per_env_flags(
            self.num_envs, random_start, random_ticks, randomize_ghosts, remove_super_pellets
        )
        self._L = L
        self._h = C.c_void_p()
        self._tape = None
        self._sample_calls = 0
        self._select_calls = 0
        check(
            L.pb_create(
                C.byref(self._h),
                self.num_envs,
                dev.index,
                self.seed & 0xFFFFFFFFFFFFFFFF,
                self.env_id_base,
                self._cfg.ctypes.data_as(C.c_void_p),
            )
        )

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.pb_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self.num_envs

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> C.c_void_p:
        # the raw handle of torch's current stream on the engine's device (the private accessor is
        # ~10x cheaper than building a torch.cuda.Stream object per call; same handle)
        try:
            return C.c_void_p(torch._C._cuda_getCurrentRawStream(self._dev_index))
        except AttributeError:
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _new(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _check_tensor(self, t: torch.Tensor, shape, dtype, name: str) -> torch.Tensor:
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape):
            raise ValueError(
                f"{name}: expected {dtype} tensor of shape {tuple(shape)} on {self.device}, "
                f"got {t.dtype} {tuple(t.shape)} on {t.device}"
            )
        if not t.is_contiguous():
            raise ValueError(f"{name}: must be contiguous")
        return t

    def _actions_u8(self, actions) -> torch.Tensor:
        if not isinstance(actions, torch.Tensor):
            a = np.asarray(actions)
            if a.shape != (self.num_envs,):
                raise ValueError(f"actions: expected shape ({self.num_envs},)")
            if a.dtype.kind not in "iu":
                raise TypeError("actions must be integers")
            if ((a < 0) | (a > 255)).any():
                # pyo3 `u8` extraction failure (env.rs:85)
                raise OverflowError("out of range integral type conversion attempted")
            return torch.from_numpy(a.astype(np.uint8)).to(self.device, non_blocking=True)
        if tuple(actions.shape) != (self.num_envs,):
            raise ValueError(f"actions: expected shape ({self.num_envs},)")
        if actions.dtype.is_floating_point or actions.dtype == torch.bool:
            raise TypeError("actions must be an integer tensor")
        a = actions.to(self.device)
        if a.dtype != torch.uint8:
            # values outside 0..4 map to 255 so the kernel reports them as invalid
            a = torch.where((a < 0) | (a > 4), torch.full_like(a, 255), a).to(torch.uint8)
        return a.contiguous()

    # ------------------------------------------------------------------ configuration
    @property
    def random_start(self) -> np.ndarray:
        """Per-env `random_start` flag (the reference exposes it get+set, env.rs:100-101)."""
        return (self._cfg & CFG_RANDOM_START).astype(bool)

    @random_start.setter
    def random_start(self, value: BoolOrSeq) -> None:
        v = np.asarray(value, dtype=bool)
        if v.ndim == 0:
            v = np.full(self.num_envs, bool(v))
        self._cfg = (self._cfg & ~np.uint8(CFG_RANDOM_START)) | v.astype(np.uint8)
        check(
            self._L.pb_set_cfg_flags(
                self._h, 0, self.num_envs, self._cfg.ctypes.data_as(C.c_void_p)
            )
        )

    def set_draw_tape(self, tape: Optional[torch.Tensor]) -> None:
        """Install (or remove) a tape of injected raw draws: uint32-valued [N, L] tensor.
        Draw number c of env i becomes tape[i, c % L] (parity tests replay the oracle's draws)."""
        if tape is None:
            self._tape = None
            check(self._L.pb_set_draw_tape(self._h, None, 0))
            return
        if tape.dtype not in (torch.int32, torch.uint32) or tape.ndim != 2:
            raise ValueError("tape must be a 2-D int32/uint32 tensor holding the raw u32 bits")
        if tape.shape[0] != self.num_envs:
            raise ValueError("tape must have one row per env")
        self._tape = tape.to(self.device).contiguous()
        check(self._L.pb_set_draw_tape(self._h, C.c_void_p(self._tape.data_ptr()), tape.shape[1]))

    # ------------------------------------------------------------------ hot path
    def reset(self, mask: Optional[torch.Tensor] = None) -> None:
        """PacmanGym.reset() for every env (or those where `mask` is True)."""
        m = None
        if mask is not None:
            m = mask.to(self.device)
            if m.dtype == torch.bool:
                m = m.view(torch.uint8)
            m = self._check_tensor(m.contiguous(), (self.num_envs,), torch.uint8, "mask")
        check(self._L.pb_reset(self._h, None if m is None else C.c_void_p(m.data_ptr()), self._stream()))

    def step(self, actions, *, check_actions: bool = True, mask_out: Optional[torch.Tensor] = None,
             reward_out: Optional[torch.Tensor] = None, done_out: Optional[torch.Tensor] = None,
             flip: bool = False):
        """PacmanGym.step for all envs -> (reward int32[N], done bool[N]).

        check_actions=True synchronises and raises ValueError("Invalid action") like the
        reference; pass False inside device-resident loops."""
        a = self._actions_u8(actions)
        n = self.num_envs
        reward = self._new((n,), torch.int32) if reward_out is None else self._check_tensor(
            reward_out, (n,), torch.int32, "reward_out")
        done = self._new((n,), torch.uint8) if done_out is None else self._check_tensor(
            done_out.view(torch.uint8), (n,), torch.uint8, "done_out")
        mptr = None
        if mask_out is not None:
            mptr = C.c_void_p(self._check_tensor(
                mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        if flip:  # double-buffered: the state just read stays intact (rollout.py)
            check(self._L.pb_step_flip(self._h, C.c_void_p(a.data_ptr()), 0, None,
                                       C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()), mptr,
                                       int(self.auto_reset), self._stream()))
        else:
            check(self._L.pb_step(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(reward.data_ptr()),
                                  C.c_void_p(done.data_ptr()), mptr, int(self.auto_reset),
                                  self._stream()))
        if check_actions:
            self.check_actions()
        return reward, done.view(torch.bool)

    def step_random(self, *, call_idx: Optional[int] = None, actions_out: Optional[torch.Tensor] = None,
                    mask_out: Optional[torch.Tensor] = None, reward_out: Optional[torch.Tensor] = None,
                    done_out: Optional[torch.Tensor] = None, first: int = 0,
                    count: Optional[int] = None, flip: bool = False):
        """sample_actions() + step() fused into one launch (random-policy rollouts).  Returns
        (reward, done); the actions taken land in `actions_out` (uint8 [N]) if given.
        `first` / `count` restrict the launch to envs [first, first+count); the arrays stay
        full size and are indexed by the absolute env index."""
        n = self.num_envs
        if count is None:
            count = n - first
        reward = self._new((n,), torch.int32) if reward_out is None else self._check_tensor(
            reward_out, (n,), torch.int32, "reward_out")
        done = self._new((n,), torch.uint8) if done_out is None else self._check_tensor(
            done_out.view(torch.uint8), (n,), torch.uint8, "done_out")
        aptr = None if actions_out is None else C.c_void_p(self._check_tensor(
            actions_out, (n,), torch.uint8, "actions_out").data_ptr())
        mptr = None if mask_out is None else C.c_void_p(self._check_tensor(
            mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        if call_idx is None:
            call_idx = self._sample_calls
            self._sample_calls += 1
        if flip:
            if first != 0 or count != n:
                raise ValueError("flip=True steps every env")
            check(self._L.pb_step_random_flip(self._h, int(call_idx), aptr, C.c_void_p(reward.data_ptr()),
                                              C.c_void_p(done.data_ptr()), mptr, int(self.auto_reset),
                                              self._stream()))
        else:
            check(self._L.pb_step_random_range(self._h, int(first), int(count), int(call_idx), aptr,
                                               C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                               mptr, int(self.auto_reset), self._stream()))
        return reward, done.view(torch.bool)

    # ------------------------------------------------------------------ pipelined rollout steps
    def rollout_random_step(self, obs_out: Optional[torch.Tensor], *, call_idx: Optional[int] = None,
                            actions_out: Optional[torch.Tensor] = None,
                            mask_out: Optional[torch.Tensor] = None,
                            reward_out: Optional[torch.Tensor] = None,
                            done_out: Optional[torch.Tensor] = None):
        """step_random() + obs() as ONE pipelined C call (pb_rollout_random_step): the step kernel
        of this call overlaps the encoder of the previous call on engine-owned streams
        (double-buffered state).  Outputs are complete after rollout_join()."""
        n = self.num_envs
        reward = self._new((n,), torch.int32) if reward_out is None else self._check_tensor(
            reward_out, (n,), torch.int32, "reward_out")
        done = self._new((n,), torch.uint8) if done_out is None else self._check_tensor(
            done_out.view(torch.uint8), (n,), torch.uint8, "done_out")
        aptr = None if actions_out is None else C.c_void_p(self._check_tensor(
            actions_out, (n,), torch.uint8, "actions_out").data_ptr())
        mptr = None if mask_out is None else C.c_void_p(self._check_tensor(
            mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        optr = None if obs_out is None else C.c_void_p(self._check_tensor(
            obs_out, (n,) + OBS_SHAPE, torch.float32, "obs_out").data_ptr())
        if call_idx is None:
            call_idx = self._sample_calls
            self._sample_calls += 1
        check(self._L.pb_rollout_random_step(self._h, int(call_idx), aptr, C.c_void_p(reward.data_ptr()),
                                             C.c_void_p(done.data_ptr()), mptr, int(self.auto_reset),
                                             optr, OBS_FLOATS, self._stream()))
        return reward, done.view(torch.bool)

    def rollout_host_step(self, actions: np.ndarray, obs_out: Optional[torch.Tensor], out: tuple):
        """step_host() as one pipelined C call (pb_rollout_host_step): `actions` uint8 [N] host
        array (pinned for full speed), `out` = (reward int32[N], done 1-byte[N], mask 1-byte[N,5]
        or None) host arrays filled when the call returns; the observation lands in the device
        tensor `obs_out` and is complete after rollout_join() (or once a later call returned)."""
        n = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        reward, done, mask = out
        if (a.shape != (n,) or reward.shape != (n,) or reward.dtype != np.int32 or done.shape != (n,)
                or done.dtype.itemsize != 1 or (mask is not None and (mask.shape != (n, NUM_ACTIONS)
                                                                      or mask.dtype.itemsize != 1))):
            raise ValueError("rollout_host_step: expected uint8[N] actions and (int32[N], 1-byte[N], "
                             "1-byte[N,5] or None) outputs")
        optr = None if obs_out is None else C.c_void_p(self._check_tensor(
            obs_out, (n,) + OBS_SHAPE, torch.float32, "obs_out").data_ptr())
        check(self._L.pb_rollout_host_step(
            self._h, a.ctypes.data_as(C.c_void_p), reward.ctypes.data_as(C.c_void_p),
            done.ctypes.data_as(C.c_void_p), None if mask is None else mask.ctypes.data_as(C.c_void_p),
            int(self.auto_reset), optr, OBS_FLOATS, self._stream()))

    def rollout_chunk_step(self, first: int, count: int, obs_out: Optional[torch.Tensor], *,
                           actions: Optional[torch.Tensor] = None, call_idx: Optional[int] = None,
                           actions_out: Optional[torch.Tensor] = None,
                           mask_out: Optional[torch.Tensor] = None,
                           reward_out: Optional[torch.Tensor] = None,
                           done_out: Optional[torch.Tensor] = None):
        """One chunk of the in-place chunked rollout (pb_rollout_chunk_step): step envs
        [first, first+count) (random valid actions unless `actions`, uint8 [N], is given) and
        encode their observations into `obs_out` (float32 [count,17,28,31]) on engine-owned
        streams; the step of a chunk overlaps the encoders of the other chunks.  The arrays stay
        full size and are indexed by the absolute env index.  Complete after rollout_join()."""
        n = self.num_envs
        if first < 0 or count < 0 or first + count > n:
            raise ValueError("rollout_chunk_step: env range out of bounds")
        reward = self._new((n,), torch.int32) if reward_out is None else self._check_tensor(
            reward_out, (n,), torch.int32, "reward_out")
        done = self._new((n,), torch.uint8) if done_out is None else self._check_tensor(
            done_out.view(torch.uint8), (n,), torch.uint8, "done_out")
        aptr = None if actions_out is None else C.c_void_p(self._check_tensor(
            actions_out, (n,), torch.uint8, "actions_out").data_ptr())
        iptr = None if actions is None else C.c_void_p(self._check_tensor(
            actions, (n,), torch.uint8, "actions").data_ptr())
        mptr = None if mask_out is None else C.c_void_p(self._check_tensor(
            mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        optr = None if obs_out is None else C.c_void_p(self._check_tensor(
            obs_out, (count,) + OBS_SHAPE, torch.float32, "obs_out").data_ptr())
        if call_idx is None:
            call_idx = self._sample_calls
        check(self._L.pb_rollout_chunk_step(self._h, int(first), int(count), iptr, int(call_idx), aptr,
                                            C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                            mptr, int(self.auto_reset), optr, OBS_FLOATS, self._stream()))
        return reward, done.view(torch.bool)

    def rollout_chunk_host_step(self, first: int, count: int, actions: np.ndarray,
                                obs_out: Optional[torch.Tensor], out: tuple) -> None:
        """rollout_chunk_step with HOST buffers (pb_rollout_chunk_host_step): `actions` uint8 [N]
        and `out` = (reward int32[N], done 1-byte[N], mask 1-byte[N,5] or None) are full-size host
        arrays (pinned for asynchronous copies); only [first, first+count) is read / written.
        Returns at once; the results are on the host after rollout_host_sync()."""
        n = self.num_envs
        reward, done, mask = out
        if (not isinstance(actions, np.ndarray) or actions.dtype != np.uint8 or actions.shape != (n,)
                or not actions.flags.c_contiguous or reward.shape != (n,) or reward.dtype != np.int32
                or done.shape != (n,) or done.dtype.itemsize != 1
                or (mask is not None and (mask.shape != (n, NUM_ACTIONS) or mask.dtype.itemsize != 1))):
            raise ValueError("rollout_chunk_host_step: expected contiguous uint8[N] actions and "
                             "(int32[N], 1-byte[N], 1-byte[N,5] or None) outputs")
        if first < 0 or count < 0 or first + count > n:
            raise ValueError("rollout_chunk_host_step: env range out of bounds")
        optr = None if obs_out is None else C.c_void_p(self._check_tensor(
            obs_out, (count,) + OBS_SHAPE, torch.float32, "obs_out").data_ptr())
        check(self._L.pb_rollout_chunk_host_step(
            self._h, int(first), int(count), actions.ctypes.data_as(C.c_void_p),
            reward.ctypes.data_as(C.c_void_p), done.ctypes.data_as(C.c_void_p),
            None if mask is None else mask.ctypes.data_as(C.c_void_p), int(self.auto_reset), optr,
            OBS_FLOATS, self._stream()))

    def rollout_host_sync(self) -> None:
        """Block until the results of every rollout_chunk_host_step issued so far are on the host;
        raises ValueError("Invalid action") if one of them carried an action > 4."""
        check(self._L.pb_rollout_host_sync(self._h))

    def rollout_join(self) -> None:
        """Make the current stream wait for every pipelined rollout step enqueued so far."""
        check(self._L.pb_rollout_join(self._h, self._stream()))

    def rollout_streams(self):
        """(step_stream, encode_stream) of the pipelined rollout as torch ExternalStreams."""
        s, e = C.c_void_p(), C.c_void_p()
        check(self._L.pb_rollout_streams(self._h, C.byref(s), C.byref(e)))
        return (torch.cuda.ExternalStream(s.value, device=self.device),
                torch.cuda.ExternalStream(e.value, device=self.device))

    def check_actions(self) -> None:
        bad = C.c_int64(-1)
        check(self._L.pb_check_actions(self._h, C.byref(bad), self._stream()))

    def obs(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """PacmanGym.obs() for all envs: float32 [N, 17, 28, 31] (written in place if `out`)."""
        n = self.num_envs
        if out is None:
            out = self._new((n,) + OBS_SHAPE, torch.float32)
        else:
            self._check_tensor(out, (n,) + OBS_SHAPE, torch.float32, "out")
        check(self._L.pb_encode_obs(self._h, C.c_void_p(out.data_ptr()), OBS_FLOATS, self._stream()))
        return out

    def obs_range(self, first: int, count: int, out: torch.Tensor) -> torch.Tensor:
        """obs() of envs [first, first+count) into `out` (float32 [count, 17, 28, 31])."""
        self._check_tensor(out, (count,) + OBS_SHAPE, torch.float32, "out")
        check(self._L.pb_encode_obs_range(self._h, first, count, C.c_void_p(out.data_ptr()),
                                          OBS_FLOATS, self._stream()))
        return out

    def obs_nhwc32(self, out: Optional[torch.Tensor] = None, first: int = 0,
                   count: Optional[int] = None) -> torch.Tensor:
        """obs() of envs [first, first+count) as the networks take it: float32 [count, 32, 28, 31]
        in channels_last memory format, channels 17..31 zero (pb_encode_obs_nhwc32: the encoder
        writes this layout itself -- no layout-conversion launch between it and the first
        convolution).  `out`: a tensor of that shape and memory format to fill."""
        if count is None:
            count = self.num_envs - first
        shape = (count, 32) + OBS_SHAPE[1:]
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device,
                              memory_format=torch.channels_last)
        elif (out.dtype != torch.float32 or out.device != self.device or tuple(out.shape) != shape
              or not out.is_contiguous(memory_format=torch.channels_last)):
            raise ValueError(f"out: expected a channels_last float32 tensor of shape {shape} on {self.device}")
        check(self._L.pb_encode_obs_nhwc32(self._h, first, count, C.c_void_p(out.data_ptr()),
                                           self._stream()))
        return out

    def obs_ring(self, ring: torch.Tensor, slot_index: torch.Tensor,
                 nhwc32_out: Optional[torch.Tensor] = None) -> None:
        """obs() of all envs into ring[slot_index] where `slot_index` is a 1-element int64 DEVICE
        tensor read at kernel time (CUDA-graph friendly).  ring: float32 [R, N, 17, 28, 31].
        `nhwc32_out` (float32 [N, 32, 28, 31], channels_last): the same observations also in the
        networks' input layout, from the same launch (pb_encode_obs_ring_nhwc32)."""
        if (ring.dtype != torch.float32 or ring.device != self.device or not ring.is_contiguous()
                or tuple(ring.shape[1:]) != (self.num_envs,) + OBS_SHAPE):
            raise ValueError("ring: expected contiguous float32 [R, N, 17, 28, 31] on the engine device")
        if slot_index.dtype != torch.int64 or slot_index.device != self.device or slot_index.numel() != 1:
            raise ValueError("slot_index: expected a 1-element int64 tensor on the engine device")
        if nhwc32_out is not None:
            shape = (self.num_envs, 32) + OBS_SHAPE[1:]
            if (nhwc32_out.dtype != torch.float32 or nhwc32_out.device != self.device
                    or tuple(nhwc32_out.shape) != shape
                    or not nhwc32_out.is_contiguous(memory_format=torch.channels_last)):
                raise ValueError(f"nhwc32_out: expected a channels_last float32 tensor of shape {shape}")
            check(self._L.pb_encode_obs_ring_nhwc32(
                self._h, C.c_void_p(ring.data_ptr()), self.num_envs * OBS_FLOATS,
                C.c_void_p(slot_index.data_ptr()), C.c_void_p(nhwc32_out.data_ptr()), self._stream()))
            return
        check(self._L.pb_encode_obs_ring(self._h, C.c_void_p(ring.data_ptr()),
                                         self.num_envs * OBS_FLOATS,
                                         C.c_void_p(slot_index.data_ptr()), self._stream()))

    SMALL_BATCH = 256   # up to here a launch is "small": one CTA per env (csrc/pb_encode_small.cuh)

    def step_encode_small(self, actions, *, reward_out: torch.Tensor, done_out: torch.Tensor,
                          mask_out: Optional[torch.Tensor] = None, obs_out: Optional[torch.Tensor] = None,
                          ring: Optional[torch.Tensor] = None, slot_index: Optional[torch.Tensor] = None,
                          nhwc32_out: Optional[torch.Tensor] = None) -> None:
        """step() and the observation(s) of the resulting states in ONE launch, one CTA per env
        (pb_step_encode_small; meant for engines of up to SMALL_BATCH envs: the DQN loop's
        experience step, an evaluation step).  Observation outputs, any subset: `obs_out` float32
        [N, 17, 28, 31]; or `ring` [R, N, 17, 28, 31] + `slot_index` (1-element int64 device tensor
        read at kernel time); `nhwc32_out` [N, 32, 28, 31] channels_last.  No action check here:
        an action > 4 leaves its env untouched and is reported by check_actions()."""
        a = self._actions_u8(actions)
        n = self.num_envs
        self._check_tensor(reward_out, (n,), torch.int32, "reward_out")
        done = self._check_tensor(done_out.view(torch.uint8), (n,), torch.uint8, "done_out")
        mptr = None
        if mask_out is not None:
            mptr = C.c_void_p(self._check_tensor(
                mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        optr, stride, sptr, sstride = None, 0, None, 0
        if ring is not None:
            if (obs_out is not None or slot_index is None or ring.dtype != torch.float32
                    or ring.device != self.device or not ring.is_contiguous()
                    or tuple(ring.shape[1:]) != (n,) + OBS_SHAPE or slot_index.dtype != torch.int64
                    or slot_index.device != self.device or slot_index.numel() != 1):
                raise ValueError("ring: contiguous float32 [R, N, 17, 28, 31] + a 1-element int64 slot_index "
                                 "on the engine device (and no obs_out)")
            optr, stride = C.c_void_p(ring.data_ptr()), OBS_FLOATS
            sptr, sstride = C.c_void_p(slot_index.data_ptr()), n * OBS_FLOATS
        elif obs_out is not None:
            self._check_tensor(obs_out, (n,) + OBS_SHAPE, torch.float32, "obs_out")
            optr, stride = C.c_void_p(obs_out.data_ptr()), OBS_FLOATS
        nptr = None
        if nhwc32_out is not None:
            shape = (n, 32) + OBS_SHAPE[1:]
            if (nhwc32_out.dtype != torch.float32 or nhwc32_out.device != self.device
                    or tuple(nhwc32_out.shape) != shape
                    or not nhwc32_out.is_contiguous(memory_format=torch.channels_last)):
                raise ValueError(f"nhwc32_out: expected a channels_last float32 tensor of shape {shape}")
            nptr = C.c_void_p(nhwc32_out.data_ptr())
        check(self._L.pb_step_encode_small(
            self._h, C.c_void_p(a.data_ptr()), C.c_void_p(reward_out.data_ptr()), C.c_void_p(done.data_ptr()),
            mptr, int(self.auto_reset), optr, stride, sptr, sstride, nptr, self._stream()))

    def step_obs(self, actions, obs_out: torch.Tensor, *, check_actions: bool = False,
                 mask_out: Optional[torch.Tensor] = None):
        """step() then obs() of the resulting states, back to back (one C-ABI call)."""
        a = self._actions_u8(actions)
        n = self.num_envs
        self._check_tensor(obs_out, (n,) + OBS_SHAPE, torch.float32, "obs_out")
        reward = self._new((n,), torch.int32)
        done = self._new((n,), torch.uint8)
        mptr = None
        if mask_out is not None:
            mptr = C.c_void_p(self._check_tensor(
                mask_out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask_out").data_ptr())
        check(self._L.pb_step_encode(self._h, C.c_void_p(a.data_ptr()),
                                     C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()),
                                     mptr, int(self.auto_reset), C.c_void_p(obs_out.data_ptr()),
                                     OBS_FLOATS, self._stream()))
        if check_actions:
            self.check_actions()
        return reward, done.view(torch.bool)

    def action_mask(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """PacmanGym.action_mask() for all envs: bool [N, 5]."""
        n = self.num_envs
        m = self._new((n, NUM_ACTIONS), torch.uint8) if out is None else self._check_tensor(
            out.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "out")
        check(self._L.pb_action_mask(self._h, C.c_void_p(m.data_ptr()), self._stream()))
        return m.view(torch.bool)

    def _query(self, what: int) -> torch.Tensor:
        out = self._new((self.num_envs,), torch.int32)
        check(self._L.pb_query(self._h, what, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def score(self) -> torch.Tensor:
        return self._query(_lib.Q_SCORE)

    def lives(self) -> torch.Tensor:
        return self._query(_lib.Q_LIVES)

    def is_done(self) -> torch.Tensor:
        return self._query(_lib.Q_IS_DONE).bool()

    def first_ai_done(self) -> torch.Tensor:
        return self._query(_lib.Q_FIRST_AI_DONE).bool()

    def remaining_pellets(self) -> torch.Tensor:
        return self._query(_lib.Q_REMAINING_PELLETS)

    def ticks_per_step(self) -> torch.Tensor:
        return self._query(_lib.Q_TICKS_PER_STEP)

    def sample_actions(self, out: Optional[torch.Tensor] = None,
                       call_idx: Optional[int] = None) -> torch.Tensor:
        """Uniformly random valid action per env (EpsilonGreedy at epsilon = 1)."""
        a = self._new((self.num_envs,), torch.uint8) if out is None else self._check_tensor(
            out, (self.num_envs,), torch.uint8, "out")
        if call_idx is None:
            call_idx = self._sample_calls
            self._sample_calls += 1
        check(self._L.pb_sample_actions(self._h, C.c_void_p(a.data_ptr()), int(call_idx),
                                        self._stream()))
        return a

    def select_actions(self, q_values: torch.Tensor, epsilon: float,
                       mask: Optional[torch.Tensor] = None,
                       call_idx: Optional[int] = None) -> torch.Tensor:
        """EpsilonGreedy(MaxQPolicy) on device (src/policies.py:16-59): uint8 [N]."""
        n = self.num_envs
        q = self._check_tensor(q_values.contiguous(), (n, NUM_ACTIONS), torch.float32, "q_values")
        mptr = None
        if mask is not None:
            mptr = C.c_void_p(self._check_tensor(
                mask.contiguous().view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask").data_ptr())
        a = self._new((n,), torch.uint8)
        if call_idx is None:
            call_idx = self._select_calls
            self._select_calls += 1
        check(self._L.pb_select_actions(self._h, C.c_void_p(q.data_ptr()), mptr, float(epsilon),
                                        C.c_void_p(a.data_ptr()), int(call_idx), self._stream()))
        return a

    def select_actions_dev(self, q_values: torch.Tensor, epsilon: torch.Tensor,
                           call_idx: torch.Tensor, mask: Optional[torch.Tensor] = None,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """select_actions() with `epsilon` (float32, 1 element) and `call_idx` (int64, 1 element)
        read from DEVICE memory at kernel time (pb_select_actions_dev): CUDA-graph friendly."""
        n = self.num_envs
        q = self._check_tensor(q_values, (n, NUM_ACTIONS), torch.float32, "q_values")
        if (epsilon.dtype != torch.float32 or epsilon.numel() != 1 or epsilon.device != self.device
                or call_idx.dtype != torch.int64 or call_idx.numel() != 1 or call_idx.device != self.device):
            raise ValueError("epsilon: 1-element float32, call_idx: 1-element int64, both on the engine device")
        mptr = None
        if mask is not None:
            mptr = C.c_void_p(self._check_tensor(
                mask.view(torch.uint8), (n, NUM_ACTIONS), torch.uint8, "mask").data_ptr())
        a = self._new((n,), torch.uint8) if out is None else self._check_tensor(
            out, (n,), torch.uint8, "out")
        check(self._L.pb_select_actions_dev(self._h, C.c_void_p(q.data_ptr()), mptr,
                                            C.c_void_p(epsilon.data_ptr()), C.c_void_p(call_idx.data_ptr()),
                                            C.c_void_p(a.data_ptr()), self._stream()))
        return a

    # ------------------------------------------------------------------ host-buffer entry
    def step_host(self, actions: np.ndarray, obs_out: Optional[torch.Tensor] = None,
                  want_mask: bool = True, out: Optional[tuple] = None, begin_only: bool = False):
        """The per-call FFI shape: numpy actions in, numpy (reward, done, mask) out, with the
        host<->device copies inside the call.  The observation, if requested, is written to the
        DEVICE tensor `obs_out` (it feeds the Q-network / replay ring there).

        `out=(reward int32[N], done bool|uint8[N], mask bool|uint8[N,5] or None)` reuses caller
        buffers (pin them for full-speed D2H); otherwise fresh pageable arrays are returned.

        `begin_only=True` (pb_step_host_begin; no `obs_out`): enqueue the copies and the step and
        return at once; the arrays are filled when `step_host_end()` returns.  Whatever the caller
        enqueues in between (chunked `obs_range` encodes) overlaps the result copies."""
        if begin_only and obs_out is not None:
            raise ValueError("begin_only: encode the observation yourself between begin and end")
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        n = self.num_envs
        if a.shape != (n,):
            raise ValueError(f"actions: expected shape ({n},)")
        if out is None:
            reward = np.empty(n, np.int32)
            done = np.empty(n, np.bool_)
            mask = np.empty((n, NUM_ACTIONS), np.bool_) if want_mask else None
        else:
            reward, done, mask = out
            if (reward.dtype != np.int32 or reward.shape != (n,) or done.itemsize != 1
                    or done.shape != (n,) or not reward.flags.c_contiguous
                    or not done.flags.c_contiguous):
                raise ValueError("out: expected (int32[N], 1-byte[N], 1-byte[N,5] or None)")
            if mask is not None and (mask.itemsize != 1 or mask.shape != (n, NUM_ACTIONS)
                                     or not mask.flags.c_contiguous):
                raise ValueError("out: mask must be a contiguous 1-byte [N,5] array")
        optr = None
        if obs_out is not None:
            optr = C.c_void_p(self._check_tensor(
                obs_out, (n,) + OBS_SHAPE, torch.float32, "obs_out").data_ptr())
        args = (self._h, a.ctypes.data_as(C.c_void_p), reward.ctypes.data_as(C.c_void_p),
                done.ctypes.data_as(C.c_void_p),
                None if mask is None else mask.ctypes.data_as(C.c_void_p), int(self.auto_reset))
        if begin_only:
            # the arrays must outlive the copies: kept until step_host_end()
            self._host_step = (a, reward, done, mask)
            check(self._L.pb_step_host_begin(*args, self._stream()))
        else:
            check(self._L.pb_step_host(*args, optr, OBS_FLOATS, self._stream()))
        return reward, done, mask

    def step_snapshot(self, actions: Optional[np.ndarray] = None, want_obs: bool = True) -> dict:
        """The per-env drop-in's round trip (pb_step_snapshot_host, engines of <= 1024 envs): step
        every env with the uint8 [N] host `actions` (None: no step) and read back, in ONE stream
        synchronisation, everything the reference's loop asks an env for afterwards.  Returns a
        dict of host arrays owned by the engine and overwritten by the next call: `reward` int32
        [N] / `done` bool [N] (only after a step), `info` int32 [N, 8] (score, lives, is_done,
        first_ai_done, remaining_pellets, ticks_per_step, level, curr_ticks), `mask` bool [N, 5],
        `obs` float32 [N, 17, 28, 31] (if want_obs)."""
        b = getattr(self, "_snap_bufs", None)
        if b is None:
            # numpy views of the engine's own pinned block (pb_snapshot_block): the kernel reads the
            # actions from, and writes every answer into, this memory -- no host-side copies
            n = self.num_envs
            blk, act, offs = C.c_void_p(), C.c_void_p(), (C.c_size_t * 6)()
            check(self._L.pb_snapshot_block(self._h, C.byref(blk), C.byref(act), offs))

            def view(addr, shape, dtype):
                nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
                return np.frombuffer((C.c_char * nbytes).from_address(addr), dtype).reshape(shape)

            b = self._snap_bufs = {
                "info": view(blk.value + offs[0], (n, 8), np.int32),
                "reward": view(blk.value + offs[1], (n,), np.int32),
                "done": view(blk.value + offs[2], (n,), np.bool_),
                "mask": view(blk.value + offs[3], (n, NUM_ACTIONS), np.bool_),
                "obs": view(blk.value + offs[4], (n,) + OBS_SHAPE, np.float32),
                "actions": view(act.value, (n,), np.uint8)}
            self._snap_ptrs = {k: C.c_void_p(v.ctypes.data) for k, v in b.items()}
        p = self._snap_ptrs
        aptr = None
        if actions is not None:
            b["actions"][:] = actions
            aptr = p["actions"]
        check(self._L.pb_step_snapshot_host(
            self._h, aptr, int(self.auto_reset), p["reward"], p["done"], p["info"], p["mask"],
            p["obs"] if want_obs else None, self._stream()))
        return b

    def step_host_end(self):
        """Wait for the results of `step_host(..., begin_only=True)` -> (reward, done, mask)."""
        check(self._L.pb_step_host_end(self._h, self._stream()))
        out, self._host_step = getattr(self, "_host_step", None), None
        return None if out is None else out[1:]

    def obs_numpy(self, first: int = 0, count: Optional[int] = None,
                  out: Optional[np.ndarray] = None) -> np.ndarray:
        """obs_numpy() of envs [first, first+count): float32 [count, 17, 28, 31] on the host.
        `out`: a C-contiguous float32 array to fill instead of a fresh one; give it PINNED memory
        (e.g. `torch.empty(...).pin_memory().numpy()`) and the chunks cross PCIe at link speed
        while the next chunk is being encoded (pb_obs_host); a pageable array goes through the
        engine's pinned bounce buffers."""
        if count is None:
            count = self.num_envs - first
        if out is None:
            out = np.empty((count,) + OBS_SHAPE, np.float32)
        elif (out.dtype != np.float32 or out.shape != (count,) + OBS_SHAPE
              or not out.flags.c_contiguous):
            raise ValueError(f"out: expected a C-contiguous float32 array of shape {(count,) + OBS_SHAPE}")
        check(self._L.pb_obs_host(self._h, first, count, out.ctypes.data_as(C.c_void_p),
                                  self._stream()))
        return out

    # ------------------------------------------------------------------ state access
    def get_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Structured array (dtype `ENV_STATE_DTYPE` == C `pb_env_state`) of envs [first, +count)."""
        if count is None:
            count = self.num_envs - first
        out = np.zeros(count, ENV_STATE_DTYPE)
        # ordered on the caller's current stream (+ the engine's own streams): no device-wide
        # synchronisation, other streams of the process keep running
        check(self._L.pb_get_state_on(self._h, first, count, out.ctypes.data_as(C.c_void_p),
                                      self._stream()))
        return out

    def set_state(self, state: np.ndarray, first: int = 0) -> None:
        st = np.ascontiguousarray(state, dtype=ENV_STATE_DTYPE)
        check(self._L.pb_set_state_on(self._h, first, st.shape[0], st.ctypes.data_as(C.c_void_p),
                                      self._stream()))
        self._cfg[first:first + st.shape[0]] = st["cfg_flags"] & 0x0F

    # ------------------------------------------------------------------ device timeline
    TRACE_DTYPE = np.dtype([("begin_ns", np.uint64), ("ready_ns", np.uint64), ("end_ns", np.uint64),
                            ("kind", np.uint32), ("reserved", np.uint32)])

    def trace_begin(self, capacity: int = 4096) -> None:
        """Record %globaltimer begin / ready / end of every k_step / k_encode_obs launch of this
        engine from now on (pb_trace_begin; diagnostics, synchronises the device)."""
        check(self._L.pb_trace_begin(self._h, int(capacity)))

    def trace_mark(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """A one-thread timestamp kernel on `stream` (default: the current stream)."""
        s = self._stream() if stream is None else C.c_void_p(stream.cuda_stream)
        check(self._L.pb_trace_mark(self._h, s))

    def trace_end(self) -> np.ndarray:
        """Stop tracing; records in launch order (kind 0 = mark, 1 = k_step, 2 = k_encode_obs)."""
        out = np.zeros(1 << 16, self.TRACE_DTYPE)
        n = C.c_int64(0)
        check(self._L.pb_trace_end(self._h, out.ctypes.data_as(C.c_void_p), len(out), C.byref(n)))
        return out[:min(n.value, len(out))].copy()

    def clone_env_from(self, dst_index: int, src: "BatchedPacmanGym", src_index: int) -> None:
        """`PacmanGym.clone()` (env.rs:96 derive(Clone)): copy one env between engines."""
        check(self._L.pb_clone_env(self._h, dst_index, src._h, src_index, self._stream()))
        self._cfg[dst_index] = src._cfg[src_index]

    def copy_envs_from(self, src: "BatchedPacmanGym", src_index: Optional[torch.Tensor] = None,
                       dst_index: Optional[torch.Tensor] = None, count: Optional[int] = None) -> None:
        """Batched clone on the device (pb_copy_envs): env src_index[k] of `src` overwrites env
        dst_index[k] of this engine; int64 device tensors, None = identity, negative = skip."""
        def ptr(t, name):
            if t is None:
                return None
            if t.dtype != torch.int64 or t.device != self.device or not t.is_contiguous() or t.ndim != 1:
                raise ValueError(f"{name}: expected a contiguous 1-D int64 tensor on {self.device}")
            return C.c_void_p(t.data_ptr())

        if count is None:
            if dst_index is not None:
                count = dst_index.shape[0]
            elif src_index is not None:
                count = src_index.shape[0]
            else:
                count = min(self.num_envs, src.num_envs)
        check(self._L.pb_copy_envs(self._h, ptr(dst_index, "dst_index"), src._h,
                                   ptr(src_index, "src_index"), int(count), self._stream()))

    def view(self, index: int) -> "PacmanGymView":
        from .gym import PacmanGymView

        return PacmanGymView(self, index)
