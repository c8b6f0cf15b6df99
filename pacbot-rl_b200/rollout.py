"""Software-pipelined rollouts: the step kernel runs while the observation encoder is still
streaming the previous observations to HBM.

A rollout step is `k_step` (10 us, latency bound: one thread per env) followed by `k_encode_obs`
(0.52 ms per 65 536 envs, HBM-write bound).  On one stream the 10 us are exposed every step.
Three drivers, all bit-identical to the one-stream order (tests/test_rollout_gpu.py):

  DoubleBufferedRandomRollout  random valid actions; the engine's second state buffer lets step t+1
                               overlap encode t (what bench.py measures by default)
  HostActionRollout            the same with host action / result buffers (bench.py `e2e`)
  PipelinedRandomRollout       in place: the batch is split into chunks that are skewed over the
                               two streams (no second buffer; pays one more encoder prologue/tail)

PipelinedRandomRollout in detail -- the envs are independent, so with two chunks A and B:

    encode stream : enc(t, A)   enc(t, B)     enc(t+1, A)   enc(t+1, B)   ...
    step stream   :      step(t+1, A)   step(t+1, B)   step(t+2, A)  ...

step(t+1, c) waits for enc(t, c) (it overwrites the state that launch reads) and enc(t+1, c) waits
for step(t+1, c); a step launch therefore always overlaps the encode of the OTHER chunk.  The
step kernel's CTAs (128 threads, 128 B of shared memory) co-reside with the encoder's persistent
CTAs (one per SM, 226 KB), and the HBM stream never pauses.  Results are bit-identical to the
sequential order: each env sees the same sequence of launches.

Per-env semantics are those of `BatchedPacmanGym.step_random` + `obs_range`."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .engine import BatchedPacmanGym


class PipelinedRandomRollout:
    def __init__(self, env: BatchedPacmanGym, num_chunks: int = 2,
                 chunks: Optional[Sequence[Tuple[int, int]]] = None) -> None:
        self.env = env
        n = env.num_envs
        if chunks is None:
            num_chunks = max(1, min(int(num_chunks), n))
            per = (n + num_chunks - 1) // num_chunks
            chunks = [(f, min(per, n - f)) for f in range(0, n, per)]
        self.chunks: List[Tuple[int, int]] = [(int(f), int(c)) for f, c in chunks]
        dev = env.device
        self.enc_stream = torch.cuda.Stream(dev)
        self.step_stream = torch.cuda.Stream(dev)
        self._enc_done = [torch.cuda.Event() for _ in self.chunks]
        self._step_done = [torch.cuda.Event() for _ in self.chunks]
        self._started = False
        self.reward = torch.empty(n, dtype=torch.int32, device=dev)
        self.done = torch.empty(n, dtype=torch.bool, device=dev)
        self.mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
        self.actions = torch.empty(n, dtype=torch.uint8, device=dev)

    def _fork(self) -> None:
        cur = torch.cuda.current_stream(self.env.device)
        self.enc_stream.wait_stream(cur)
        self.step_stream.wait_stream(cur)
        self._started = True

    def step(self, call_idx: int, obs_out: Callable[[int, int, int], torch.Tensor],
             on_encode: Optional[Callable[[int, torch.cuda.Stream, bool], None]] = None) -> None:
        """Enqueue one rollout step of every env.  obs_out(chunk_index, first, count) returns the
        float32 [count,17,28,31] tensor that chunk's observations go to.  Outputs of the step
        kernel land in self.reward / done / mask / actions (valid after join(); a later step
        overwrites them).  on_encode(chunk_index, stream, before) lets a caller time the encodes."""
        if not self._started:
            self._fork()
        env = self.env
        for c, (first, count) in enumerate(self.chunks):
            with torch.cuda.stream(self.step_stream):
                self.step_stream.wait_event(self._enc_done[c])   # enc(t-1, c) has read the old state
                env.step_random(call_idx=call_idx, actions_out=self.actions, mask_out=self.mask,
                                reward_out=self.reward, done_out=self.done, first=first, count=count)
                self._step_done[c].record(self.step_stream)
            with torch.cuda.stream(self.enc_stream):
                self.enc_stream.wait_event(self._step_done[c])
                if on_encode is not None:
                    on_encode(c, self.enc_stream, True)
                env.obs_range(first, count, obs_out(c, first, count))
                if on_encode is not None:
                    on_encode(c, self.enc_stream, False)
                self._enc_done[c].record(self.enc_stream)

    def join(self) -> None:
        """Make the caller's current stream wait for everything enqueued so far."""
        cur = torch.cuda.current_stream(self.env.device)
        cur.wait_stream(self.enc_stream)
        cur.wait_stream(self.step_stream)
        self._started = False


class DoubleBufferedRandomRollout:
    """The same overlap without splitting the encode launch: the engine keeps TWO state buffers
    (pb_step_random_flip).  step(t+1) reads buffer t and writes buffer t+1 while enc(t) is still
    streaming buffer t to HBM; it only has to wait for enc(t-1), whose buffer it overwrites.

        encode stream : enc(t)                enc(t+1)               ...
        step stream   :     step(t+1)              step(t+2)         ...

    One full-size encode launch per step (half-size launches lose to their own prologue / tail
    what the overlap wins), the step kernel never on the critical path."""

    def __init__(self, env: BatchedPacmanGym, chunks: Optional[Sequence[Tuple[int, int]]] = None) -> None:
        self.env = env
        n = env.num_envs
        self.chunks: List[Tuple[int, int]] = [(0, n)] if chunks is None else [(int(f), int(c)) for f, c in chunks]
        dev = env.device
        self.enc_stream = torch.cuda.Stream(dev)
        self.step_stream = torch.cuda.Stream(dev)
        self._enc_done = [torch.cuda.Event(), torch.cuda.Event()]  # of the last two steps
        self._step_done = torch.cuda.Event()
        self._t = 0
        self._started = False
        self.reward = torch.empty(n, dtype=torch.int32, device=dev)
        self.done = torch.empty(n, dtype=torch.bool, device=dev)
        self.mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
        self.actions = torch.empty(n, dtype=torch.uint8, device=dev)

    def step(self, call_idx: int, obs_out: Callable[[int, int, int], torch.Tensor],
             on_encode: Optional[Callable[[int, torch.cuda.Stream, bool], None]] = None) -> None:
        """Enqueue one rollout step of every env; obs_out(chunk_index, first, count) as in
        PipelinedRandomRollout.  self.reward / done / mask / actions are valid after join()."""
        env = self.env
        if not self._started:
            cur = torch.cuda.current_stream(env.device)
            self.enc_stream.wait_stream(cur)
            self.step_stream.wait_stream(cur)
            self._started = True
        with torch.cuda.stream(self.step_stream):
            # this launch overwrites the buffer that the encode of two steps ago read
            self.step_stream.wait_event(self._enc_done[self._t & 1])
            env.step_random(call_idx=call_idx, actions_out=self.actions, mask_out=self.mask,
                            reward_out=self.reward, done_out=self.done, flip=True)
            self._step_done.record(self.step_stream)
        with torch.cuda.stream(self.enc_stream):
            self.enc_stream.wait_event(self._step_done)
            for c, (first, count) in enumerate(self.chunks):
                if on_encode is not None:
                    on_encode(c, self.enc_stream, True)
                env.obs_range(first, count, obs_out(c, first, count))
                if on_encode is not None:
                    on_encode(c, self.enc_stream, False)
            self._enc_done[self._t & 1].record(self.enc_stream)
        self._t += 1

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.env.device)
        cur.wait_stream(self.enc_stream)
        cur.wait_stream(self.step_stream)
        self._started = False


class HostActionRollout:
    """End-to-end stepping with HOST action / result buffers, pipelined like
    DoubleBufferedRandomRollout: what `BatchedPacmanGym.step_host` does (pinned actions host->device,
    step, observation encode into a device buffer, reward / done / mask device->host, read by the
    host every step), except that the step of call t+1 no longer queues behind the encode of call
    t.  step() returns when reward / done / mask of THIS step are in the pinned host arrays; the
    observation is complete after join() (or once a later step() has returned)."""

    def __init__(self, env: BatchedPacmanGym) -> None:
        self.env = env
        n, dev = env.num_envs, env.device
        self.step_stream = torch.cuda.Stream(dev)
        self.enc_stream = torch.cuda.Stream(dev)
        self.copy_stream = torch.cuda.Stream(dev)
        self._enc_done = [torch.cuda.Event(), torch.cuda.Event()]
        self._step_done = torch.cuda.Event()
        self._t = 0
        self._started = False
        self._actions = torch.empty(n, dtype=torch.uint8, device=dev)
        self._reward = torch.empty(n, dtype=torch.int32, device=dev)
        self._done = torch.empty(n, dtype=torch.bool, device=dev)
        self._mask = torch.empty((n, 5), dtype=torch.bool, device=dev)
        self.reward = torch.empty(n, dtype=torch.int32).pin_memory()
        self.done = torch.empty(n, dtype=torch.bool).pin_memory()
        self.mask = torch.empty((n, 5), dtype=torch.bool).pin_memory()

    def step(self, actions_pinned: torch.Tensor, obs_out: torch.Tensor):
        """actions_pinned: uint8 [n] host tensor (pinned for an asynchronous copy); obs_out:
        float32 [n,17,28,31] device tensor.  Returns (reward, done, mask) pinned host tensors,
        overwritten by the next call."""
        env = self.env
        if not self._started:
            cur = torch.cuda.current_stream(env.device)
            for s in (self.step_stream, self.enc_stream, self.copy_stream):
                s.wait_stream(cur)
            self._started = True
        with torch.cuda.stream(self.step_stream):
            self._actions.copy_(actions_pinned, non_blocking=True)
            self.step_stream.wait_event(self._enc_done[self._t & 1])
            env.step(self._actions, check_actions=False, mask_out=self._mask, reward_out=self._reward,
                     done_out=self._done, flip=True)
            self._step_done.record(self.step_stream)
        with torch.cuda.stream(self.enc_stream):
            self.enc_stream.wait_event(self._step_done)
            env.obs(out=obs_out)
            self._enc_done[self._t & 1].record(self.enc_stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._step_done)
            self.reward.copy_(self._reward, non_blocking=True)
            self.done.copy_(self._done, non_blocking=True)
            self.mask.copy_(self._mask, non_blocking=True)
        self.copy_stream.synchronize()   # the host reads the results of every step
        self._t += 1
        return self.reward, self.done, self.mask

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.env.device)
        for s in (self.step_stream, self.enc_stream, self.copy_stream):
            cur.wait_stream(s)
        self._started = False
