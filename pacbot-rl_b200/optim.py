"""The optimizer tail of a DQN iteration on the repo's kernels.

Reference: /root/reference/src/algorithms/train_dqn.py:133 (`torch.optim.Adam(q_net.parameters(),
lr)`) and :200-206 (`clip_grad_norm_(..., max_norm)` then `optimizer.step()`).  `ClipAdam` does
both in two launches (`pb_clip_adam`, csrc/pb_dqn_abi.inl) instead of torch's ~8; same
hyper-parameters and defaults as `torch.optim.Adam` (betas (0.9, 0.999), eps 1e-8, no weight decay,
no amsgrad), results within float32 rounding of the torch pair (tests/test_dqn_gpu.py)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _lib


def _dense(t: torch.Tensor) -> bool:
    """non-overlapping and dense: some permutation of the dims is contiguous"""
    dims = sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1)
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


class ClipAdam:
    """`grad_norm = clip_grad_norm_(params, max_norm); Adam.step()` as `clip_and_step(max_norm)`.

    The moments are two flat float32 buffers in parameter order; the step counter lives on the
    device, so the call can be captured in a CUDA graph and replayed.  `p.grad` is read, not
    rewritten with the clipped values."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        self.params = [p for p in params]
        if not 0 < len(self.params) <= 32:
            raise ValueError("ClipAdam: 1..32 parameter tensors")
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32 or dev.type != "cuda":
                raise ValueError("ClipAdam: float32 CUDA parameters on one device")
            if not _dense(p):
                raise ValueError("ClipAdam: parameters must be non-overlapping and dense")
        self.device = dev
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps}]
        total = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._partials = torch.zeros(8 * len(self.params) + 2, dtype=torch.float32, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        k = len(self.params)
        self._sizes = (C.c_int64 * k)(*[p.numel() for p in self.params])
        self._pp = (C.c_void_p * k)(*[p.data_ptr() for p in self.params])
        self._gp = (C.c_void_p * k)()

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def clip_and_step(self, max_norm: float) -> torch.Tensor:
        """Returns the total gradient norm before clipping (0-dim device tensor, overwritten by
        the next call), like `clip_grad_norm_`."""
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                raise RuntimeError("ClipAdam: a parameter has no gradient")
            if g.stride() != p.stride() or g.dtype != torch.float32:
                # storage order is what the kernel pairs: re-lay the gradient out like its parameter
                # (autograd's layout contract makes this the rare case)
                g = p.grad = torch.empty_like(p).copy_(g)
            self._pp[i] = p.data_ptr()
            self._gp[i] = g.data_ptr()
        lr = float(self.param_groups[0]["lr"])
        _lib.check(_lib.load().pb_clip_adam(
            len(self.params), self._pp, self._gp, self._sizes, C.c_void_p(self.exp_avg.data_ptr()),
            C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.step_count.data_ptr()),
            C.c_void_p(self._partials.data_ptr()), float(max_norm), lr, self.betas[0], self.betas[1],
            self.eps, C.c_void_p(self._norm.data_ptr()), self.device.index, _lib.raw_stream(self.device)))
        return self._norm[0]

    def step(self) -> None:
        """`optimizer.step()` without clipping."""
        self.clip_and_step(float("inf"))

    def state_dict(self) -> dict:
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "lr": self.param_groups[0]["lr"],
                "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, state: dict) -> None:
        self.step_count.copy_(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.param_groups[0]["lr"] = float(state["lr"])
