"""Q-network of the DQN loop.  Stays PyTorch/cuDNN (north star): tensor cores are used only here.

`QNetV2` reproduces the architecture, initialisation and state-dict key names of the reference
(/root/reference/src/models.py:66-110; keys `backbone.{0,2,5,8,11,15}`, `value_head.0`,
`advantage_head.0` as consumed by pacbot_rs/src/candle_qnet.rs:18-44) so checkpoints written by
either side load in the other (`q_net-latest.pt` / `.safetensors`, train_dqn.py:239-253)."""
from __future__ import annotations

import copy

import torch
from torch import nn


def init_orthogonal(module: nn.Module) -> None:
    """Orthogonal init of every weight with >= 2 dims (reference models.py:6-14)."""
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() >= 2:
                nn.init.orthogonal_(p)


def portable_copy(net: nn.Module) -> nn.Module:
    """A detached copy of `net` with plain row-major parameters, for checkpoints.  The trainers
    keep conv weights in NHWC (`use_channels_last`); a pickled module written that way breaks the
    reference's export path (`to_safetensors.py`: `save_file(q_net.state_dict())` refuses
    non-contiguous tensors) and any other consumer that expects the layout the reference writes."""
    out = copy.deepcopy(net).to(memory_format=torch.contiguous_format)
    if getattr(out, "_channels_last", False):
        out._channels_last = False
    if getattr(out, "_fused_blocks", False):
        out._fused_blocks = False
    for m in out.modules():          # scratch buffers of the fused forward do not belong in a checkpoint
        m.__dict__.pop("_pb_pad_buffers", None)
    return out


class _BiasPoolSiLU(torch.autograd.Function):
    """`y = SiLU(maxpool2x2_ceil(x + bias))` (pool=True) or `y = SiLU(x + bias)` (pool=False) for a
    channels-last float32 CUDA activation, forward and backward each ONE launch of the repo's
    kernels (`pb_bias_pool_silu_fwd / _bwd`, csrc/pb_qnet_abi.inl) instead of torch's add_ +
    max_pool2d + silu and their three backward kernels.  Same arithmetic as that composition
    (reference models.py:70-90), see the kernel header."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, bias: torch.Tensor, pool: bool) -> torch.Tensor:
        import ctypes as C

        from . import _lib

        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise ValueError("bias_pool_silu: float32 CUDA [N, C, H, W] activation expected")
        x = x.contiguous(memory_format=torch.channels_last)
        n, c, h, w = x.shape
        ho, wo = ((h + 1) // 2, (w + 1) // 2) if pool else (h, w)
        dev = x.device
        y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        need_bwd = any(ctx.needs_input_grad[:2])
        z = torch.empty_like(y) if need_bwd else None
        idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=dev) if (need_bwd and pool) else None
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(_lib.load().pb_bias_pool_silu_fwd(
            p(x), p(bias.contiguous()), n, h, w, c, int(pool), p(z), p(y), p(idx), dev.index,
            _lib.raw_stream(dev)))
        if need_bwd:
            ctx.save_for_backward(z, idx) if pool else ctx.save_for_backward(z)
            ctx.pool, ctx.in_shape = pool, (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        import ctypes as C

        from . import _lib

        saved = ctx.saved_tensors
        z, idx = saved[0], (saved[1] if ctx.pool else None)
        n, c, h, w = ctx.in_shape
        dev = dy.device
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        db = torch.empty(c, dtype=torch.float32, device=dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(_lib.load().pb_bias_pool_silu_bwd(
            p(dy), p(z), p(idx), n, h, w, c, int(ctx.pool), p(dx), p(db), dev.index,
            _lib.raw_stream(dev)))
        return dx, db, None


def bias_pool_silu(x: torch.Tensor, bias: torch.Tensor, pool: bool) -> torch.Tensor:
    return _BiasPoolSiLU.apply(x, bias, pool)


def qnet_head(x: torch.Tensor, conv_bias: torch.Tensor, lin: nn.Linear, out: nn.Linear,
              value: "nn.Linear | None" = None, greedy_mask: "torch.Tensor | None" = None):
    """Inference-only head of QNetV2 / NetV2 as ONE launch (`pb_qnet_head_fwd`,
    csrc/pb_qnet_abi.inl): global max of the last convolution's bias-free NHWC output `x`
    [N, 128, H, W] + `conv_bias`, SiLU, `lin` (128 -> 256), SiLU, then `value` / `out` combined as
    `V - mean(A) + A` (reference models.py:92-110) or, without `value`, the plain layer `out`
    (models.py:139-141).  No autograd graph: callers use it under `torch.no_grad()`.
    With `greedy_mask` (bool / uint8 [N, actions]) the same launch also takes the masked arg-max
    (`MaxQPolicy`, policies.py:44-59) and the result is `(q, actions uint8 [N])`."""
    import ctypes as C

    from . import _lib

    n, c, h, w = x.shape
    if not (x.is_cuda and x.dtype == torch.float32 and c == 128 and lin.in_features == 128
            and lin.out_features == 256 and x.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("qnet_head: channels-last float32 CUDA [N, 128, H, W] input and Linear(128, 256) expected")
    q = torch.empty((n, out.out_features), dtype=torch.float32, device=x.device)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
    head = (p(x), n, h * w, p(conv_bias), p(lin.weight), p(lin.bias),
            p(value.weight if value is not None else None), p(value.bias if value is not None else None),
            p(out.weight), p(out.bias), out.out_features, int(value is not None), p(q))
    tail = (x.device.index, _lib.raw_stream(x.device))
    if greedy_mask is None:
        _lib.check(_lib.load().pb_qnet_head_fwd(*head, *tail))
        return q
    m = greedy_mask.contiguous().view(torch.uint8)
    if tuple(m.shape) != (n, out.out_features) or m.device != x.device:
        raise ValueError("qnet_head: greedy_mask must be [N, actions] on the input's device")
    actions = torch.empty(n, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.load().pb_qnet_head_greedy(*head, p(m), p(actions), *tail))
    return q, actions


class _QNetHead(torch.autograd.Function):
    """`qnet_head` with gradients: forward = `pb_qnet_head_fwd_train` (one launch; keeps the two
    SiLU inputs and the arg-max positions), backward = `pb_qnet_head_bwd` (two launches: the
    per-sample chain incl. the scatter to the arg-max cells, and the batch reductions for the
    weight / bias gradients) instead of ~25 torch launches (max-pool, SiLU, three `addmm` and their
    backward GEMMs, reductions, the convolution's bias add and bias-gradient sum)."""

    @staticmethod
    def forward(ctx, x, conv_bias, w1, b1, wv, bv, wa, ba):
        import ctypes as C

        from . import _lib

        n, c, h, w = x.shape
        dev = x.device
        dueling = wv is not None
        n_out = wa.shape[0]
        q = torch.empty((n, n_out), dtype=torch.float32, device=dev)
        pre = torch.empty((n, 128), dtype=torch.float32, device=dev)
        idx = torch.empty((n, 128), dtype=torch.uint8, device=dev)
        u = torch.empty((n, 256), dtype=torch.float32, device=dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(_lib.load().pb_qnet_head_fwd_train(
            p(x), n, h * w, p(conv_bias), p(w1), p(b1), p(wv), p(bv), p(wa), p(ba), n_out, int(dueling),
            p(q), p(pre), p(idx), p(u), dev.index, _lib.raw_stream(dev)))
        ctx.save_for_backward(pre, idx, u, w1, wv, wa)
        ctx.shape = (n, c, h, w)
        return q

    @staticmethod
    def backward(ctx, dq):
        import ctypes as C

        from . import _lib

        pre, idx, u, w1, wv, wa = ctx.saved_tensors
        n, c, h, w = ctx.shape
        dev = dq.device
        L = _lib.load()
        dueling = wv is not None
        n_out = wa.shape[0]
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)   # noqa: E731
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        dcb, dw1, db1, dwa, dba = new(128), new(256, 128), new(256), new(n_out, 256), new(n_out)
        dwv, dbv = (new(1, 256), new(1)) if dueling else (None, None)
        scratch = new(int(L.pb_qnet_head_bwd_scratch_floats(n)))
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(L.pb_qnet_head_bwd(
            p(dq.contiguous()), p(pre), p(idx), p(u), p(w1), p(wv), p(wa), n, h * w, n_out, int(dueling),
            p(dx), p(dcb), p(dw1), p(db1), p(dwv), p(dbv), p(dwa), p(dba), p(scratch), dev.index,
            _lib.raw_stream(dev)))
        return dx, dcb, dw1, db1, dwv, dbv, dwa, dba


def _head_args_ok(x: torch.Tensor, lin: nn.Linear) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 128
            and lin.in_features == 128 and lin.out_features == 256 and x.shape[2] * x.shape[3] <= 255
            and x.is_contiguous(memory_format=torch.channels_last))


def qnet_head_train(x: torch.Tensor, conv_bias: torch.Tensor, lin: nn.Linear, out: nn.Linear,
                    value: "nn.Linear | None" = None) -> torch.Tensor:
    """`qnet_head` for the forward that needs gradients (`_QNetHead`): same arguments."""
    if not _head_args_ok(x, lin):
        raise ValueError("qnet_head_train: channels-last float32 CUDA [N, 128, H, W] input (H*W <= 255) and "
                         "Linear(128, 256) expected")
    return _QNetHead.apply(x.contiguous(memory_format=torch.channels_last), conv_bias, lin.weight, lin.bias,
                           None if value is None else value.weight, None if value is None else value.bias,
                           out.weight, out.bias)


class GlobalMaxPool2d(nn.Module):
    """`nn.AdaptiveMaxPool2d((1, 1))` (reference models.py:96) computed as ONE max-pool window over
    the whole feature map: same values, same arg-max choice (first maximum in row-major order) and
    therefore the same gradient routing, but torch's NHWC max-pool kernel instead of the adaptive
    one (33 us -> a few us for the [512, 128, 4, 4] map).  No parameters: the module sits at the
    same `backbone` index, state-dict keys are unchanged."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return nn.functional.max_pool2d(x, kernel_size=tuple(x.shape[-2:]))


_ZERO_I64: dict = {}


def obs_to_padded_nhwc(obs: torch.Tensor, channels: int = 32) -> torch.Tensor:
    """A channel-major observation batch [N, 17, 28, 31] (what the encoder writes) as a
    channels-last batch zero-padded to `channels` channels, in ONE launch (`pb_replay_slot_obs`, the
    transposing copy of the replay ring): torch would run a layout-conversion kernel and cuDNN its
    own channel-padding kernel before a first convolution that, with 17 input channels, has no
    tensor-core kernel.  The networks pad their first weight to match (see `_fused_trunk`)."""
    import ctypes as C

    from . import _lib

    n, dev = obs.shape[0], obs.device
    out = torch.empty((n, channels) + tuple(obs.shape[2:]), dtype=torch.float32, device=dev,
                      memory_format=torch.channels_last)
    zero = _ZERO_I64.get(dev)
    if zero is None:
        zero = _ZERO_I64[dev] = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(_lib.load().pb_replay_slot_obs(
        C.c_void_p(obs.data_ptr()), n, 1, C.c_void_p(zero.data_ptr()), C.c_void_p(out.data_ptr()), channels,
        dev.index, _lib.raw_stream(dev)))
    return out


def _is_raw_obs(obs: torch.Tensor) -> bool:
    return (obs.is_cuda and obs.dtype == torch.float32 and obs.dim() == 4 and tuple(obs.shape[1:]) == (17, 28, 31)
            and obs.is_contiguous() and not obs.requires_grad and obs.shape[0] > 0)


def _pad_buffer(conv: nn.Conv2d, channels: int) -> torch.Tensor:
    """Persistent channels-last weight buffer [out, channels, kh, kw] of `conv`, zero where the
    input channels are padding (allocated once per module and channel count; not a parameter, not
    in the state dict).  The no-grad forwards refresh its live part with ONE strided copy
    (`F.pad` + `.contiguous(channels_last)` are a fill and two copies, six times per DQN
    iteration)."""
    bufs = conv.__dict__.setdefault("_pb_pad_buffers", {})
    w = conv.weight
    key = (channels, w.device)
    buf = bufs.get(key)
    if buf is None:
        buf = bufs[key] = torch.zeros((w.shape[0], channels) + tuple(w.shape[2:]), dtype=w.dtype,
                                      device=w.device).contiguous(memory_format=torch.channels_last)
    return buf


def _fused_trunk(bb: nn.Sequential, obs: torch.Tensor) -> torch.Tensor:
    """Blocks 0..8 of the QNetV2 / NetV2 layer list (conv5x5 + SiLU, 3 x [conv3x3, max pool, SiLU])
    with every block tail as one `bias_pool_silu` launch; the convolutions run without bias."""
    first = bb[0]
    w = first.weight
    if obs.shape[1] > first.in_channels:   # zero-padded observation channels, see QNetV2.forward()
        if torch.is_grad_enabled():        # a fresh tensor: autograd may keep it for a later backward
            w = nn.functional.pad(w, (0, 0, 0, 0, 0, obs.shape[1] - first.in_channels)).contiguous(
                memory_format=torch.channels_last)
        else:                              # one strided copy into the module's persistent buffer
            w = _pad_buffer(first, obs.shape[1])
            w[:, :first.in_channels].copy_(first.weight)
    else:
        w = w.contiguous(memory_format=torch.channels_last)
    h = bias_pool_silu(nn.functional.conv2d(obs, w, None, padding="same"), first.bias, False)
    for i in (2, 5, 8):                     # Conv2d, MaxPool2d(2, ceil_mode=True), SiLU
        conv = bb[i]
        h = bias_pool_silu(nn.functional.conv2d(h, conv.weight, None, padding="same"), conv.bias, True)
    return h


class QNetV2(nn.Module):
    """conv5x5(C->16) SiLU, 3 x [conv3x3, maxpool2(ceil), SiLU] (32, 64, 128), grouped conv3x3
    (128, 8 groups), global max, SiLU, Linear 128->256, SiLU, dueling heads V(1) / A(actions);
    Q = V - mean(A) + A.  156 934 parameters for the (17, 28, 31) observation and 5 actions."""

    def __init__(self, obs_shape, action_count: int) -> None:
        super().__init__()
        in_ch = int(obs_shape[0])
        layers: list[nn.Module] = [nn.Conv2d(in_ch, 16, 5, padding="same"), nn.SiLU()]
        for cin, cout in ((16, 32), (32, 64), (64, 128)):
            layers += [nn.Conv2d(cin, cout, 3, padding="same"), nn.MaxPool2d(2, ceil_mode=True), nn.SiLU()]
        layers += [
            nn.Conv2d(128, 128, 3, groups=8, padding="same"),
            GlobalMaxPool2d(),
            nn.Flatten(),
            nn.SiLU(),
            nn.Linear(128, 256),
            nn.SiLU(),
        ]
        self.backbone = nn.Sequential(*layers)
        self.value_head = nn.Sequential(nn.Linear(256, 1))
        self.advantage_head = nn.Sequential(nn.Linear(256, action_count))
        init_orthogonal(self)
        with torch.no_grad():  # both heads start at zero => Q == 0 at init
            for head in (self.value_head[-1], self.advantage_head[-1]):
                head.weight.zero_()
                head.bias.zero_()

    def use_channels_last(self) -> "QNetV2":
        """Keep weights (and run activations) in NHWC: cuDNN's tensor-core conv kernels want it
        and otherwise transpose on every call (fwd+bwd at batch 512: 2.51 -> 1.96 ms on B200,
        profiles/r01_dqn_torch_profiler.txt).  Parameter names and shapes are unchanged."""
        self.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def use_fused_blocks(self, enabled: bool = True) -> "QNetV2":
        """Run the tail of every convolution block (bias, 2x2 max pool, SiLU) as the repo's fused
        kernels (`bias_pool_silu`), forward and backward; the convolutions stay cuDNN.  Needs the
        channels-last mode and CUDA float32 inputs; parameters and checkpoints are unchanged."""
        if enabled and not getattr(self, "_channels_last", False):
            self.use_channels_last()
        self._fused_blocks = bool(enabled)
        return self

    def _forward_fused(self, obs: torch.Tensor) -> torch.Tensor:
        bb = self.backbone
        h = _fused_trunk(bb, obs)
        if not torch.is_grad_enabled():         # policy / target / evaluation forwards: fused head
            last = bb[11]
            x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
            return qnet_head(x5, last.bias, bb[15], self.advantage_head[0], self.value_head[0])
        last = bb[11]                            # the forward that needs gradients: fused head + backward
        x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
        return qnet_head_train(x5, last.bias, bb[15], self.advantage_head[0], self.value_head[0])

    @torch.no_grad()
    def greedy_actions(self, obs: torch.Tensor, action_masks: torch.Tensor) -> torch.Tensor:
        """`MaxQPolicy(self)(obs, action_masks)` as uint8 [N]: with the fused blocks on, the masked
        arg-max comes out of the head's own launch (no masked_fill / argmax / cast kernels)."""
        if getattr(self, "_fused_blocks", False) and obs.is_cuda and obs.dtype == torch.float32:
            if _is_raw_obs(obs):
                obs = obs_to_padded_nhwc(obs)
            obs = obs.contiguous(memory_format=torch.channels_last)
            bb = self.backbone
            h = _fused_trunk(bb, obs)
            last = bb[11]
            x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
            return qnet_head(x5, last.bias, bb[15], self.advantage_head[0], self.value_head[0],
                             greedy_mask=action_masks)[1]
        return self(obs).masked_fill(~action_masks, -torch.inf).argmax(dim=1).to(torch.uint8)

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        if getattr(self, "_fused_blocks", False) and _is_raw_obs(obs):
            obs = obs_to_padded_nhwc(obs)          # encoder output -> NHWC, 32 channels, one launch
        if getattr(self, "_channels_last", False):
            obs = obs.contiguous(memory_format=torch.channels_last)
        if getattr(self, "_fused_blocks", False) and obs.is_cuda and obs.dtype == torch.float32:
            return self._forward_fused(obs)
        first = self.backbone[0]
        if obs.shape[1] > first.in_channels:
            # zero-padded observation channels (ReplayBuffer(pad_channels=True)): the first weight
            # gets matching zero input channels, so the result is that of the 17-channel
            # convolution while cuDNN sees a channel count its tensor-core kernels accept
            w = nn.functional.pad(first.weight, (0, 0, 0, 0, 0, obs.shape[1] - first.in_channels))
            if getattr(self, "_channels_last", False):
                w = w.contiguous(memory_format=torch.channels_last)
            feat = self.backbone[1:](nn.functional.conv2d(obs, w, first.bias, padding="same"))
        else:
            feat = self.backbone(obs)
        value = self.value_head(feat)
        adv = self.advantage_head(feat)
        return value - adv.mean(dim=1, keepdim=True) + adv


class QNet(nn.Module):
    """The reference's first Q-network (models.py:17-63), selectable with `--model QNet`: two
    conv+SiLU blocks (5x5 C->16, 3x3 16->32) and a 3x3 conv to 64 channels at full resolution,
    global max over the board, SiLU, Linear 64->128, SiLU, dueling heads; same module nesting
    (`backbone.{0.0,1.0,2}`, `fc.1`, `value_head.0`, `advantage_head.0`) for checkpoint exchange."""

    def __init__(self, obs_shape, action_count: int) -> None:
        super().__init__()
        in_ch = int(obs_shape[0])
        self.backbone = nn.Sequential(
            nn.Sequential(nn.Conv2d(in_ch, 16, 5, padding="same"), nn.SiLU()),
            nn.Sequential(nn.Conv2d(16, 32, 3, padding="same"), nn.SiLU()),
            nn.Conv2d(32, 64, 3, padding="same"),
        )
        self.fc = nn.Sequential(nn.SiLU(), nn.Linear(64, 128), nn.SiLU())
        self.value_head = nn.Sequential(nn.Linear(128, 1))
        self.advantage_head = nn.Sequential(nn.Linear(128, action_count))
        init_orthogonal(self)
        with torch.no_grad():
            for head in (self.value_head[-1], self.advantage_head[-1]):
                head.weight.zero_()
                head.bias.zero_()

    def use_channels_last(self) -> "QNet":
        self.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        if getattr(self, "_channels_last", False):
            obs = obs.contiguous(memory_format=torch.channels_last)
        feat = self.fc(self.backbone(obs).amax(dim=(-2, -1)))
        adv = self.advantage_head(feat)
        return self.value_head(feat) - adv.mean(dim=1, keepdim=True) + adv


class DebugMLPQNet(nn.Module):
    """The reference's debugging MLP (models.py:150-167), `--model DebugMLPQNet`: flatten,
    Linear -> 128, SiLU, Linear -> actions with a zeroed last layer (default init otherwise)."""

    def __init__(self, obs_shape, action_count: int) -> None:
        super().__init__()
        n_in = 1
        for d in obs_shape:
            n_in *= int(d)
        self.mlp = nn.Sequential(nn.Linear(n_in, 128), nn.SiLU(), nn.Linear(128, action_count))
        with torch.no_grad():
            self.mlp[-1].weight.zero_()
            self.mlp[-1].bias.zero_()

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        # logical (C, H, W) order whatever the memory format of the batch
        return self.mlp(obs.contiguous().view(obs.shape[0], -1))


class NetV2(nn.Module):
    """The AlphaZero network (reference models.py:113-147): QNetV2's trunk with one linear head of
    `output_dim` outputs (5 policy logits + 1 value, train_alphazero.py:75), last layer zero
    initialised.  Same module indices as the reference (`network.{0,2,5,8,11,15,17}`) so that
    `torch.save(model)` / state dicts interchange."""

    def __init__(self, obs_shape, output_dim: int) -> None:
        super().__init__()
        in_ch = int(obs_shape[0])
        layers: list[nn.Module] = [nn.Conv2d(in_ch, 16, 5, padding="same"), nn.SiLU()]
        for cin, cout in ((16, 32), (32, 64), (64, 128)):
            layers += [nn.Conv2d(cin, cout, 3, padding="same"), nn.MaxPool2d(2, ceil_mode=True), nn.SiLU()]
        layers += [
            nn.Conv2d(128, 128, 3, groups=128 // 16, padding="same"),
            GlobalMaxPool2d(),
            nn.Flatten(),
            nn.SiLU(),
            nn.Linear(128, 256),
            nn.SiLU(),
            nn.Linear(256, output_dim),
        ]
        self.network = nn.Sequential(*layers)
        init_orthogonal(self)
        with torch.no_grad():
            self.network[-1].weight.zero_()
            self.network[-1].bias.zero_()

    def use_channels_last(self) -> "NetV2":
        self.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def use_fused_blocks(self, enabled: bool = True) -> "NetV2":
        """As QNetV2.use_fused_blocks: block tails as `bias_pool_silu`, and under `torch.no_grad()`
        (the MCTS evaluator) the whole head as one `qnet_head` launch."""
        if enabled and not getattr(self, "_channels_last", False):
            self.use_channels_last()
        self._fused_blocks = bool(enabled)
        return self

    @torch.no_grad()
    def policy_value(self, obs: torch.Tensor, action_masks: torch.Tensor, value_scale: float = 1.0):
        """The evaluator of the AlphaZero loop (reference train_alphazero.py:83-95): `(value [N],
        policy [N, actions])` with `value = out[:, -1] * value_scale` and `policy =
        softmax(out[:, :-1] masked with -inf)`.  With the fused blocks on, both come out of the
        head's own launch (`pb_qnet_head_policy_value`)."""
        if getattr(self, "_fused_blocks", False) and obs.is_cuda and obs.dtype == torch.float32:
            import ctypes as C

            from . import _lib

            if _is_raw_obs(obs):
                obs = obs_to_padded_nhwc(obs)
            obs = obs.contiguous(memory_format=torch.channels_last)
            net = self.network
            h = _fused_trunk(net, obs)
            last, lin, out = net[11], net[15], net[17]
            x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
            n, dev, rows = x5.shape[0], x5.device, out.out_features
            m = action_masks.contiguous().view(torch.uint8)
            if tuple(m.shape) != (n, rows - 1):
                raise ValueError("policy_value: action_masks must be [N, actions]")
            raw = torch.empty((n, rows), dtype=torch.float32, device=dev)
            value = torch.empty(n, dtype=torch.float32, device=dev)
            policy = torch.empty((n, rows - 1), dtype=torch.float32, device=dev)
            p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
            _lib.check(_lib.load().pb_qnet_head_policy_value(
                p(x5), n, x5.shape[2] * x5.shape[3], p(last.bias), p(lin.weight), p(lin.bias), p(out.weight),
                p(out.bias), rows, p(m), float(value_scale), p(raw), p(value), p(policy), dev.index,
                _lib.raw_stream(dev)))
            return value, policy
        pred = self(obs)
        return (pred[:, -1] * value_scale,
                torch.softmax(pred[:, :-1].masked_fill(~action_masks.bool(), -torch.inf), dim=-1))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        if getattr(self, "_fused_blocks", False) and _is_raw_obs(obs):
            obs = obs_to_padded_nhwc(obs)          # encoder output -> NHWC, 32 channels, one launch
        if getattr(self, "_channels_last", False):
            obs = obs.contiguous(memory_format=torch.channels_last)
        if getattr(self, "_fused_blocks", False) and obs.is_cuda and obs.dtype == torch.float32:
            net = self.network
            h = _fused_trunk(net, obs)
            if not torch.is_grad_enabled():
                last = net[11]
                x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
                return qnet_head(x5, last.bias, net[15], net[17])
            last = net[11]
            x5 = nn.functional.conv2d(h, last.weight, None, padding="same", groups=last.groups)
            return qnet_head_train(x5, last.bias, net[15], net[17])
        return self.network(obs)
