"""The Fourier layer around the spectral convolution (tanh stabilizer of f3): B200-native drop-ins
for `neuralop.layers.fno_block.FNOBlocks`, `neuralop.layers.channel_mlp.ChannelMLP` and the skip connections of
`neuralop.layers.skip_connections` -- same constructor arguments, same parameter names (state dicts load both ways), same forward
semantics (fno_block.py:371-453), with everything after the spectral convolution running in the fused kernels of
`csrc/sc_layer.cu`:

    f1   x1  = act( SpectralConv(x) + W_skip x )                 ONE launch: reads x and the conv output, writes x1
    f2   h   = gelu( W1 x1 + b1 );  out = act( W2 h + b2 + gate * x )     two launches

instead of one tensor pass per torch op (conv1d, add, gelu, mul, add, gelu ...).  No CPU / PyTorch fallback: the modules raise on
CPU tensors, and configurations the kernels do not cover (complex data, dropout, activations other than GELU, conv_bias_kernel > 1)
raise `NotImplementedError` at construction (activations: gelu, relu, silu, tanh; `conv_bias_kernel > 1` keeps torch's Conv{n}d for
that skip).  The four `norm` options are composed from the same kernels (one statistics pass + one
fused affine / add / activation launch per normalisation).
"""
import math
from typing import Union

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd.function import once_differentiable

from . import _lib
from .spectral_conv import SpectralConv, _ptr, _stream_ptr

Number = Union[int, float]
ACT_IDENTITY, ACT_GELU = _lib.ACT_IDENTITY, _lib.ACT_GELU


def activation_code(non_linearity) -> int:
    """The kernels' activation code for a `non_linearity` callable of the reference API (fno_block.py:150, channel_mlp.py:49)."""
    table = {F.gelu: _lib.ACT_GELU, F.relu: _lib.ACT_RELU, torch.relu: _lib.ACT_RELU, F.silu: _lib.ACT_SILU, torch.tanh: _lib.ACT_TANH,
             F.tanh: _lib.ACT_TANH}
    for fn, code in table.items():
        if non_linearity is fn:
            return code
    raise NotImplementedError(f"non_linearity {non_linearity!r}: the fused kernels implement F.gelu (exact erf form), F.relu, F.silu "
                              "and torch.tanh")


def _require_device_tensor(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"neuraloperator_b200 has no CPU path: {what} must live on a B200 (got {t.device})")
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32 (full precision, real data), got {t.dtype}")


def _require_complex_input(x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError(f"neuraloperator_b200 has no CPU path: FNOBlocks input must live on a B200 (got {x.device})")
    if x.dtype != torch.complex64:
        raise TypeError(f"FNOBlocks(complex_data=True) expects complex64 input, got {x.dtype}")


def set_tensor_core_mixing(enable: bool) -> None:
    """Route `channel_mix` (forward and input gradient; Ci <= 256, Co <= 128) through the tcgen05 bf16x3 kernel instead of the exact
    fp32 SIMT kernel.  OPT-IN: that kernel was written without hardware access (also: SC_MIX_TC=1 in the environment)."""
    _lib.check(_lib.load().sc_layer_set_tensor_cores(int(bool(enable))), "sc_layer_set_tensor_cores")


def uses_tensor_core_mixing() -> bool:
    return bool(_lib.load().sc_layer_uses_tensor_cores())


def _launch_channel_mix(x, weight, w_stride_o, w_stride_i, bias, add, gate, gated, act, out, pre, B, Ci, Co, P):
    lib = _lib.load()
    dev = out.device
    with torch.cuda.device(dev):
        _lib.check(lib.sc_channel_mix(_ptr(x), _ptr(weight), w_stride_o, w_stride_i, _ptr(bias), _ptr(add), _ptr(gate), _ptr(gated), act,
                                      _ptr(out), _ptr(pre), B, Ci, Co, P, _stream_ptr(dev)), "sc_channel_mix")


class _ChannelMix(torch.autograd.Function):
    """out = act( W x + bias + add + gate * gated )  on (B, C, *S) tensors; any of x/W, bias, add, gate, gated may be None
    (gate None with gated given: coefficient 1).  W: (Co, Ci) or a Conv1d weight (Co, Ci, 1); gate: any shape with Co elements."""

    @staticmethod
    def forward(ctx, x, weight, bias, add, gate, gated, act):
        ref = add if add is not None else (gated if gated is not None else x)
        if x is None and ref is None:
            raise ValueError("channel_mix needs at least one tensor operand")
        B, spatial = ref.shape[0], tuple(ref.shape[2:])
        P = math.prod(spatial)
        Ci = x.shape[1] if x is not None else 0
        Co = weight.shape[0] if weight is not None else ref.shape[1]
        for name, t in (("x", x), ("add", add), ("gated", gated), ("weight", weight), ("bias", bias), ("gate", gate)):
            if t is not None:
                _require_device_tensor(t, f"channel_mix {name}")
        if x is not None:
            if weight is None or weight.numel() != Co * Ci:
                raise ValueError(f"channel_mix: weight must hold (Co, Ci) = ({Co}, {Ci}) elements")
            if tuple(x.shape[2:]) != spatial or x.shape[0] != B:
                raise ValueError(f"channel_mix: x {tuple(x.shape)} does not match the other operands (B={B}, grid={spatial})")
            x = x.contiguous()
            weight = weight.contiguous()
        for name, t in (("add", add), ("gated", gated)):
            if t is not None and tuple(t.shape) != (B, Co, *spatial):
                raise ValueError(f"channel_mix: {name} must be {(B, Co, *spatial)}, got {tuple(t.shape)}")
        if bias is not None and bias.numel() != Co:
            raise ValueError(f"channel_mix: bias must hold {Co} elements")
        if gate is not None and (gated is None or gate.numel() != Co):
            raise ValueError(f"channel_mix: gate must hold {Co} elements and needs the tensor it gates")
        add = add.contiguous() if add is not None else None
        gated = gated.contiguous() if gated is not None else None
        bias_c = bias.contiguous() if bias is not None else None
        gate_c = gate.contiguous() if gate is not None else None
        out = torch.empty((B, Co, *spatial), dtype=torch.float32, device=ref.device)
        # the GELU derivative needs the pre-activation: stored only when some input asks for a gradient
        pre = torch.empty_like(out) if act != ACT_IDENTITY and any(ctx.needs_input_grad) else None
        if out.numel():
            _launch_channel_mix(x, weight, Ci, 1, bias_c, add, gate_c, gated, act, out, pre, B, Ci, Co, P)
        ctx.act, ctx.dims = act, (B, Ci, Co, P)
        ctx.shapes = (weight.shape if weight is not None else None, bias.shape if bias is not None else None,
                      gate.shape if gate is not None else None)
        ctx.save_for_backward(x, weight, pre, gate_c, gated)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        x, weight, pre, gate_c, gated = ctx.saved_tensors
        act = ctx.act
        B, Ci, Co, P = ctx.dims
        w_shape, b_shape, g_shape = ctx.shapes
        need_x, need_w, need_b, need_add, need_gate, need_gated = ctx.needs_input_grad[:6]
        need_x = need_x and x is not None
        need_w = need_w and weight is not None
        dev = gout.device
        gout = gout.contiguous()
        if gout.dtype != torch.float32:
            gout = gout.float()
        if gout.numel() == 0:
            z = lambda t, need: torch.zeros_like(t) if need and t is not None else None      # noqa: E731
            return (z(x, need_x), z(weight, need_w), torch.zeros(b_shape, device=dev) if need_b and b_shape is not None else None,
                    gout if need_add else None, torch.zeros(g_shape, device=dev) if need_gate and g_shape is not None else None,
                    z(gated, need_gated), None)
        f32 = dict(dtype=torch.float32, device=dev)
        gpre = gout if act == ACT_IDENTITY else torch.empty_like(gout)
        dbias = torch.empty(Co, **f32) if need_b and b_shape is not None else None
        dgate = torch.empty(Co, **f32) if need_gate and g_shape is not None else None
        dgated = torch.empty_like(gout) if need_gated and gated is not None else None
        st = _stream_ptr(dev)
        with torch.cuda.device(dev):
            if act != ACT_IDENTITY or dbias is not None or dgate is not None or dgated is not None:
                _lib.check(lib.sc_channel_mix_act_backward(_ptr(gout), _ptr(pre), act, _ptr(gate_c), _ptr(gated if dgate is not None else None),
                                                           _ptr(gpre if act != ACT_IDENTITY else None), _ptr(dgated), _ptr(dbias), _ptr(dgate),
                                                           B, Co, P, st), "sc_channel_mix_act_backward")
            dx = None
            if need_x:
                # din[b, i, p] = sum_o W[o, i] gpre[b, o, p]: the mixing kernel with the weight read through transposed strides
                dx = torch.empty_like(x)
                _lib.check(lib.sc_channel_mix(_ptr(gpre), _ptr(weight), 1, Ci, None, None, None, None, ACT_IDENTITY, _ptr(dx), None,
                                              B, Co, Ci, P, st), "sc_channel_mix (input gradient)")
            dw = None
            if need_w:
                dw = torch.empty(Co, Ci, **f32)
                _lib.check(lib.sc_channel_mix_weight_grad(_ptr(gpre), _ptr(x), _ptr(dw), B, Ci, Co, P, st), "sc_channel_mix_weight_grad")
                dw = dw.view(w_shape)
        return (dx, dw, dbias.view(b_shape) if dbias is not None else None, gpre if need_add else None,
                dgate.view(g_shape) if dgate is not None else None, dgated, None)


def channel_mix(x=None, weight=None, bias=None, add=None, gate=None, gated=None, act: int = ACT_IDENTITY):
    """Functional form of the fused pointwise op (differentiable in every tensor argument)."""
    return _ChannelMix.apply(x, weight, bias, add, gate, gated, int(act))


class _Tanh(torch.autograd.Function):
    """The "tanh" stabilizer in front of the spectral conv (fno_block.py:386-390)."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        _require_device_tensor(x, "tanh stabilizer input")
        x = x.contiguous()
        out = torch.empty_like(x)
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.check(lib.sc_pointwise(_lib.POINTWISE_TANH, _ptr(x), None, _ptr(out), x.numel(), _stream_ptr(x.device)), "sc_pointwise")
        ctx.save_for_backward(out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        (out,) = ctx.saved_tensors
        g = g.contiguous().float()
        dx = torch.empty_like(out)
        if out.numel():
            with torch.cuda.device(g.device):
                _lib.check(lib.sc_pointwise(_lib.POINTWISE_TANH_BACKWARD, _ptr(g), _ptr(out), _ptr(dx), out.numel(), _stream_ptr(g.device)),
                           "sc_pointwise")
        return dx


class _RowSums(torch.autograd.Function):
    """(sum_p x, sum_p x^2) of every (b, c) row of a (B, C, *S) tensor in ONE pass over x: the reduction half of
    `k_channel_act_backward` on the tensor viewed as (1, B*C, P) -- with g = gated = x its dbias is sum g and its dgate sum g * gated.
    Backward: dx = g1[b,c] + 2 g2[b,c] x, the mixing kernel without a mixing term (bias + gate * gated) on the same view."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        _require_device_tensor(x, "normalisation input")
        x = x.contiguous()
        B, C = x.shape[:2]
        P = math.prod(x.shape[2:])
        s1 = torch.empty(B * C, dtype=torch.float32, device=x.device)
        s2 = torch.empty(B * C, dtype=torch.float32, device=x.device)
        if x.numel():
            with torch.cuda.device(x.device):
                _lib.check(lib.sc_channel_mix_act_backward(_ptr(x), None, ACT_IDENTITY, None, _ptr(x), None, None, _ptr(s1), _ptr(s2),
                                                           1, B * C, P, _stream_ptr(x.device)), "sc_channel_mix_act_backward (row sums)")
        else:
            s1.zero_()
            s2.zero_()
        ctx.save_for_backward(x)
        return s1.view(B, C), s2.view(B, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, g1, g2):
        (x,) = ctx.saved_tensors
        B, C = x.shape[:2]
        P = math.prod(x.shape[2:])
        dx = torch.empty_like(x)
        if x.numel():
            b = g1.reshape(-1).float().contiguous()
            g = (2.0 * g2).reshape(-1).float().contiguous()
            _launch_channel_mix(None, None, 0, 0, b, None, g, x, ACT_IDENTITY, dx, None, 1, 0, B * C, P)
        return dx


def _norm_scale_shift(norm: nn.Module, x: torch.Tensor):
    """Normalisation of x (B, C, *S) as a per-row affine map  y = scale[b,c] * x + shift[b,c]:  the statistics come from one pass over
    x (`_RowSums`); everything else is arithmetic on (B, C) tensors, differentiated by autograd (d weight / d bias of a GroupNorm
    fall out of it).  instance_norm: F.instance_norm without affine / running statistics (normalization_layers.py:60-96);
    group_norm: nn.GroupNorm(num_groups, C) (fno_block.py:318-326); batch_norm: nn.BatchNorm{n}d in training / eval mode incl. the
    running statistics (:99-158); ada_in: instance statistics with weight / bias from the embedding MLP (:5-57).  Biased variance, eps
    inside the square root, as all of them do."""
    B, C = x.shape[:2]
    P = math.prod(x.shape[2:])
    s1, s2 = _RowSums.apply(x)
    if isinstance(norm, nn.GroupNorm):
        G = norm.num_groups
        n = P * (C // G)
        mean = s1.view(B, G, -1).sum(-1) / n
        var = (s2.view(B, G, -1).sum(-1) / n - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + norm.eps)
        mean_c, rstd_c = mean.repeat_interleave(C // G, dim=1), rstd.repeat_interleave(C // G, dim=1)
        scale = rstd_c * norm.weight.view(1, C) if norm.weight is not None else rstd_c
        shift = -mean_c * scale
        if norm.bias is not None:
            shift = shift + norm.bias.view(1, C)
        return scale, shift
    if isinstance(norm, BatchNorm):
        bn = norm.norm                                               # F.batch_norm semantics (torch/nn/modules/batchnorm.py)
        use_batch = bn.training or not bn.track_running_stats
        if use_batch:
            n = B * P
            mean = s1.sum(0) / n
            var = (s2.sum(0) / n - mean * mean).clamp_min(0.0)       # biased: what normalises the batch
            if bn.training and bn.track_running_stats:
                with torch.no_grad():
                    bn.num_batches_tracked += 1
                    m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                    bn.running_mean.mul_(1 - m).add_(m * mean.detach())
                    bn.running_var.mul_(1 - m).add_(m * var.detach() * (n / max(n - 1, 1)))      # unbiased in the running average
        else:
            mean, var = bn.running_mean, bn.running_var
        scale = torch.rsqrt(var + bn.eps)
        if bn.weight is not None:
            scale = scale * bn.weight
        shift = -mean * scale
        if bn.bias is not None:
            shift = shift + bn.bias
        return scale.view(1, C).expand(B, C), shift.view(1, C).expand(B, C)
    mean = s1 / P
    var = (s2 / P - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + norm.eps)
    if isinstance(norm, AdaIN):
        if norm.embedding is None:
            raise RuntimeError("AdaIN: update embeddding before running forward")
        weight, bias = torch.split(norm.mlp(norm.embedding), C, dim=0)
        scale = rstd * weight.view(1, C)
        return scale, bias.view(1, C) - mean * scale
    return rstd, -mean * rstd


def _apply_norm(x, scale, shift, add=None, act=ACT_IDENTITY):
    """act( scale[b,c] * x + shift[b,c] + add ): ONE launch of the mixing kernel on the (1, B*C, *S) view (gate = scale, bias = shift)."""
    B, C = x.shape[:2]
    sp = tuple(x.shape[2:])
    out = channel_mix(bias=shift.reshape(-1), add=add.reshape(1, B * C, *sp) if add is not None else None, gate=scale.reshape(-1),
                      gated=x.reshape(1, B * C, *sp), act=act)
    return out.view(B, C, *sp)


class BatchNorm(nn.Module):
    """Parameter container for batch normalisation (normalization_layers.py:99-158): `self.norm` holds weight, bias and the running
    statistics under the reference's names (a BatchNorm1d: the parameters of BatchNorm{1,2,3}d are the same)."""

    def __init__(self, n_dim: int, num_features: int, **kwargs):
        super().__init__()
        self.n_dim, self.num_features, self.kwargs = n_dim, num_features, kwargs
        self.norm = nn.BatchNorm1d(num_features=num_features, **kwargs)

    def forward(self, x):
        scale, shift = _norm_scale_shift(self, x)
        return _apply_norm(x, scale, shift)


class AdaIN(nn.Module):
    """Adaptive instance normalisation (normalization_layers.py:5-57): instance statistics, weight and bias from an embedding through
    a small MLP (host-sized: embed_dim -> 512 -> 2 C)."""

    def __init__(self, embed_dim, in_channels, mlp=None, eps=1e-5):
        super().__init__()
        self.in_channels, self.embed_dim, self.eps = in_channels, embed_dim, eps
        self.mlp = mlp if mlp is not None else nn.Sequential(nn.Linear(embed_dim, 512), nn.GELU(), nn.Linear(512, 2 * in_channels))
        self.embedding = None

    def set_embedding(self, x):
        self.embedding = x.reshape(self.embed_dim,)

    def forward(self, x):
        scale, shift = _norm_scale_shift(self, x)
        return _apply_norm(x, scale, shift)


class InstanceNorm(nn.Module):
    """Parameter-free instance normalisation (normalization_layers.py:60-96) for the fused block: statistics per (sample, channel)."""

    def __init__(self, **kwargs):
        super().__init__()
        extra = set(kwargs) - {"eps"}
        if extra:
            raise NotImplementedError(f"InstanceNorm: only eps is supported, got {sorted(extra)}")
        self.eps = float(kwargs.get("eps", 1e-5))
        self.kwargs = kwargs

    def forward(self, x):
        scale, shift = _norm_scale_shift(self, x)
        return _apply_norm(x, scale, shift)


class _Dropout(torch.autograd.Function):
    """F.dropout(x, p, training=True) (channel_mlp.py:54-58, 110-111): the Bernoulli mask comes from torch's generator --
    `F.dropout(ones)` draws exactly the mask `F.dropout(x)` would, so a run seeded like the reference drops the same elements --
    and is applied (forward and backward) by the pointwise product kernel."""

    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        scaled_mask = F.dropout(torch.ones_like(x), p, True)             # mask / (1 - p)
        ctx.save_for_backward(scaled_mask)
        return _Dropout._mul(x, scaled_mask)

    @staticmethod
    def _mul(a, b):
        lib = _lib.load()
        out = torch.empty_like(a)
        if a.numel():
            with torch.cuda.device(a.device):
                _lib.check(lib.sc_pointwise(_lib.POINTWISE_MUL, _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream_ptr(a.device)), "sc_pointwise")
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (scaled_mask,) = ctx.saved_tensors
        return _Dropout._mul(g.contiguous().float(), scaled_mask), None


# --------------------------------------------------------------------------------------------------
# parameter containers with the reference's names (state-dict compatible) and fused forwards
# --------------------------------------------------------------------------------------------------
class SoftGating(nn.Module):
    """`x * w` with w of shape (1, C, 1, ..) (skip_connections.py:53-93)."""

    def __init__(self, in_features, out_features=None, n_dim=2, bias=False):
        super().__init__()
        if out_features is not None and in_features != out_features:
            raise ValueError(f"Got in_features={in_features} and out_features={out_features}, "
                             "but these two must be the same for soft-gating")
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.ones(1, self.in_features, *(1,) * n_dim))
        self.bias = nn.Parameter(torch.ones(1, self.in_features, *(1,) * n_dim)) if bias else None

    def forward(self, x):
        return channel_mix(bias=self.bias, gate=self.weight, gated=x)


class Flattened1dConv(nn.Module):
    """The "linear" skip: a Conv1d with kernel size 1 over the flattened grid (skip_connections.py:96-130)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size=1, bias=False):
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError("Flattened1dConv: only kernel_size=1 (a pointwise channel mixing) is built")
        self.conv = nn.Conv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, bias=bias)     # parameter container

    def forward(self, x):
        return channel_mix(x, self.conv.weight, self.conv.bias)


def skip_connection(in_features, out_features, n_dim=2, bias=False, skip_type="soft-gating"):
    """Same wrapper as skip_connections.py:5-50."""
    kind = skip_type.lower()
    if kind == "soft-gating":
        return SoftGating(in_features=in_features, out_features=out_features, bias=bias, n_dim=n_dim)
    if kind == "linear":
        return Flattened1dConv(in_channels=in_features, out_channels=out_features, kernel_size=1, bias=bias)
    if kind == "identity":
        return nn.Identity()
    raise ValueError(f"Got skip-connection type={skip_type}, expected one of {'soft-gating', 'linear', 'id'}.")


class ChannelMLP(nn.Module):
    """Pointwise MLP over the channels (channel_mlp.py:6-119): every layer is one fused launch (mixing + bias + GELU); the last one
    can also take the block's skip term and final activation (`_forward_fused`)."""

    def __init__(self, in_channels, out_channels=None, hidden_channels=None, n_layers=2, n_dim=2, non_linearity=F.gelu, dropout=0.0):
        super().__init__()
        self._act = activation_code(non_linearity)
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.hidden_channels = in_channels if hidden_channels is None else hidden_channels
        self.non_linearity = non_linearity
        self.dropout = nn.ModuleList([nn.Dropout(dropout) for _ in range(n_layers)]) if dropout > 0.0 else None
        self.fcs = nn.ModuleList()                                                       # Conv1d modules as parameter containers
        for i in range(n_layers):
            cin = self.in_channels if i == 0 else self.hidden_channels
            cout = self.out_channels if i == n_layers - 1 else self.hidden_channels
            self.fcs.append(nn.Conv1d(cin, cout, 1))

    def _forward_fused(self, x, gate=None, gated=None, final_act=ACT_IDENTITY):
        dropping = self.dropout is not None and self.training and self.dropout[0].p > 0.0
        for i, fc in enumerate(self.fcs):
            last = i == self.n_layers - 1
            if not last:
                x = channel_mix(x, fc.weight, fc.bias, act=self._act)
            elif dropping:
                # the reference drops the output of the last layer BEFORE the block adds its skip (channel_mlp.py:104-111): the skip
                # and the final activation get a launch of their own
                x = channel_mix(x, fc.weight, fc.bias)
            else:
                x = channel_mix(x, fc.weight, fc.bias, gate=gate, gated=gated, act=final_act)
            if dropping:
                x = _Dropout.apply(x, self.dropout[i].p)
        if dropping and (gated is not None or final_act != ACT_IDENTITY):
            x = channel_mix(add=x, gate=gate, gated=gated, act=final_act)
        return x

    def forward(self, x):
        return self._forward_fused(x)


# --------------------------------------------------------------------------------------------------
# complex-valued data at block level (neuralop/layers/complex.py): every layer-epilogue op on the (..., 2) real view of a complex
# tensor -- the kernels are dimension-agnostic, so the trailing (re, im) axis is just one more grid axis
# --------------------------------------------------------------------------------------------------
class _AddITimes(torch.autograd.Function):
    """view_as_complex(a) + 1j * view_as_complex(b) on real (..., 2) views: how `apply_complex` (complex.py:55-62) recombines."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        _require_device_tensor(a, "complex recombination operand")
        _require_device_tensor(b, "complex recombination operand")
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        if a.numel():
            with torch.cuda.device(a.device):
                _lib.check(lib.sc_pointwise(_lib.POINTWISE_ADD_I_TIMES, _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream_ptr(a.device)),
                           "sc_pointwise")
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        g = g.contiguous().float()
        gb = torch.empty_like(g)
        if g.numel():
            with torch.cuda.device(g.device):
                _lib.check(lib.sc_pointwise(_lib.POINTWISE_MUL_NEG_I, _ptr(g), None, _ptr(gb), g.numel(), _stream_ptr(g.device)), "sc_pointwise")
        return g, gb


def _as_real(z: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(z.contiguous()) if z.is_complex() else z


class ComplexValued(nn.Module):
    """`ComplexValued(module)` of the reference (complex.py:65-79): a real module applied to complex data as
    (fr(x.real) - fi(x.imag)) + 1j (fr(x.imag) + fi(x.real)), with `fr` / `fi` two copies of the module (same parameter names).
    Here each copy runs ONCE on the (..., 2) real view of x -- which yields fr / fi of the real and the imaginary part side by
    side -- and the two results are recombined as a + 1j * b by one pointwise launch."""

    def __init__(self, module):
        super().__init__()
        import copy
        self.fr = copy.deepcopy(module)
        self.fi = copy.deepcopy(module)

    def forward(self, x):
        xv = _as_real(x)
        return torch.view_as_complex(_AddITimes.apply(self.fr(xv), self.fi(xv)))


def _complex_act(z, act):
    """CGELU (complex.py:12-31): the activation on the real and the imaginary part separately = on the real view."""
    return z if act == ACT_IDENTITY else torch.view_as_complex(channel_mix(add=_as_real(z), act=act))


def _complex_add_act(z1, z2, act):
    """act(z1 + z2) in one launch on the real views."""
    return torch.view_as_complex(channel_mix(add=_as_real(z1), gated=_as_real(z2), act=act))


# --------------------------------------------------------------------------------------------------
# FNOBlocks
# --------------------------------------------------------------------------------------------------
def _validate_scaling_factor(factor, n_dim, n_layers):
    """neuralop/utils.py:151-197 with n_layers given: a list (per layer) of lists (per dim) of floats, or None."""
    if factor is None:
        return None
    if isinstance(factor, (float, int)):
        return [[float(factor)] * n_dim] * n_layers
    if isinstance(factor, list) and len(factor) > 0 and all(isinstance(s, (float, int)) for s in factor):
        return [[float(s)] * n_dim for s in factor]
    if isinstance(factor, list) and len(factor) > 0 and all(isinstance(s, list) for s in factor):
        return [[float(v) for v in s] for s in factor]
    return None


class FNOBlocks(nn.Module):
    """Drop-in for `neuralop.layers.fno_block.FNOBlocks` (fno_block.py:46-470) on the B200 kernels: n_layers Fourier layers, each
    SpectralConv + skip + (ChannelMLP + skip), applied one at a time with `forward(x, index)`.  Same arguments; unsupported choices
    raise NotImplementedError here instead of silently running something else."""

    def __init__(
        self,
        in_channels,
        out_channels,
        n_modes,
        resolution_scaling_factor=None,
        n_layers=1,
        max_n_modes=None,
        fno_block_precision="full",
        use_channel_mlp=True,
        channel_mlp_dropout=0,
        channel_mlp_expansion=0.5,
        non_linearity=F.gelu,
        stabilizer=None,
        norm=None,
        norm_groups=1,
        ada_in_features=None,
        preactivation=False,
        fno_skip="linear",
        conv_bias_kernel=1,
        channel_mlp_skip="soft-gating",
        complex_data=False,
        separable=False,
        factorization=None,
        rank=1.0,
        conv_module=SpectralConv,
        fixed_rank_modes=False,
        implementation="factorized",
        decomposition_kwargs=dict(),
        enforce_hermitian_symmetry=True,
    ):
        super().__init__()
        if isinstance(n_modes, int):
            n_modes = [n_modes]
        self._n_modes = n_modes
        self.n_dim = len(n_modes)
        if norm not in (None, "instance_norm", "group_norm", "batch_norm", "ada_in"):
            raise ValueError(f"Got norm={norm} but expected None or one of [instance_norm, group_norm, batch_norm, ada_in]")
        if complex_data and norm is not None:
            raise NotImplementedError("FNOBlocks(complex_data=True) with a normalisation layer is not built")
        if complex_data and resolution_scaling_factor is not None:
            raise NotImplementedError("FNOBlocks(complex_data=True) with a resolution change is not built (the reference's `resample` "
                                      "interpolates real tensors)")
        self._act = ACT_GELU if complex_data else activation_code(non_linearity)      # (complex data: CGELU whatever was passed, :204-207)
        if conv_bias_kernel < 1:
            raise ValueError(f"conv_bias_kernel must be >= 1, got {conv_bias_kernel}")
        if conv_bias_kernel != 1:
            if fno_skip is None or fno_skip.lower() != "linear":
                raise ValueError("conv_bias_kernel can only differ from 1 when fno_skip='linear'.")
            if self.n_dim > 3:
                raise NotImplementedError("conv_bias_kernel > 1 is only implemented for 1D, 2D, and 3D FNO blocks.")
            if complex_data:
                raise NotImplementedError("FNOBlocks(complex_data=True) with conv_bias_kernel > 1 is not built")
        if stabilizer not in (None, "tanh"):
            raise ValueError(f"unknown stabilizer {stabilizer!r}")
        for name, kind in (("fno_skip", fno_skip), ("channel_mlp_skip", channel_mlp_skip)):
            if kind is not None and kind.lower() not in ("linear", "soft-gating", "identity"):
                raise ValueError(f"Got {name}={kind}, expected one of 'soft-gating', 'linear', 'identity' or None")
        self.resolution_scaling_factor = _validate_scaling_factor(resolution_scaling_factor, self.n_dim, n_layers)
        self.max_n_modes = max_n_modes
        self.fno_block_precision = fno_block_precision
        self.in_channels, self.out_channels, self.n_layers = in_channels, out_channels, n_layers
        self.stabilizer, self.rank, self.factorization = stabilizer, rank, factorization
        self.fixed_rank_modes, self.decomposition_kwargs = fixed_rank_modes, decomposition_kwargs
        self.fno_skip = fno_skip.lower() if fno_skip is not None else None
        self.conv_bias_kernel = conv_bias_kernel
        self.channel_mlp_skip = channel_mlp_skip.lower() if channel_mlp_skip is not None else None
        self.complex_data = complex_data
        self.use_channel_mlp = use_channel_mlp
        self.channel_mlp_expansion, self.channel_mlp_dropout = channel_mlp_expansion, channel_mlp_dropout
        self.implementation, self.separable, self.preactivation = implementation, separable, preactivation
        self.ada_in_features, self.enforce_hermitian_symmetry = ada_in_features, enforce_hermitian_symmetry
        self.non_linearity = non_linearity
        self.n_norms = 2
        if norm is None:
            self.norm = None
        elif norm == "instance_norm":
            self.norm = nn.ModuleList([InstanceNorm() for _ in range(n_layers * self.n_norms)])
        elif norm == "group_norm":   # nn.GroupNorm modules as parameter containers (their forward is never called)
            self.norm = nn.ModuleList([nn.GroupNorm(num_groups=norm_groups, num_channels=self.out_channels)
                                       for _ in range(n_layers * self.n_norms)])
        elif norm == "batch_norm":
            self.norm = nn.ModuleList([BatchNorm(n_dim=self.n_dim, num_features=self.out_channels) for _ in range(n_layers * self.n_norms)])
        else:
            self.norm = nn.ModuleList([AdaIN(ada_in_features, out_channels) for _ in range(n_layers * self.n_norms)])

        self.convs = nn.ModuleList([
            conv_module(
                self.in_channels, self.out_channels, self.n_modes,
                resolution_scaling_factor=(self.resolution_scaling_factor[i] if resolution_scaling_factor is not None else None),
                max_n_modes=max_n_modes, rank=rank, fixed_rank_modes=fixed_rank_modes, implementation=implementation,
                separable=separable, factorization=factorization, fno_block_precision=fno_block_precision,
                decomposition_kwargs=decomposition_kwargs, complex_data=complex_data,
                **({"enforce_hermitian_symmetry": enforce_hermitian_symmetry} if issubclass(conv_module, SpectralConv) else {}),
            )
            for i in range(n_layers)
        ])
        if self.fno_skip is not None and conv_bias_kernel != 1:
            # a LOCAL convolution as the skip (fno_block.py:18-43): the one piece of the block that is a library call here (torch's
            # Conv{n}d = cuDNN, like the reference); its output joins the fused add + activation launch as a materialised skip
            self.fno_skips = nn.ModuleList([getattr(nn, f"Conv{self.n_dim}d")(self.in_channels, self.out_channels,
                                                                               kernel_size=conv_bias_kernel, padding="same", bias=False)
                                            for _ in range(n_layers)])
        elif self.fno_skip is not None:
            self.fno_skips = nn.ModuleList([skip_connection(self.in_channels, self.out_channels, skip_type=self.fno_skip, n_dim=self.n_dim)
                                            for _ in range(n_layers)])
        else:
            self.fno_skips = None
        if self.use_channel_mlp:
            self.channel_mlp = nn.ModuleList([
                ChannelMLP(in_channels=self.out_channels, hidden_channels=round(self.out_channels * channel_mlp_expansion),
                           dropout=channel_mlp_dropout, n_dim=self.n_dim)
                for _ in range(n_layers)])
            if self.channel_mlp_skip is not None:
                self.channel_mlp_skips = nn.ModuleList([
                    skip_connection(self.in_channels, self.out_channels, skip_type=self.channel_mlp_skip, n_dim=self.n_dim)
                    for _ in range(n_layers)])
            else:
                self.channel_mlp_skips = None
        if self.complex_data:          # fno_block.py:275-276, 293-311: every epilogue module becomes a ComplexValued pair (fr, fi)
            if self.fno_skips is not None:
                self.fno_skips = nn.ModuleList([ComplexValued(m) for m in self.fno_skips])
            if self.use_channel_mlp:
                self.channel_mlp = nn.ModuleList([ComplexValued(m) for m in self.channel_mlp])
                if self.channel_mlp_skips is not None:
                    self.channel_mlp_skips = nn.ModuleList([ComplexValued(m) for m in self.channel_mlp_skips])

    def set_ada_in_embeddings(self, *embeddings):
        """Sets the embeddings of the AdaIN layers (fno_block.py:354-369): one for all, or one per norm layer."""
        if self.norm is not None:
            if len(embeddings) == 1:
                for norm in self.norm:
                    norm.set_embedding(embeddings[0])
            else:
                for norm, embedding in zip(self.norm, embeddings):
                    norm.set_embedding(embedding)

    # -- helpers ------------------------------------------------------------------------------------
    @staticmethod
    def _skip_terms(kind, module, x):
        """The skip as (x_mix, weight, gate, gated) operands of the fused op, without materialising it."""
        if kind == "linear":
            if not isinstance(module, Flattened1dConv) or module.conv.bias is not None:
                return None          # a local convolution (conv_bias_kernel > 1) or a skip with a bias: materialise
            return x, module.conv.weight, None, None
        if kind == "soft-gating":
            if module.bias is not None:
                return None
            return None, None, module.weight, x
        return None, None, None, x   # identity

    def _resamples(self, conv, x, output_shape):
        grid = list(x.shape[2:])
        if hasattr(conv, "_output_grid"):
            return [int(s) for s in conv._output_grid(grid, output_shape)] != grid
        return output_shape is not None and list(output_shape) != grid

    def _fourier_step(self, x, index, output_shape, act, norm=None):
        """f1: act( norm(conv(stabilizer(x))) + fno_skip(x) )."""
        conv = self.convs[index]
        x_conv = _Tanh.apply(x) if self.stabilizer == "tanh" else x
        x_fno = conv(x_conv, output_shape=output_shape)
        if norm is not None:
            # the normalisation is a per-(sample, channel) affine map of the conv output: one pass for its statistics, then ONE launch
            # for  act( scale * x_fno + shift + skip )  -- the skip is materialised first (its mixing is per channel, the map per row)
            scale, shift = _norm_scale_shift(norm, x_fno)
            x_skip = None
            if self.fno_skips is not None:
                x_skip = conv.transform(self.fno_skips[index](x), output_shape=output_shape)
            return _apply_norm(x_fno, scale, shift, add=x_skip, act=act)
        if self.fno_skips is None:
            return x_fno if act == ACT_IDENTITY else channel_mix(add=x_fno, act=act)
        terms = None if self._resamples(conv, x, output_shape) else self._skip_terms(self.fno_skip, self.fno_skips[index], x)
        if terms is None:
            # the skip lives on the input grid and the conv changed the resolution (or the skip carries a bias): materialise it,
            # resample it as the reference does (`convs[i].transform`, fno_block.py:380), then ONE launch for add + activation
            x_skip = conv.transform(self.fno_skips[index](x), output_shape=output_shape)
            return channel_mix(add=x_fno, gated=x_skip, act=act)
        xm, w, gate, gated = terms
        return channel_mix(xm, w, add=x_fno, gate=gate, gated=gated, act=act)

    def _mlp_step(self, x1, x, index, output_shape, act):
        """f2: act( channel_mlp(x1) + channel_mlp_skip(x) )."""
        mlp = self.channel_mlp[index]
        if self.channel_mlp_skips is None:
            return mlp._forward_fused(x1, final_act=act)
        conv = self.convs[index]
        kind, module = self.channel_mlp_skip, self.channel_mlp_skips[index]
        if kind == "linear" or self._resamples(conv, x, output_shape) or (kind == "soft-gating" and module.bias is not None):
            x_skip = conv.transform(module(x), output_shape=output_shape)
            return mlp._forward_fused(x1, gated=x_skip, final_act=act)
        gate = module.weight if kind == "soft-gating" else None
        return mlp._forward_fused(x1, gate=gate, gated=x, final_act=act)

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, x, index=0, output_shape=None):
        if self.complex_data:
            return self._forward_complex(x, index, output_shape)
        if self.preactivation:
            return self.forward_with_preactivation(x, index, output_shape)
        return self.forward_with_postactivation(x, index, output_shape)

    def _forward_complex(self, x, index, output_shape):
        """Both forward orders for complex data (CGELU, ctanh, ComplexValued skips / MLP; fno_block.py:377-453 with :204-207, :275-311):
        the same sequence as the real case with every epilogue op on the (..., 2) real view."""
        _require_complex_input(x)
        if output_shape is not None and list(output_shape) != list(x.shape[2:]):
            raise NotImplementedError("FNOBlocks(complex_data=True) with a resolution change is not built")
        act = ACT_GELU if index < (self.n_layers - 1) else ACT_IDENTITY
        x = x.contiguous()
        if self.preactivation:
            x = _complex_act(x, ACT_GELU)
        x_skip_fno = self.fno_skips[index](x) if self.fno_skips is not None else None
        x_skip_mlp = self.channel_mlp_skips[index](x) if self.use_channel_mlp and self.channel_mlp_skips is not None else None
        xc = torch.view_as_complex(_Tanh.apply(_as_real(x))) if self.stabilizer == "tanh" else x         # ctanh, complex.py:45-52
        x_fno = self.convs[index](xc, output_shape=output_shape)
        y = _complex_add_act(x_fno, x_skip_fno, act) if x_skip_fno is not None else _complex_act(x_fno, act)
        if self.use_channel_mlp:
            m = self.channel_mlp[index](y)
            final = ACT_IDENTITY if self.preactivation else act
            return _complex_add_act(m, x_skip_mlp, final) if x_skip_mlp is not None else _complex_act(m, final)
        return y if self.preactivation else _complex_act(y, act)

    def forward_with_postactivation(self, x, index=0, output_shape=None):
        """fno_block.py:377-414."""
        _require_device_tensor(x, "FNOBlocks input")
        x = x.contiguous()
        act = self._act if index < (self.n_layers - 1) else ACT_IDENTITY
        if self.norm is not None:
            x1 = self._fourier_step(x, index, output_shape, act, norm=self.norm[self.n_norms * index])
            pre = self._mlp_step(x1, x, index, output_shape, ACT_IDENTITY) if self.use_channel_mlp else x1
            scale, shift = _norm_scale_shift(self.norm[self.n_norms * index + 1], pre)           # fno_block.py:408-412
            return _apply_norm(pre, scale, shift, act=act)
        x1 = self._fourier_step(x, index, output_shape, act)
        if self.use_channel_mlp:
            return self._mlp_step(x1, x, index, output_shape, act)
        # no channel MLP: the reference applies the non-linearity a second time (fno_block.py:411-412)
        return x1 if act == ACT_IDENTITY else channel_mix(add=x1, act=act)

    def forward_with_preactivation(self, x, index=0, output_shape=None):
        """fno_block.py:416-453: activation first, then conv + skip (+ activation unless last), then the channel MLP + skip."""
        _require_device_tensor(x, "FNOBlocks input")
        x = channel_mix(add=x.contiguous(), act=self._act)
        if self.norm is not None:                                                                # fno_block.py:421-422
            x = _apply_norm(x, *_norm_scale_shift(self.norm[self.n_norms * index], x))
        act = self._act if index < (self.n_layers - 1) else ACT_IDENTITY
        x1 = self._fourier_step(x, index, output_shape, act)
        if self.norm is not None:                                                                # fno_block.py:444-445
            x1 = _apply_norm(x1, *_norm_scale_shift(self.norm[self.n_norms * index + 1], x1))
        if self.use_channel_mlp:
            return self._mlp_step(x1, x, index, output_shape, ACT_IDENTITY)
        return x1

    @property
    def n_modes(self):
        return self._n_modes

    @n_modes.setter
    def n_modes(self, n_modes):
        for i in range(self.n_layers):
            self.convs[i].n_modes = n_modes
        self._n_modes = n_modes

    def get_block(self, indices):
        """One layer of the jointly parametrised block as a module of its own (fno_block.py:466-479): shares the parameters."""
        if self.n_layers == 1:
            raise ValueError("A single layer is parametrized, directly use the main class.")
        return LayerView(self, indices)

    def __getitem__(self, indices):
        return self.get_block(indices)


class LayerView(nn.Module):
    """`FNOBlocks[i]`: forward(x) = blocks.forward(x, i) (the reference's SubModule, fno_block.py:482-500)."""

    def __init__(self, blocks: FNOBlocks, index: int):
        super().__init__()
        self.main_module = blocks
        self.indices = index

    def forward(self, x, output_shape=None):
        return self.main_module.forward(x, self.indices, output_shape=output_shape)
