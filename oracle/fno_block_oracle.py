"""CPU oracle for the Fourier layer around the spectral convolution.  TEST INFRASTRUCTURE --
NOT PRODUCT CODE: only `tests/`, `bench.py`'s baseline legs and `oracle/make_golden_block.py` import it.

What it restates (reference = neuraloperator @ 93d3f06, paths relative to /root/reference):

  * `fno_block_forward`   <- neuralop/layers/fno_block.py:377-414 (`forward_with_postactivation`) and :416-453
                             (`forward_with_preactivation`), for norm in {None, instance_norm, group_norm}, real data, non_linearity=F.gelu
  * `_skip`               <- neuralop/layers/skip_connections.py:85-93 (SoftGating.forward), :118-130 (Flattened1dConv.forward),
                             nn.Identity
  * `_channel_mlp`        <- neuralop/layers/channel_mlp.py:92-116 (ChannelMLP.forward, dropout = 0)
  * the conv              <- `oracle.spectral_conv_oracle.spectral_conv_forward` (pinned on its own)
  * the skip resampling   <- `SpectralConv.transform` (spectral_convolution.py:383-398) -> `resample_restated`

The backward is whatever torch.autograd records for these CPU ops -- as in the reference, which has no hand-written backward.
Pinning: `tests/test_block_oracle.py` compares this restatement with golden vectors minted by `oracle/make_golden_block.py` from the
UNMODIFIED `FNOBlocks` (imported through `oracle/load_reference.py`) and, where /root/reference exists, with the live class.
Parameters are passed as a dict keyed by the reference's state-dict names (`fno_skips.0.conv.weight`, `channel_mlp.0.fcs.1.bias`, ...).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from . import spectral_conv_oracle as O


def _skip(kind: Optional[str], params: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    if kind == "linear" and prefix + ".weight" in params:                      # conv_bias_kernel > 1: Conv{n}d, padding="same", no bias (:36-43)
        return getattr(F, f"conv{x.ndim - 2}d")(x, params[prefix + ".weight"], padding="same")
    if kind == "linear":
        w = params[prefix + ".conv.weight"]                                   # (Co, Ci, 1), no bias (skip_connection default)
        size = list(x.shape)
        return F.conv1d(x.reshape(size[0], size[1], -1), w).view(size[0], w.shape[0], *size[2:])
    if kind == "soft-gating":
        return params[prefix + ".weight"] * x
    if kind == "identity":
        return x
    raise ValueError(kind)


def _channel_mlp(params: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, n_layers: int = 2) -> torch.Tensor:
    size = list(x.shape)
    h = x.reshape(size[0], size[1], -1)
    for i in range(n_layers):
        h = F.conv1d(h, params[f"{prefix}.fcs.{i}.weight"], params[f"{prefix}.fcs.{i}.bias"])
        if i < n_layers - 1:
            h = F.gelu(h)
    return h.reshape(size[0], h.shape[1], *size[2:])


def _norm(kind: Optional[str], params: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """InstanceNorm.forward (normalization_layers.py:91-96) / nn.GroupNorm(norm_groups, C) (fno_block.py:318-326)."""
    if kind == "instance_norm":
        return F.instance_norm(x)
    if kind == "group_norm":
        return F.group_norm(x, groups, params[prefix + ".weight"], params[prefix + ".bias"])
    if kind == "batch_norm":      # BatchNorm.forward (normalization_layers.py:146-158) in TRAINING mode: batch statistics
        return F.batch_norm(x, None, None, params[prefix + ".norm.weight"], params[prefix + ".norm.bias"], training=True, eps=1e-5)
    if kind == "ada_in":          # AdaIN.forward (:51-57): weight, bias = split(mlp(embedding)); group_norm with C groups
        e = params["ada_in_embedding"]
        h = F.gelu(F.linear(e, params[prefix + ".mlp.0.weight"], params[prefix + ".mlp.0.bias"]))
        wb = F.linear(h, params[prefix + ".mlp.2.weight"], params[prefix + ".mlp.2.bias"])
        C = x.shape[1]
        return F.group_norm(x, C, wb[:C], wb[C:], eps=1e-5)
    raise ValueError(kind)


def _transform(x, in_grid, out_grid):
    return x if list(in_grid) == list(out_grid) else O.resample_restated(x, out_grid)


def conv_weight_from_params(params: Dict[str, torch.Tensor], index: int, kind: str = "dense") -> O.Weight:
    """The oracle's Weight for layer `index` from reference-named parameters (`convs.<i>.weight.tensor` / `.core` / `.factors.<j>`)."""
    pre = f"convs.{index}.weight."
    fs = sorted((k for k in params if k.startswith(pre + "factors.")), key=lambda k: int(k.rsplit(".", 1)[1].replace("factor_", "")))
    factors = [params[k] for k in fs]
    if kind == "dense":
        return O.Weight("dense", tensor=params[pre + "tensor"])
    if kind == "tucker":
        return O.Weight("tucker", core=params[pre + "core"], factors=factors)
    if kind == "cp":
        return O.Weight("cp", weights=params[pre + "weights"], factors=factors)
    if kind == "tt":
        return O.Weight("tt", factors=factors)
    raise ValueError(kind)


def fno_block_forward(x: torch.Tensor, params: Dict[str, torch.Tensor], index: int, *, n_modes: Sequence[int], n_layers: int,
                      weight_kind: str = "dense", fno_skip: Optional[str] = "linear", channel_mlp_skip: Optional[str] = "soft-gating",
                      use_channel_mlp: bool = True, stabilizer: Optional[str] = None, preactivation: bool = False,
                      output_shape: Optional[Sequence[int]] = None, resolution_scaling_factor=None,
                      max_n_modes: Optional[Sequence[int]] = None, norm: Optional[str] = None, norm_groups: int = 1,
                      non_linearity=F.gelu) -> torch.Tensor:
    """One Fourier layer: `FNOBlocks.forward(x, index, output_shape)` for real data (the ChannelMLP inside keeps GELU, fno_block.py:280-290)."""
    grid = list(x.shape[2:])
    rsf = resolution_scaling_factor
    if rsf is not None and not isinstance(rsf, (list, tuple)):
        rsf = [float(rsf)] * len(grid)
    out_grid = O.resolve_output_grid(grid, rsf, output_shape)
    nonlin = index < n_layers - 1
    if preactivation:
        x = non_linearity(x)                                                    # :419
        if norm is not None:
            x = _norm(norm, params, f"norm.{2 * index}", x, norm_groups)         # :421-422
    x_skip_fno = None
    if fno_skip is not None:                                                    # :378-380 / :424-426
        x_skip_fno = _transform(_skip(fno_skip, params, f"fno_skips.{index}", x), grid, out_grid)
    x_skip_mlp = None
    if use_channel_mlp and channel_mlp_skip is not None:                        # :382-384 / :428-430
        x_skip_mlp = _transform(_skip(channel_mlp_skip, params, f"channel_mlp_skips.{index}", x), grid, out_grid)
    xc = torch.tanh(x) if stabilizer == "tanh" else x                           # :386-390
    w = conv_weight_from_params(params, index, weight_kind)
    x_fno = O.spectral_conv_forward(xc, w, params.get(f"convs.{index}.bias"), n_modes, max_n_modes=max_n_modes,
                                    output_shape=output_shape, resolution_scaling_factor=rsf)      # :392
    if norm is not None and not preactivation:
        x_fno = _norm(norm, params, f"norm.{2 * index}", x_fno, norm_groups)     # :394-395
    y = x_fno + x_skip_fno if x_skip_fno is not None else x_fno                 # :397
    if nonlin:
        y = non_linearity(y)                                                    # :399-400
    if norm is not None and preactivation:
        y = _norm(norm, params, f"norm.{2 * index + 1}", y, norm_groups)         # :444-445
    if use_channel_mlp:                                                         # :402-406
        y = _channel_mlp(params, f"channel_mlp.{index}", y)
        if x_skip_mlp is not None:
            y = y + x_skip_mlp
    if norm is not None and not preactivation:
        y = _norm(norm, params, f"norm.{2 * index + 1}", y, norm_groups)         # :408-409
    if nonlin and not preactivation:                                            # :411-412 (the pre-activation form ends without it)
        y = non_linearity(y)
    return y


def _cgelu(z):                                                                  # complex.py:12-31
    return torch.complex(F.gelu(z.real), F.gelu(z.imag))


def _complex_valued(fr, fi, z):                                                 # apply_complex, complex.py:55-62
    return torch.complex(fr(z.real) - fi(z.imag), fr(z.imag) + fi(z.real))


def fno_block_forward_complex(x: torch.Tensor, params: Dict[str, torch.Tensor], index: int, *, n_modes: Sequence[int], n_layers: int,
                              fno_skip: Optional[str] = "linear", channel_mlp_skip: Optional[str] = "soft-gating",
                              use_channel_mlp: bool = True, stabilizer: Optional[str] = None, preactivation: bool = False,
                              max_n_modes: Optional[Sequence[int]] = None) -> torch.Tensor:
    """`FNOBlocks.forward(x, index)` with complex_data=True (fno_block.py:204-207 CGELU, :275-276 / :293-311 ComplexValued skips and
    MLP, :386-388 ctanh), dense conv weight, norm=None, no resolution change."""
    nonlin = index < n_layers - 1

    def cv(kind_fn, prefix):
        return lambda z: _complex_valued(lambda t: kind_fn(prefix + ".fr", t), lambda t: kind_fn(prefix + ".fi", t), z)

    if preactivation:
        x = _cgelu(x)
    x_skip_fno = cv(lambda pre, t: _skip(fno_skip, params, pre, t), f"fno_skips.{index}")(x) if fno_skip is not None else None
    x_skip_mlp = None
    if use_channel_mlp and channel_mlp_skip is not None:
        x_skip_mlp = cv(lambda pre, t: _skip(channel_mlp_skip, params, pre, t), f"channel_mlp_skips.{index}")(x)
    xc = torch.complex(torch.tanh(x.real), torch.tanh(x.imag)) if stabilizer == "tanh" else x
    x_fno = O.spectral_conv_forward_complex(xc, params[f"convs.{index}.weight.tensor"], params.get(f"convs.{index}.bias"), list(n_modes),
                                            max_n_modes=max_n_modes)
    y = x_fno + x_skip_fno if x_skip_fno is not None else x_fno
    if nonlin:
        y = _cgelu(y)
    if use_channel_mlp:
        y = cv(lambda pre, t: _channel_mlp(params, pre, t), f"channel_mlp.{index}")(y)
        if x_skip_mlp is not None:
            y = y + x_skip_mlp
    if nonlin and not preactivation:
        y = _cgelu(y)
    return y


def fno_block_fwd_bwd(x, params, index, grad_y, **kw):
    """Forward + autograd backward. Returns y, dx, {name: grad} for every parameter the layer touched."""
    x = x.detach().clone().requires_grad_(True)
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    y = (fno_block_forward_complex if x.is_complex() else fno_block_forward)(x, ps, index, **kw)
    y.backward(grad_y)
    return y.detach(), x.grad, {k: v.grad for k, v in ps.items() if v.grad is not None}
