"""Switching an existing reference model over to the B200 kernels.

`use_b200_layers(model)` takes an instance of the reference's `neuralop.models.FNO` / `TFNO` (or any module that holds a reference
`FNOBlocks` / `ChannelMLP`) and replaces, in place, every sub-module this package has a drop-in for:

    neuralop.layers.fno_block.FNOBlocks        -> neuraloperator_b200.FNOBlocks   (SpectralConv + fused layer epilogue)
    neuralop.layers.channel_mlp.ChannelMLP     -> neuraloperator_b200.ChannelMLP  (lifting / projection: one fused launch per layer)
    neuralop.layers.skip_connections.Flattened1dConv / SoftGating -> the same-named drop-ins (stand-alone skips, e.g. UNO's)
    neuralop.layers.spectral_convolution.SpectralConv -> neuraloperator_b200.SpectralConv  (wherever the block around it has no
                                                  drop-in -- norm layers, complex data, dropout: the block stays, its convs move over)

Parameters are carried over by name (`load_state_dict`: the drop-ins use the reference's parameter names), so a trained checkpoint
keeps working; the model's own `forward` (positional embedding, domain padding, the loop over layers, fno.py:346-406) stays the
reference's Python.  Modules are recognised by class NAME and constructor attributes, so this file does not import the reference.
"""
import warnings

import torch.nn.functional as F
from torch import nn

from .fno_block import ChannelMLP, ComplexValued, Flattened1dConv, FNOBlocks, SoftGating
from .spectral_conv import SpectralConv


def _convert_channel_mlp(ref: nn.Module) -> ChannelMLP:
    drop = ref.dropout[0].p if getattr(ref, "dropout", None) is not None else 0.0      # (eval-mode only on the B200 side)
    new = ChannelMLP(ref.in_channels, out_channels=ref.out_channels, hidden_channels=ref.hidden_channels, n_layers=ref.n_layers,
                     non_linearity=ref.non_linearity, dropout=drop)
    new.train(ref.training)
    new.load_state_dict(ref.state_dict())
    return new


def _convert_fno_blocks(ref: nn.Module) -> FNOBlocks:
    norm, norm_groups = None, 1
    if getattr(ref, "norm", None) is not None:
        kind = type(ref.norm[0]).__name__
        if kind == "InstanceNorm" and not getattr(ref.norm[0], "kwargs", {}):
            norm = "instance_norm"
        elif kind == "GroupNorm":
            norm, norm_groups = "group_norm", ref.norm[0].num_groups
        elif kind == "BatchNorm" and not getattr(ref.norm[0], "kwargs", {}):
            norm = "batch_norm"
        elif kind == "AdaIN":
            norm = "ada_in"
        else:
            raise NotImplementedError(f"FNOBlocks with {kind} normalisation layers has no B200 drop-in")
    rsf = ref.resolution_scaling_factor
    new = FNOBlocks(
        ref.in_channels, ref.out_channels, ref.n_modes, resolution_scaling_factor=rsf, n_layers=ref.n_layers,
        max_n_modes=ref.max_n_modes, fno_block_precision=ref.fno_block_precision, use_channel_mlp=ref.use_channel_mlp,
        channel_mlp_dropout=ref.channel_mlp_dropout, channel_mlp_expansion=ref.channel_mlp_expansion,
        # (for complex data the reference replaces whatever activation was passed by CGELU, fno_block.py:204-207: ours does the same)
        non_linearity=F.gelu if ref.complex_data else ref.non_linearity, stabilizer=ref.stabilizer, norm=norm, norm_groups=norm_groups,
        ada_in_features=ref.ada_in_features, preactivation=ref.preactivation,
        fno_skip=ref.fno_skip, conv_bias_kernel=ref.conv_bias_kernel, channel_mlp_skip=ref.channel_mlp_skip,
        complex_data=ref.complex_data, separable=ref.separable, factorization=ref.factorization, rank=ref.rank,
        fixed_rank_modes=ref.fixed_rank_modes, implementation=ref.implementation, decomposition_kwargs=ref.decomposition_kwargs,
        enforce_hermitian_symmetry=getattr(ref, "enforce_hermitian_symmetry", True))
    state = {}
    for k, v in ref.state_dict().items():                       # tltorch names its factors `factors.factor_<j>`; accept `factors.<j>` too
        parts = k.split(".")
        if len(parts) >= 2 and parts[-2] == "factors" and parts[-1].isdigit():
            parts[-1] = "factor_" + parts[-1]
        state[".".join(parts)] = v
    new.load_state_dict(state)
    if norm == "ada_in":
        for ours, theirs in zip(new.norm, ref.norm):
            ours.embedding = theirs.embedding
    new.train(ref.training)
    return new


def _convert_spectral_conv(ref: nn.Module) -> SpectralConv:
    """A reference `SpectralConv` (spectral_convolution.py:183-570) -> the B200 class with the same configuration and parameters.
    Used when the block around it cannot be replaced as a whole (norm layers, complex data ...): the hot path still moves over."""
    stored = list(ref.n_modes)                                   # the reference stores the last dim already halved (:404-415)
    user_modes = stored if ref.complex_data else stored[:-1] + [(stored[-1] - 1) * 2]
    rsf = ref.resolution_scaling_factor
    new = SpectralConv(ref.in_channels, ref.out_channels, tuple(user_modes), complex_data=ref.complex_data,
                       max_n_modes=list(ref.max_n_modes), bias=ref.bias is not None, separable=ref.separable,
                       resolution_scaling_factor=rsf, fno_block_precision=ref.fno_block_precision, rank=ref.rank,
                       factorization=ref.factorization, implementation=ref.implementation,
                       enforce_hermitian_symmetry=getattr(ref, "enforce_hermitian_symmetry", True), fft_norm=ref.fft_norm)
    state = {}
    for k, v in ref.state_dict().items():
        parts = k.split(".")
        if len(parts) >= 2 and parts[-2] == "factors" and parts[-1].isdigit():
            parts[-1] = "factor_" + parts[-1]
        state[".".join(parts)] = v
    new.load_state_dict(state)
    return new


def _convert_flattened_conv(ref: nn.Module) -> Flattened1dConv:
    """A stand-alone linear skip (skip_connections.py:96-130), e.g. the horizontal skips of UNO."""
    if tuple(ref.conv.kernel_size) != (1,):
        raise NotImplementedError("Flattened1dConv with kernel_size != 1 has no B200 drop-in")
    new = Flattened1dConv(ref.conv.in_channels, ref.conv.out_channels, kernel_size=1, bias=ref.conv.bias is not None)
    new.load_state_dict(ref.state_dict())
    return new


def _convert_soft_gating(ref: nn.Module) -> SoftGating:
    new = SoftGating(ref.in_features, ref.out_features, n_dim=ref.weight.ndim - 2, bias=ref.bias is not None)
    new.load_state_dict(ref.state_dict())
    return new


def _convert_complex_valued(ref: nn.Module) -> ComplexValued:
    """`ComplexValued(module)` (complex.py:65-79, e.g. the lifting / projection of a complex FNO): the pair (fr, fi) of converted
    modules, evaluated once each on the real view of the input."""
    parts = []
    for part in (ref.fr, ref.fi):
        conv = _CONVERTERS.get(type(part).__name__)
        if conv is None or type(part).__name__ == "ComplexValued":
            raise NotImplementedError(f"ComplexValued({type(part).__name__}) has no B200 drop-in")
        parts.append(conv(part))
    new = ComplexValued(nn.Identity())
    new.fr, new.fi = parts
    return new


_CONVERTERS = {"ComplexValued": _convert_complex_valued, "FNOBlocks": _convert_fno_blocks, "ChannelMLP": _convert_channel_mlp, "SpectralConv": _convert_spectral_conv,
               "Flattened1dConv": _convert_flattened_conv, "SoftGating": _convert_soft_gating}


def use_b200_layers(model: nn.Module) -> nn.Module:
    """Replaces, in place and recursively, every reference `FNOBlocks` / `ChannelMLP` inside `model` by its B200 drop-in with the same
    parameters; returns `model`.  Move the model to the GPU afterwards (or before: device and dtype of the parameters are kept)."""
    for name, child in list(model.named_children()):
        conv = _CONVERTERS.get(type(child).__name__)
        if conv is not None and not type(child).__module__.startswith("neuraloperator_b200"):
            ref_param = next(child.parameters(), None)
            try:
                new = conv(child)
            except NotImplementedError as why:
                if type(child).__name__ == "SpectralConv":
                    raise
                # e.g. a block with norm layers: keep the reference module, move what is inside it (its SpectralConvs) over
                warnings.warn(f"use_b200_layers: {name} ({type(child).__name__}) stays the reference module: {why}", stacklevel=2)
                use_b200_layers(child)
                continue
            if ref_param is not None:
                new = new.to(ref_param.device)
            if isinstance(model, (nn.ModuleList, nn.Sequential)):
                model[int(name)] = new
            else:
                setattr(model, name, new)
        else:
            use_b200_layers(child)
    return model
