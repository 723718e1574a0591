"""Model-level integration without a GPU (kernels through their host checks, conv through the CPU oracle, as in
tests/test_block_host_logic.py):
(1) the drop-ins stacked as the reference `FNO` stacks them (lifting -> n_layers x FNOBlocks -> projection) against goldens minted from
    the unmodified reference model (oracle/make_golden_fno.py): y, dx, every parameter gradient;
(2) `neuraloperator_b200.use_b200_layers(model)` on a live reference FNO / TFNO: same outputs and gradients before and after the swap."""
import pytest
import torch

import neuraloperator_b200 as nb
from conftest import build_fno_stack, fno_golden_index, load_fno_golden
from oracle.load_reference import reference_available
from test_block_host_logic import grad_err, host, rel_err  # noqa: F401  (fixture)


@pytest.mark.parametrize("name", sorted(fno_golden_index().keys()))
def test_stacked_drop_ins_match_reference_fno_golden(host, name):  # noqa: F811
    meta, io, params, grads = load_fno_golden(name)
    mods, forward = build_fno_stack(meta, params)
    x = io["x"].clone().requires_grad_(True)
    y = forward(x)
    y.backward(io["gy"])
    assert rel_err(y, io["y"]) < 3e-5 and rel_err(x.grad, io["dx"]) < 3e-5
    ours = dict(mods.named_parameters())
    for k, g in grads.items():
        assert rel_err(ours[k.replace("weight.factors.", "weight.factors.factor_")].grad, g) < 5e-5, k


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kw", [dict(n_modes=(8, 8), in_channels=2, out_channels=3, hidden_channels=8, n_layers=2),
                                dict(n_modes=(8, 6), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, factorization="tucker",
                                     implementation="factorized", rank=[3, 3, 4, 3]),
                                dict(n_modes=(10,), in_channels=1, out_channels=1, hidden_channels=4, n_layers=3, stabilizer="tanh",
                                     fno_skip="soft-gating", channel_mlp_skip="linear"),
                                dict(n_modes=(8, 8), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, norm="group_norm"),
                                dict(n_modes=(8, 8), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, norm="instance_norm"),
                                dict(n_modes=(8, 8), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, norm="batch_norm"),
                                dict(n_modes=(8, 8), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, complex_data=True,
                                     positional_embedding=None),
                                dict(n_modes=(8, 8), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, conv_bias_kernel=3,
                                     channel_mlp_dropout=0.2, non_linearity=torch.nn.functional.silu)])
def test_use_b200_layers_on_a_live_reference_model(host, kw):  # noqa: F811
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__file__)), "oracle"))
    from make_golden_fno import load_reference_fno
    fno = load_reference_fno()
    torch.manual_seed(21)
    model = fno.FNO(**kw)                                  # default positional embedding (grid) + no padding: stays reference code
    if kw.get("channel_mlp_dropout"):
        model.eval()                                       # (training-mode dropout is compared under equal seeds in test_block_host_logic)
    grid = (16,) * len(kw["n_modes"])
    x = torch.randn(2, kw["in_channels"], *grid, dtype=torch.cfloat if kw.get("complex_data") else torch.float32)
    gy = None
    xr = x.clone().requires_grad_(True)
    y_ref = model(xr)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    dx_ref = xr.grad.clone()
    names = sorted(k for k, _ in model.named_parameters())
    model.zero_grad(set_to_none=True)

    out = nb.use_b200_layers(model)
    assert out is model
    mlp_type = nb.ComplexValued if kw.get("complex_data") else nb.ChannelMLP
    assert type(model.fno_blocks) is nb.FNOBlocks and type(model.lifting) is mlp_type and type(model.projection) is mlp_type
    assert sorted(k.replace("factors.factor_", "factors.") for k, _ in model.named_parameters()) == names
    xo = x.clone().requires_grad_(True)
    y = model(xo)
    y.backward(gy)
    assert rel_err(y, y_ref.detach()) < 3e-5 and rel_err(xo.grad, dx_ref) < 3e-5
    for k, p in model.named_parameters():
        assert grad_err(p.grad, ref_grads[k.replace("factors.factor_", "factors.")], ref_grads) < 5e-5, k
    nb.use_b200_layers(model)                              # idempotent: nothing left to convert
    assert type(model.fno_blocks) is nb.FNOBlocks


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_blocks_without_a_drop_in_keep_the_reference_block_and_swap_its_convs(host):  # noqa: F811
    """non_linearity=F.elu (an activation the kernels do not have): the reference FNOBlocks and the lifting / projection MLPs stay, with a
    warning each; the SpectralConvs and the (GELU) ChannelMLPs inside the block move over; outputs and gradients are unchanged."""
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__file__)), "oracle"))
    from make_golden_fno import load_reference_fno
    fno = load_reference_fno()
    torch.manual_seed(5)
    model = fno.FNO(n_modes=(8, 6), in_channels=1, out_channels=1, hidden_channels=8, n_layers=2, non_linearity=torch.nn.functional.elu,
                    max_n_modes=(10, 8))
    x = torch.randn(2, 1, 16, 12)
    xr = x.clone().requires_grad_(True)
    y_ref = model(xr)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    dx_ref = xr.grad.clone()
    model.zero_grad(set_to_none=True)
    nb.use_b200_layers(model)
    assert type(model.fno_blocks).__module__.startswith("neuralop.")                    # the block is still the reference's
    assert all(type(c) is nb.SpectralConv for c in model.fno_blocks.convs)              # ... its convs are ours
    assert all(type(m) is nb.ChannelMLP for m in model.fno_blocks.channel_mlp)          # ... and its (GELU) channel MLPs
    assert type(model.lifting).__module__.startswith("neuralop.")                       # an elu MLP has no drop-in
    assert model.fno_blocks.convs[0].n_modes == [8, 4] and list(model.fno_blocks.convs[0].max_n_modes) == [10, 8]
    xo = x.clone().requires_grad_(True)
    y = model(xo)
    y.backward(gy)
    assert rel_err(y, y_ref.detach()) < 3e-5 and rel_err(xo.grad, dx_ref) < 3e-5
    for k, p in model.named_parameters():
        assert grad_err(p.grad, ref_grads[k], ref_grads) < 5e-5, k       # (the conv bias in front of a norm has a zero gradient)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_use_b200_layers_on_another_model_family_uno(host):  # noqa: F811
    """The swap is by module class, not by model: a reference U-shaped neural operator (neuralop/models/uno.py: FNOBlocks with
    different channel counts / modes / scalings per layer, stand-alone linear skips) ends up without a single reference layer class
    from this package's list, and computes the same function."""
    import importlib
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__file__)), "oracle"))
    from make_golden_fno import load_reference_fno
    load_reference_fno()
    uno = importlib.import_module("neuralop.models.uno")
    torch.manual_seed(0)
    model = uno.UNO(in_channels=1, out_channels=1, hidden_channels=8, n_layers=3, uno_out_channels=[8, 8, 8],
                    uno_n_modes=[[6, 6], [4, 4], [6, 6]], uno_scalings=[[0.5, 0.5], [1, 1], [2, 2]], channel_mlp_skip="linear")
    x = torch.randn(2, 1, 16, 16)
    xr = x.clone().requires_grad_(True)
    y_ref = model(xr)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    dx_ref = xr.grad.clone()
    model.zero_grad(set_to_none=True)
    nb.use_b200_layers(model)
    left = {type(m).__name__ for m in model.modules() if type(m).__module__.startswith("neuralop.")}
    assert not left & {"FNOBlocks", "SpectralConv", "ChannelMLP", "Flattened1dConv", "SoftGating"}, left
    xo = x.clone().requires_grad_(True)
    y = model(xo)
    y.backward(gy)
    assert rel_err(y, y_ref.detach()) < 3e-5 and rel_err(xo.grad, dx_ref) < 3e-5
    for k, p in model.named_parameters():
        assert grad_err(p.grad, ref_grads[k], ref_grads) < 5e-5, k


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_batch_norm_eval_mode_uses_the_running_statistics(host):  # noqa: F811
    import importlib
    from oracle.load_reference import load_reference_spectral_conv
    load_reference_spectral_conv()
    ref_fb = importlib.import_module("neuralop.layers.fno_block")
    torch.manual_seed(8)
    ref = ref_fb.FNOBlocks(4, 4, (6, 6), n_layers=2, norm="batch_norm", implementation="reconstructed")
    ours = nb.FNOBlocks(4, 4, (6, 6), n_layers=2, norm="batch_norm", implementation="reconstructed")
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(3, 4, 12, 12)
    for i in range(2):                                       # two training steps: the running statistics move the same way
        assert rel_err(ours(x, i), ref(x, i).detach()) < 3e-5
    for (na, a), (nb_, b) in zip(ours.named_buffers(), ref.named_buffers()):
        assert na == nb_ and rel_err(a.float(), b.float()) < 1e-5, na
    ours.eval()
    ref.eval()
    x2 = torch.randn(2, 4, 12, 12)
    with torch.no_grad():
        assert rel_err(ours(x2, 0), ref(x2, 0)) < 3e-5     # eval: running statistics instead of the batch's
        assert rel_err(ours(x2, 0), ours.train()(x2, 0)) > 1e-3
    ours.eval()


def test_unsupported_reference_modules_are_reported():
    class ChannelMLP(torch.nn.Module):                     # a reference-like ChannelMLP with an activation the kernels do not have
        in_channels = out_channels = hidden_channels = 4
        n_layers = 2
        non_linearity = staticmethod(torch.nn.functional.elu)
        dropout = None
    holder = torch.nn.Module()
    holder.mlp = ChannelMLP()
    with pytest.warns(UserWarning, match="stays the reference module"):
        nb.use_b200_layers(holder)
    assert type(holder.mlp) is ChannelMLP                  # left in place, and said so
