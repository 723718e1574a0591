"""The Fourier-layer epilogue (tanh stabilizer) on the GPU: the fused kernels of csrc/sc_layer.cu against
float64 CPU restatements of the reference ops, and `neuraloperator_b200.FNOBlocks` (CUDA spectral conv + fused epilogue) against
golden vectors minted from the unmodified reference FNOBlocks (oracle/make_golden_block.py): y, dx, every parameter gradient.
(Named zzz so that it runs after every tier that was validated on hardware: this file was written after the round's GPU minutes were
spent; the kernels' tile code is checked on CPU by tests/test_layer_cpu.py, the module logic by tests/test_block_host_logic.py.)"""
import pytest
import torch
import torch.nn.functional as F

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib
from conftest import block_ctor_kwargs, block_golden_index, load_block_golden

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]
REL_TOL = 1e-4
CASES = sorted(block_golden_index().keys())


def rel_err(a, ref):
    wide = lambda t: t.detach().cpu().to(torch.complex128 if t.is_complex() else torch.float64)       # noqa: E731
    a, ref = wide(a), wide(ref)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return (a - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)


def grad_err(got, ref, all_refs):
    """rel_err of one parameter gradient; a reference gradient that vanishes identically (the conv bias in front of an instance /
    group norm) is rounding noise on both sides: then ours has to be as negligible, on the scale of the case's largest gradient."""
    scale = max(float(g.abs().max()) for g in all_refs.values())
    if float(ref.abs().max()) < 1e-5 * scale:
        assert got.shape == ref.shape
        return float(got.detach().abs().max()) / scale
    return rel_err(got, ref)


@pytest.mark.parametrize("B,Ci,Co,grid", [(2, 3, 5, (7,)), (2, 17, 65, (129,)), (3, 70, 130, (3, 11)), (4, 64, 64, (128, 128)),
                                           (2, 32, 16, (16, 16, 16)), (1, 64, 32, (100, 100))])
@pytest.mark.parametrize("opts", ["plain", "all"])
def test_channel_mix_forward_backward_vs_float64(cuda_device, B, Ci, Co, grid, opts):
    g = torch.Generator().manual_seed(Ci * 31 + Co)
    dd = dict(dtype=torch.float64)
    x = torch.randn(B, Ci, *grid, generator=g, **dd).requires_grad_(True)
    w = (torch.randn(Co, Ci, 1, generator=g, **dd) / Ci ** 0.5).requires_grad_(True)
    full = opts == "all"
    bias = torch.randn(Co, generator=g, **dd).requires_grad_(True) if full else None
    add = torch.randn(B, Co, *grid, generator=g, **dd).requires_grad_(True) if full else None
    gate = torch.randn(1, Co, *[1] * len(grid), generator=g, **dd).requires_grad_(True) if full else None
    gated = torch.randn(B, Co, *grid, generator=g, **dd).requires_grad_(True) if full else None
    gout = torch.randn(B, Co, *grid, generator=g, **dd)
    pre = F.conv1d(x.flatten(2), w, bias).view(B, Co, *grid)
    if full:
        pre = pre + add + gate * gated
    ref = F.gelu(pre) if full else pre
    ref.backward(gout)
    leaves = [t for t in (x, w, bias, add, gate, gated) if t is not None]
    dev = [t.detach().float().to(cuda_device).requires_grad_(True) if t is not None else None for t in (x, w, bias, add, gate, gated)]
    before = _lib.launch_count()
    out = nb.channel_mix(*dev, act=_lib.ACT_GELU if full else _lib.ACT_IDENTITY)
    out.backward(gout.float().to(cuda_device))
    torch.cuda.synchronize()
    assert _lib.launch_count() > before
    assert rel_err(out, ref) < REL_TOL, "out"
    for name, t_dev, t_ref in zip(("x", "w", "bias", "add", "gate", "gated"), dev, (x, w, bias, add, gate, gated)):
        if t_ref is not None:
            assert t_dev.grad is not None, name
            assert rel_err(t_dev.grad, t_ref.grad) < REL_TOL, name
    assert len(leaves) >= 2


def test_channel_mix_no_grad_and_empty(cuda_device):
    x = torch.randn(2, 4, 9, device=cuda_device)
    w = torch.randn(6, 4, 1, device=cuda_device)
    with torch.no_grad():
        y = nb.channel_mix(x, w, act=_lib.ACT_GELU)
    assert rel_err(y, F.gelu(F.conv1d(x.cpu().double(), w.cpu().double()))) < REL_TOL
    e = nb.channel_mix(torch.empty(0, 4, 9, device=cuda_device), w)
    assert tuple(e.shape) == (0, 6, 9)


def test_tanh_stabilizer_kernel(cuda_device):
    x = (3 * torch.randn(3, 5, 33, device=cuda_device)).requires_grad_(True)
    from neuraloperator_b200.fno_block import _Tanh
    y = _Tanh.apply(x)
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().cpu().double().requires_grad_(True)
    yr = torch.tanh(xr)
    yr.backward(g.cpu().double())
    assert rel_err(y, yr) < 1e-6 and rel_err(x.grad, xr.grad) < 1e-5


def _our_name(pname):
    return pname.replace("weight.factors.", "weight.factors.factor_")


@pytest.mark.parametrize("name", CASES)
def test_block_module_matches_reference_golden(cuda_device, name):
    meta, io, params, grads = load_block_golden(name)
    ctor = block_ctor_kwargs(meta)
    blk = nb.FNOBlocks(meta["in_channels"], meta["out_channels"], tuple(meta["n_modes"]), n_layers=meta["n_layers"], **ctor).to(cuda_device)
    ours = dict(blk.named_parameters())
    with torch.no_grad():
        for pname, val in params.items():
            assert ours[_our_name(pname)].shape == val.shape, pname
            ours[_our_name(pname)].copy_(val.to(cuda_device))
    if "ada_in_embedding" in io:
        blk.set_ada_in_embeddings(io["ada_in_embedding"].to(cuda_device))
    x = io["x"].to(cuda_device).requires_grad_(True)
    kw = {k: tuple(v) for k, v in meta["forward"].items()}
    y = blk(x, meta["index"], **kw)
    assert y.dtype == (torch.complex64 if meta["ctor"].get("complex_data") else torch.float32) and list(y.shape[2:]) == meta["out_grid"]
    y.backward(io["gy"].to(cuda_device))
    torch.cuda.synchronize()
    for bname, buf in blk.named_buffers():                       # batch norm: running statistics after this (training-mode) forward
        assert rel_err(buf.float(), io["b__" + bname.replace(".", "__")].float()) < REL_TOL, bname
    assert rel_err(y, io["y"]) < REL_TOL, "y"
    assert rel_err(x.grad, io["dx"]) < REL_TOL, "dx"
    for pname in meta["touched"]:
        p = ours[_our_name(pname)]
        assert p.grad is not None, pname
        assert grad_err(p.grad, grads[pname], grads) < REL_TOL, pname


def test_headline_shape_layer_against_oracle(cuda_device):
    """One full Fourier layer at the tile shapes of the headline config (channels 64, 128 x 128, modes 32 x 32; batch 4): the fused
    tcgen05 conv kernels + the fused epilogue, against the CPU oracle of the reference block."""
    from oracle import fno_block_oracle as BO
    torch.manual_seed(12)
    blk = nb.FNOBlocks(64, 64, (32, 32), n_layers=2, implementation="reconstructed").to(cuda_device)
    with torch.no_grad():
        for pname, p in blk.named_parameters():
            if "skips" in pname:
                p.add_(0.2 * torch.randn_like(p))
    x = torch.randn(4, 64, 128, 128, device=cuda_device, requires_grad=True)
    y = blk(x, 0)
    gy = torch.randn_like(y)
    y.backward(gy)
    torch.cuda.synchronize()
    params = {k: v.detach().cpu() for k, v in blk.named_parameters()}
    y_ref, dx_ref, g_ref = BO.fno_block_fwd_bwd(x.detach().cpu(), params, 0, gy.cpu(), n_modes=(32, 32), n_layers=2)
    assert rel_err(y, y_ref) < REL_TOL and rel_err(x.grad, dx_ref) < REL_TOL
    for pname, p in blk.named_parameters():
        if pname in g_ref:
            assert rel_err(p.grad, g_ref[pname]) < REL_TOL, pname


# ---- f3: fno_block_precision "half" / "mixed" ----------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["mixed", "half"])
@pytest.mark.parametrize("B,Ci,Co,grid,modes", [(2, 4, 4, (16, 12), (8, 6)), (2, 3, 5, (64,), (16,)), (2, 8, 8, (128, 128), (32, 32)),
                                                (1, 4, 4, (8, 8, 8), (4, 4, 4))])
def test_reduced_precision_matches_rounding_point_oracle(cuda_device, precision, B, Ci, Co, grid, modes):
    """The module with fp16 rounding points against the oracle's statement of the same pipeline (oracle.spectral_conv_forward_reduced;
    its contraction stage is pinned to the reference's einsum_complexhalf on CPU).  Tolerance: fp16 rounding noise (a mode that lands
    within 1e-6 of a rounding boundary may flip by one fp16 ulp between the two transforms)."""
    from oracle import spectral_conv_oracle as O
    x, w, bias, gy = O.make_inputs(B, Ci, Co, grid, modes, seed=4)
    xr = x.clone().requires_grad_(True)
    wt = w.tensor.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    y_ref = O.spectral_conv_forward_reduced(xr, O.Weight("dense", tensor=wt), br, modes, precision)
    y_ref.backward(gy)
    y_full = O.spectral_conv_forward(x, w, bias, modes)
    conv = nb.SpectralConv(Ci, Co, modes, fno_block_precision=precision).to(cuda_device)
    with torch.no_grad():
        conv.weight.tensor.copy_(w.tensor.to(cuda_device))
        conv.bias.copy_(bias.to(cuda_device))
    xd = x.to(cuda_device).requires_grad_(True)
    y = conv(xd)
    assert y.dtype == torch.float32
    y.backward(gy.to(cuda_device))
    torch.cuda.synchronize()
    tol = 2e-3
    assert rel_err(y, y_ref) < 4e-3, "y"        # a kept mode within ~1e-5 of an fp16 rounding boundary lands one ulp (4.9e-4) apart
    assert rel_err(xd.grad, xr.grad) < tol, "dx"
    assert rel_err(conv.weight.tensor.grad, wt.grad) < tol, "dW"
    assert rel_err(conv.bias.grad, br.grad) < tol, "db"
    assert 1e-7 < rel_err(y, y_full) < 5e-3                      # reduced precision really happened, and stays fp16-close to full
    assert torch.equal(xd.detach().cpu(), x)                     # "half" rounds a copy, never the caller's tensor


def test_block_with_mixed_precision_and_tanh_runs(cuda_device):
    """The configuration the reference recommends for reduced precision (stabilizer="tanh", fno_block.py:95-99) through FNOBlocks."""
    torch.manual_seed(5)
    blk = nb.FNOBlocks(8, 8, (8, 8), n_layers=2, fno_block_precision="mixed", stabilizer="tanh").to(cuda_device)
    full = nb.FNOBlocks(8, 8, (8, 8), n_layers=2, stabilizer="tanh").to(cuda_device)
    full.load_state_dict(blk.state_dict())
    x = torch.randn(2, 8, 16, 16, device=cuda_device, requires_grad=True)
    y = blk(x, 0)
    y.sum().backward()
    with torch.no_grad():
        y_full = full(x, 0)
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
    assert rel_err(y, y_full) < 5e-3


# ---- model level: the drop-ins stacked as the reference FNO stacks them ----------------------------------------------------------
@pytest.mark.parametrize("name", ["fno_d1_small", "fno_d2_small", "tfno_d2_small"])
def test_stacked_drop_ins_match_reference_fno_golden(cuda_device, name):
    """lifting (ChannelMLP) -> n_layers x FNOBlocks -> projection (ChannelMLP), every piece from this package, against y, dx and every
    parameter gradient of the unmodified reference `neuralop.models.FNO` / TFNO (oracle/make_golden_fno.py)."""
    from conftest import build_fno_stack, load_fno_golden
    meta, io, params, grads = load_fno_golden(name)
    mods, forward = build_fno_stack(meta, params, device=cuda_device)
    x = io["x"].to(cuda_device).requires_grad_(True)
    y = forward(x)
    y.backward(io["gy"].to(cuda_device))
    torch.cuda.synchronize()
    assert rel_err(y, io["y"]) < REL_TOL, "y"
    assert rel_err(x.grad, io["dx"]) < REL_TOL, "dx"
    ours = dict(mods.named_parameters())
    for k, g in grads.items():
        assert rel_err(ours[k.replace("weight.factors.", "weight.factors.factor_")].grad, g) < 2 * REL_TOL, k
