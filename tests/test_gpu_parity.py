"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden
vectors minted from the unmodified reference.

Tolerance: the contract (BASELINE.json north_star) is 1e-3 relative in fp32, measured as
max|d| / max|ref| per tensor; these tests hold the kernels to REL_TOL = 1e-4, ten times tighter.
The kept-mode index set is compared bit-exactly.
"""
import ctypes
import os

import pytest
import torch

import neuraloperator_b200 as nb
from conftest import forward_kwargs, golden_grads, golden_index, golden_weight, load_golden
from oracle import spectral_conv_oracle as O

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
CASES = sorted(golden_index().keys())


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    a = a.detach().cpu()
    ref = ref.detach().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    scale = ref.abs().max().item()
    return (a - ref).abs().max().item() / max(scale, 1e-20)


def test_extension_is_loaded_and_counts_launches(cuda_device):
    from neuraloperator_b200 import _lib
    lib = _lib.load()
    before = _lib.launch_count()
    plan = nb.get_plan(cuda_device, (16, 12), (16, 12), [8, 4], [8, 4])
    x = torch.randn(2, 3, 16, 12, device=cuda_device)
    nb.analyze(plan, x)
    torch.cuda.synchronize()
    assert _lib.launch_count() > before
    loaded = open(f"/proc/{os.getpid()}/maps").read()
    assert "libspectral_conv_b200.so" in loaded


@pytest.mark.parametrize("grid,modes,maxm", [
    ((16,), (8,), None), ((1024,), (16,), None), ((128, 128), (32, 32), None), ((9, 9), (4, 5), None),
    ((8, 8), (12, 12), None), ((64, 64, 64), (16, 16, 16), None), ((16, 12), (5, 4), (8, 6)),
    ((16, 12), (6, 4), (8, 6)), ((6, 6, 6, 6), (4, 4, 4, 4), None), ((256, 256), (64, 64), None),
])
def test_kept_mode_index_set_bit_exact(cuda_device, grid, modes, maxm):
    stored = O.stored_n_modes(modes)
    maxes = list(maxm) if maxm is not None else stored
    oracle = O.kept_mode_plan(grid, stored, maxes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, maxes)
    assert list(plan.kept) == [p.kept for p in oracle]
    for j, p in enumerate(oracle):
        bins, rows = plan.mode_bins(j)
        assert bins == p.in_bins
        assert rows == p.w_index


@pytest.mark.parametrize("grid,modes,out_grid", [
    ((64,), (16,), None), ((16, 12), (8, 6), None), ((9, 11), (4, 5), None), ((8, 9, 10), (4, 5, 7), None),
    ((12, 12), (10, 8), (24, 24)), ((12, 12), (10, 8), (6, 6)), ((12, 13), (6, 6), (9, 16)), ((6, 6, 6, 6), (4, 4, 4, 4), None),
    ((32, 32), (16, 16), None), ((128, 128), (32, 32), None),
])
def test_transforms_match_closed_form(cuda_device, grid, modes, out_grid):
    """sc_analyze / sc_synthesize (and their adjoints) against the float64 DFT-sum statement."""
    import numpy as np
    torch.manual_seed(3)
    stored = O.stored_n_modes(modes)
    og = list(out_grid) if out_grid is not None else list(grid)
    plans = O.kept_mode_plan(grid, stored)
    plan = nb.get_plan(cuda_device, grid, og, stored, stored)
    B, C = 2, 3
    x = torch.randn(B, C, *grid)
    # forward analysis == rfftn + gather
    xm_ref = torch.fft.rfftn(x.double(), dim=list(range(2, 2 + len(grid))), norm="forward")
    for j, p in enumerate(plans):
        xm_ref = xm_ref.index_select(2 + j, torch.tensor(p.in_bins))
    xm = nb.analyze(plan, x.to(cuda_device))
    assert rel_err(xm, xm_ref.to(torch.cfloat)) < REL_TOL
    # synthesis of random modes == closed form with identity weight (Ci == Co, W = delta)
    ym = torch.randn(B, C, *plan.kept, dtype=torch.cfloat)
    bias = torch.randn(C)
    lead, Sre, Sim = O.synthesis_matrices_f64(plans, og)
    t = ym.numpy().astype(np.complex128)
    for j, S in enumerate(lead):
        t = np.moveaxis(np.tensordot(S, t, axes=([1], [2 + j])), 0, 2 + j)
    y_ref = np.tensordot(t.real, Sre, axes=([t.ndim - 1], [1])) + np.tensordot(t.imag, Sim, axes=([t.ndim - 1], [1]))
    y_ref = y_ref + bias.numpy().reshape(1, C, *[1] * len(grid))
    y = nb.synthesize(plan, ym.to(cuda_device), bias.to(cuda_device))
    assert rel_err(y, torch.from_numpy(y_ref).float()) < REL_TOL
    # adjoint pairs: <S m, g> == <m, S^H g>  and  <A x, m> == <x, A^H m>  (real inner products)
    g = torch.randn(B, C, *og)
    gm = nb.analyze(plan, g.to(cuda_device), adjoint=True).cpu()
    y0 = nb.synthesize(plan, ym.to(cuda_device)).cpu()
    lhs = (y0.double() * g.double()).sum().item()
    rhs = (ym.real.double() * gm.real.double() + ym.imag.double() * gm.imag.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)
    dx = nb.synthesize(plan, ym.to(cuda_device), adjoint=True).cpu()
    lhs = (xm.cpu().real.double() * ym.real.double() + xm.cpu().imag.double() * ym.imag.double()).sum().item()
    rhs = (x.double() * dx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)


def _module_from_golden(meta, arr, device):
    ctor = dict(meta["ctor"])
    if "rank" in ctor and isinstance(ctor["rank"], list) and meta["weight_kind"] == "cp":
        ctor["rank"] = ctor["rank"][0]
    conv = nb.SpectralConv(meta["in_channels"], meta["out_channels"], tuple(meta["n_modes"]), **ctor).to(device)
    w = golden_weight(meta, arr)
    with torch.no_grad():
        for dst, src in zip(conv.weight.decomposition(), w.params()):
            assert dst.shape == src.shape, (dst.shape, src.shape)
            dst.copy_(src.to(device))
        if conv.bias is not None:
            conv.bias.copy_(arr["p__bias"].to(device))
    return conv


@pytest.mark.parametrize("name", CASES)
def test_module_matches_golden(cuda_device, name):
    """Forward + backward through the nn.Module / autograd.Function against the reference's own outputs."""
    meta, arr = load_golden(name)
    conv = _module_from_golden(meta, arr, cuda_device)
    x = arr["x"].to(cuda_device).requires_grad_(True)
    kw = {}
    if "output_shape" in meta["forward"]:
        kw["output_shape"] = tuple(meta["forward"]["output_shape"])
    y = conv(x, **kw)
    assert y.dtype == torch.float32 and not torch.is_complex(y)
    assert list(y.shape[2:]) == meta["out_grid"]
    y.backward(arr["gy"].to(cuda_device))
    assert rel_err(y, arr["y"]) < REL_TOL, "y"
    assert rel_err(x.grad, arr["dx"]) < REL_TOL, "dx"
    gws, gb = golden_grads(meta, arr)
    for i, (p, g) in enumerate(zip(conv.weight.decomposition(), gws)):
        assert p.grad is not None, f"param {i} received no grad"
        assert rel_err(p.grad, g) < REL_TOL, f"dparam{i}"
    if conv.bias is not None:
        assert rel_err(conv.bias.grad, gb) < REL_TOL, "dbias"


def _fwd_bwd_vs_oracle(device, B, Ci, Co, grid, modes, seed=0, tol=REL_TOL, **kw):
    x, w, bias, gy = O.make_inputs(B, Ci, Co, grid, modes, seed=seed)
    y_ref, dx_ref, dws_ref, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes, **kw)
    conv = nb.SpectralConv(Ci, Co, modes).to(device)
    with torch.no_grad():
        conv.weight.tensor.copy_(w.tensor.to(device))
        conv.bias.copy_(bias.to(device))
    xd = x.to(device).requires_grad_(True)
    y = conv(xd)
    y.backward(gy.to(device))
    errs = {"y": rel_err(y, y_ref), "dx": rel_err(xd.grad, dx_ref),
            "dW": rel_err(conv.weight.tensor.grad, dws_ref[0]), "db": rel_err(conv.bias.grad, db_ref)}
    for k, v in errs.items():
        assert v < tol, f"{k}: rel err {v:.3e} (all: {errs})"
    return errs


@pytest.mark.parametrize("B,Ci,Co,grid,modes", [
    (16, 32, 32, (1024,), (16,)),            # BASELINE config 1 (FNO1d Burgers) at full size
    (4, 16, 24, (64, 64), (32, 32)),
    (3, 5, 7, (30, 20), (12, 9)),            # ragged: nothing is a multiple of a tile
    (2, 8, 8, (32, 32, 32), (16, 16, 16)),
    (1, 33, 17, (17, 19), (9, 11)),
])
def test_fwd_bwd_matches_oracle(cuda_device, B, Ci, Co, grid, modes):
    _fwd_bwd_vs_oracle(cuda_device, B, Ci, Co, grid, modes)


def test_headline_config_full_size(cuda_device):
    """BASELINE config 2: (B,C,H,W)=(32,64,128,128), modes (32,32), dense -- direct comparison with the oracle."""
    _fwd_bwd_vs_oracle(cuda_device, 32, 64, 64, (128, 128), (32, 32))


def test_fno3d_config_full_size(cuda_device):
    """BASELINE config 4: (8,32,64^3), modes 16^3."""
    _fwd_bwd_vs_oracle(cuda_device, 8, 32, 32, (64, 64, 64), (16, 16, 16))


def test_highres_config_256(cuda_device):
    """BASELINE config 5 at R=256: (16,64,256,256), modes (64,64)."""
    _fwd_bwd_vs_oracle(cuda_device, 16, 64, 64, (256, 256), (64, 64))


@pytest.mark.parametrize("B,R", [(2, 512), (1, 1024)])
def test_highres_config_512_1024(cuda_device, B, R):
    """BASELINE config 5 at R = 512 and 1024: grid and modes as named, channels 64; the batch is cut (16 -> 2 / 1) so that the
    CPU oracle finishes in seconds.  The last-dim tensor-core kernels + the leading-dim table kernels run here."""
    _fwd_bwd_vs_oracle(cuda_device, B, 64, 64, (R, R), (64, 64))


def test_hidden_channels_128_and_batch_96(cuda_device):
    """Extents above the 64 x 64 tile of the tensor-core contraction: they run the tiled quad kernel, not a fallback."""
    _fwd_bwd_vs_oracle(cuda_device, 4, 128, 128, (64, 64), (16, 16))
    _fwd_bwd_vs_oracle(cuda_device, 96, 16, 24, (64, 64), (16, 16))


def test_tfno_config_full_size(cuda_device):
    """BASELINE config 3: TFNO2d, (B,C,H,W) = (32,64,128,128), modes (32,32), Tucker ranks (36,36,18,10) (= rank 0.1),
    implementation="factorized": y, dx, the core gradient, every factor gradient and dbias against the CPU oracle."""
    dev = cuda_device
    B, C, grid, modes, ranks = 32, 64, (128, 128), (32, 32), [36, 36, 18, 10]
    x, w, bias, gy = O.make_inputs(B, C, C, grid, modes, seed=3, kind="tucker", ranks=ranks)
    y_ref, dx_ref, dws_ref, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes)
    conv = nb.SpectralConv(C, C, modes, factorization="tucker", rank=ranks, implementation="factorized").to(dev)
    with torch.no_grad():
        for dst, src in zip(conv.weight.decomposition(), w.params()):
            assert dst.shape == src.shape, (dst.shape, src.shape)
            dst.copy_(src.to(dev))
        conv.bias.copy_(bias.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = conv(xd)
    y.backward(gy.to(dev))
    assert rel_err(y, y_ref) < REL_TOL, "y"
    assert rel_err(xd.grad, dx_ref) < REL_TOL, "dx"
    assert rel_err(conv.bias.grad, db_ref) < REL_TOL, "dbias"
    for i, (p, g) in enumerate(zip(conv.weight.decomposition(), dws_ref)):
        assert p.grad is not None, f"param {i} received no grad"
        assert rel_err(p.grad, g) < REL_TOL, f"dparam{i}"


def test_properties_at_full_size(cuda_device):
    """Size-independent properties at the headline size: linearity in x, zero response to modes outside
    the kept set, dweight support, bias gradient == sum of the upstream gradient."""
    dev = cuda_device
    torch.manual_seed(11)
    B, C, H, W = 32, 64, 128, 128
    conv = nb.SpectralConv(C, C, (32, 32)).to(dev)
    x1 = torch.randn(B, C, H, W, device=dev)
    x2 = torch.randn(B, C, H, W, device=dev)
    with torch.no_grad():
        y1, y2, y12 = conv(x1), conv(x2), conv(2.0 * x1 - 3.0 * x2)
        b = conv.bias
        lin = 2.0 * (y1 - b) - 3.0 * (y2 - b) + b
        assert rel_err(y12, lin) < REL_TOL
        # a pure tone outside the kept block (ky = 40, kx = 5) and one with kx = 20 > 16 produce bias only
        hh = torch.arange(H, device=dev).view(H, 1).float()
        ww = torch.arange(W, device=dev).view(1, W).float()
        tone = torch.cos(2 * torch.pi * (40 * hh / H + 5 * ww / W)) + torch.sin(2 * torch.pi * (3 * hh / H + 20 * ww / W))
        yt = conv(tone.expand(1, C, H, W).contiguous())
        assert (yt - b).abs().max().item() < 1e-4 * max(conv.weight.tensor.abs().max().item() * C, 1.0)
    xg = x1.clone().requires_grad_(True)
    g = torch.randn(B, C, H, W, device=dev)
    conv(xg).backward(g)
    assert rel_err(conv.bias.grad.reshape(-1), g.sum(dim=(0, 2, 3))) < REL_TOL
    assert conv.weight.tensor.grad.shape == conv.weight.tensor.shape
    assert torch.isfinite(xg.grad).all() and torch.isfinite(conv.weight.tensor.grad).all()


def test_n_modes_can_shrink_at_runtime(cuda_device):
    """Reference test_spectral_convolution.py:68-70: shrinking n_modes keeps the output shape; and the result
    equals the oracle with n_modes < max_n_modes (central weight block)."""
    dev = cuda_device
    x, w, bias, gy = O.make_inputs(2, 3, 3, (12, 12), (10, 8), seed=5)
    conv = nb.SpectralConv(3, 3, (10, 8)).to(dev)
    with torch.no_grad():
        conv.weight.tensor.copy_(w.tensor.to(dev))
        conv.bias.copy_(bias.to(dev))
    y_full = conv(x.to(dev))
    conv.n_modes = (6, 6)
    y_small = conv(x.to(dev))
    assert y_small.shape == y_full.shape
    y_ref = O.spectral_conv_forward(x, w, bias, (6, 6), max_n_modes=[10, 5])
    assert rel_err(y_small, y_ref) < REL_TOL


def test_fno_block_style_usage(cuda_device):
    """The conv is constructed the way FNOBlocks does (fno_block.py:212-237) and trains: every parameter gets a grad."""
    dev = cuda_device
    conv = nb.SpectralConv(8, 8, (12, 12), resolution_scaling_factor=None, max_n_modes=None, rank=1.0,
                           fixed_rank_modes=False, implementation="factorized", separable=False, factorization=None,
                           fno_block_precision="full", decomposition_kwargs={}, complex_data=False,
                           enforce_hermitian_symmetry=True).to(dev)
    x = torch.randn(4, 8, 24, 24, device=dev, requires_grad=True)
    y = conv(x)
    y.square().mean().backward()
    for n, p in conv.named_parameters():
        assert p.grad is not None and torch.isfinite(torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).all(), n


@pytest.mark.parametrize("fact,rank,Ci,Co,grid,modes,max_modes", [
    ("cp", 9, 12, 10, (40, 36), (16, 12), None),
    ("cp", 5, 6, 7, (64,), (20,), None),
    ("cp", 6, 4, 5, (12, 10, 16), (6, 6, 8), None),
    ("cp", 8, 8, 8, (32, 32), (8, 8), (16, 12)),        # kept rows are a slice of the mode factors
    ("tt", 0.3, 12, 10, (40, 36), (16, 12), None),
    ("tt", 0.5, 6, 7, (64,), (20,), None),
    ("tt", 0.4, 4, 5, (12, 10, 16), (6, 6, 8), None),
    ("tt", 0.3, 8, 8, (32, 32), (8, 8), (16, 12)),
])
def test_factor_by_factor_matches_reconstructed(cuda_device, fact, rank, Ci, Co, grid, modes, max_modes):
    """implementation="factorized" (CP / TT contracted factor by factor on the device, reference :55-73 / :106-127) against
    implementation="reconstructed" (weight rebuilt, dense kernels) on the same parameters: outputs and every gradient."""
    dev = cuda_device
    torch.manual_seed(11)
    kw = dict(factorization=fact, rank=rank, max_n_modes=max_modes)
    a = nb.SpectralConv(Ci, Co, modes, implementation="factorized", **kw).to(dev)
    b = nb.SpectralConv(Ci, Co, modes, implementation="reconstructed", **kw).to(dev)
    with torch.no_grad():
        for pa, pb in zip(a.weight.decomposition(), b.weight.decomposition()):
            pa.copy_(torch.randn_like(pa))           # O(1) entries so every factor gradient is well above rounding
            pb.copy_(pa)
        b.bias.copy_(a.bias.normal_())
    B = 3
    x = torch.randn(B, Ci, *grid, device=dev)
    gy = torch.randn(B, Co, *grid, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = a(xa), b(xb)
    ya.backward(gy)
    yb.backward(gy)
    assert rel_err(ya, yb) < REL_TOL, "y"
    assert rel_err(xa.grad, xb.grad) < REL_TOL, "dx"
    assert rel_err(a.bias.grad, b.bias.grad) < REL_TOL, "dbias"
    for i, (pa, pb) in enumerate(zip(a.weight.decomposition(), b.weight.decomposition())):
        assert pa.grad is not None and pb.grad is not None
        assert rel_err(pa.grad, pb.grad) < REL_TOL, f"dparam{i}"


@pytest.mark.parametrize("O_,P,Q,I", [(1, 3, 5, 1), (4, 36, 64, 544), (2, 7, 9, 600), (3, 4, 4, 1500), (5, 1, 2, 33), (32, 5, 6, 7)])
def test_pair_reduce_and_table_contract_primitives(cuda_device, O_, P, Q, I):
    """sc_pair_reduce / sc_table_contract against torch.einsum, including strided (transposed) tables and outputs."""
    from neuraloperator_b200.spectral_conv import _pair_reduce, _table_contract
    dev = cuda_device
    torch.manual_seed(3)
    a = torch.randn(O_, P, I, dtype=torch.complex64, device=dev)
    b = torch.randn(O_, Q, I, dtype=torch.complex64, device=dev)
    ref = torch.einsum("opi,oqi->pq", a.conj().to(torch.complex128), b.to(torch.complex128))
    with torch.cuda.device(dev):
        out = _pair_reduce(a, b, torch.empty(P, Q, dtype=torch.complex64, device=dev), Q, 1, O_, P, Q, I)
        out_t = _pair_reduce(a, b, torch.empty(Q, P, dtype=torch.complex64, device=dev), 1, P, O_, P, Q, I)
    assert rel_err(torch.view_as_real(out), torch.view_as_real(ref.to(torch.complex64))) < REL_TOL
    assert rel_err(torch.view_as_real(out_t.t().contiguous()), torch.view_as_real(ref.to(torch.complex64))) < REL_TOL
    t = torch.randn(P, Q, dtype=torch.complex64, device=dev)
    ref2 = torch.einsum("pq,oqi->opi", t, b)
    with torch.cuda.device(dev):
        got2 = _table_contract(t, Q, 1, False, b, O_, P, Q, I).view(O_, P, I)
        got3 = _table_contract(t.t().contiguous(), 1, P, True, b, O_, P, Q, I).view(O_, P, I)     # table stored transposed
    assert rel_err(torch.view_as_real(got2), torch.view_as_real(ref2)) < REL_TOL
    assert rel_err(torch.view_as_real(got3), torch.view_as_real(torch.einsum("pq,oqi->opi", t.conj(), b))) < REL_TOL


@pytest.mark.parametrize("B,C,grid,modes", [(4, 8, (64, 64), (16, 16)), (2, 6, (30, 20), (12, 9)), (2, 4, (16, 16, 16), (8, 8, 8))])
def test_separable_matches_oracle(cuda_device, B, C, grid, modes):
    """separable=True (depthwise, reference :49-52) on the fused-transform path and on the generic one, against the CPU oracle."""
    torch.manual_seed(21)
    conv = nb.SpectralConv(C, C, modes, separable=True).to(cuda_device)
    with torch.no_grad():
        conv.weight.tensor.copy_(torch.randn_like(conv.weight.tensor))
    w = O.Weight("dense", tensor=conv.weight.tensor.detach().cpu(), separable=True)
    x, gy = torch.randn(B, C, *grid), torch.randn(B, C, *grid)
    y_ref, dx_ref, dws_ref, db_ref = O.spectral_conv_fwd_bwd(x, w, conv.bias.detach().cpu(), gy, modes)
    xd = x.to(cuda_device).requires_grad_(True)
    y = conv(xd)
    y.backward(gy.to(cuda_device))
    assert rel_err(y, y_ref) < REL_TOL and rel_err(xd.grad, dx_ref) < REL_TOL
    assert rel_err(conv.weight.tensor.grad, dws_ref[0]) < REL_TOL and rel_err(conv.bias.grad, db_ref) < REL_TOL


def test_empty_batch(cuda_device):
    """B = 0 is accepted (the reference's torch.fft calls accept it): empty output of the right shape, zero gradients."""
    conv = nb.SpectralConv(4, 6, (8, 8)).to(cuda_device)
    x = torch.empty(0, 4, 16, 16, device=cuda_device, requires_grad=True)
    y = conv(x)
    assert tuple(y.shape) == (0, 6, 16, 16) and y.dtype == torch.float32
    y.sum().backward()
    assert conv.weight.tensor.grad is not None and float(conv.weight.tensor.grad.abs().max()) == 0.0
    assert tuple(x.grad.shape) == (0, 4, 16, 16)


@pytest.mark.parametrize("shape,out", [((2, 3, 16), (24,)), ((2, 3, 12, 10), (18, 20)), ((1, 2, 8, 8, 8), (12, 12, 12)),
                                       ((1, 2, 12, 8, 10), (8, 8, 6)), ((1, 2, 8, 6, 10), (8, 12, 16))])
def test_transform_resamples_like_the_reference(cuda_device, shape, out):
    """`SpectralConv.transform` with a resolution change (the FNO block calls it on every skip path, fno_block.py:377-384):
    reference `resample` (resample.py:7-71; restatement pinned bit-exactly in tests/test_oracle_vs_reference.py), incl. its gradient."""
    torch.manual_seed(4)
    d = len(out)
    conv = nb.SpectralConv(shape[1], shape[1], tuple([4] * d)).to(cuda_device)
    x = torch.randn(*shape)
    ref_in = x.clone().requires_grad_(True)
    ref = O.resample_restated(ref_in, out)
    g = torch.randn_like(ref)
    ref.backward(g)
    xd = x.to(cuda_device).requires_grad_(True)
    got = conv.transform(xd, output_shape=out)
    assert tuple(got.shape) == tuple(ref.shape)
    got.backward(g.to(cuda_device))
    assert rel_err(got, ref) < REL_TOL
    assert rel_err(xd.grad, ref_in.grad) < REL_TOL
    assert conv.transform(xd, output_shape=shape[2:]) is xd       # identity when nothing changes
