"""GPU tests of the tcgen05 fast path: primitive bring-up check, then fast-vs-generic agreement."""
import ctypes

import pytest
import torch

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(48, 64), (48, 256), (96, 128)])
def test_umma_selftest_tmem_a_operand(cuda_device, N, K):
    """A operand written to tensor memory with tcgen05.st and consumed by a TMEM-A tcgen05.mma."""
    lib = _lib.load()
    torch.manual_seed(N * 1000 + K + 1)
    a = torch.randn(128, K, device=cuda_device)
    b = torch.randn(N, K, device=cuda_device)
    d = torch.full((128, N), float("nan"), device=cuda_device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    _lib.check(lib.sc_selftest_umma_ts(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                       ctypes.c_void_p(d.data_ptr()), N, K, stream), "sc_selftest_umma_ts")
    torch.cuda.synchronize()
    ref = a.bfloat16().double() @ b.bfloat16().double().T
    err = (d.double() - ref).abs().max().item()
    assert err < 1e-3 * max(ref.abs().max().item(), 1.0), f"max err {err}"


@pytest.mark.parametrize("N,K", [(16, 64), (48, 64), (96, 128), (128, 128), (48, 256)])
def test_umma_selftest(cuda_device, N, K):
    """tcgen05.mma with hand-swizzled bf16 operands, FP32 accumulation in TMEM, tcgen05.ld read-back."""
    lib = _lib.load()
    torch.manual_seed(N * 1000 + K)
    a = torch.randn(128, K, device=cuda_device)
    b = torch.randn(N, K, device=cuda_device)
    d = torch.full((128, N), float("nan"), device=cuda_device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    _lib.check(lib.sc_selftest_umma(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                    ctypes.c_void_p(d.data_ptr()), N, K, stream), "sc_selftest_umma")
    torch.cuda.synchronize()
    ref = a.bfloat16().double() @ b.bfloat16().double().T
    err = (d.double() - ref).abs().max().item()
    assert err < 1e-3 * max(ref.abs().max().item(), 1.0), f"max err {err}"


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


@pytest.mark.parametrize("grid,modes,n0,n1", [
    ((128, 128), (32, 32), 2, 5),       # headline tile shape: one image per 128-row tile
    ((128, 128), (32, 32), 32, 64),     # full headline batch (2048 tiles over 148 CTAs)
    ((128, 64), (16, 16), 3, 4),        # single slab
    ((64, 64), (16, 16), 2, 6),         # two images per tile
    ((128, 128), (12, 20), 1, 3),       # fewer / odd mode counts inside the same tile shape
    ((32, 128), (8, 32), 2, 4),         # four images per tile
])
@pytest.mark.parametrize("adjoint", [False, True])
def test_fast_analysis_matches_generic_and_fp64(cuda_device, grid, modes, n0, n1, adjoint):
    """Fused tcgen05 analysis vs the generic SIMT chain (same tables) and vs a float64 DFT."""
    from oracle import spectral_conv_oracle as O
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, stored)
    mask = plan.uses_fast_path()
    assert mask & (4 if adjoint else 1), f"fast analysis not selected for {grid} (mask {mask})"
    torch.manual_seed(5)
    x = torch.randn(n0, n1, *grid, device=cuda_device)
    fast = nb.analyze(plan, x, adjoint=adjoint)
    plan.set_fast_path(False)
    try:
        slow = nb.analyze(plan, x, adjoint=adjoint)
    finally:
        plan.set_fast_path(True)
    torch.cuda.synchronize()
    assert torch.isfinite(torch.view_as_real(fast)).all()
    assert _rel(torch.view_as_real(fast), torch.view_as_real(slow)) < 1e-4
    if not adjoint and n0 * n1 <= 64:
        plans = O.kept_mode_plan(grid, stored)
        ref = torch.fft.rfftn(x.double().cpu(), dim=(2, 3), norm="forward")
        for j, p in enumerate(plans):
            ref = ref.index_select(2 + j, torch.tensor(p.in_bins))
        assert _rel(torch.view_as_real(fast.cpu().to(torch.complex128)), torch.view_as_real(ref)) < 5e-5


@pytest.mark.parametrize("grid,modes,n0,n1", [
    ((128, 128), (32, 32), 2, 5),
    ((128, 128), (32, 32), 32, 64),
    ((128, 64), (16, 16), 3, 4),
    ((64, 64), (16, 16), 2, 6),
    ((128, 128), (12, 20), 1, 3),
    ((32, 128), (8, 32), 2, 4),
])
@pytest.mark.parametrize("adjoint", [False, True])
def test_fast_synthesis_matches_generic(cuda_device, grid, modes, n0, n1, adjoint):
    from oracle import spectral_conv_oracle as O
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, stored)
    mask = plan.uses_fast_path()
    assert mask & (8 if adjoint else 2), f"fast synthesis not selected for {grid} (mask {mask})"
    torch.manual_seed(9)
    ym = torch.randn(n0, n1, *plan.kept, dtype=torch.cfloat, device=cuda_device)
    bias = None if adjoint else torch.randn(n1, device=cuda_device)
    fast = nb.synthesize(plan, ym, bias, adjoint=adjoint)
    plan.set_fast_path(False)
    try:
        slow = nb.synthesize(plan, ym, bias, adjoint=adjoint)
    finally:
        plan.set_fast_path(True)
    torch.cuda.synchronize()
    assert torch.isfinite(fast).all()
    assert _rel(fast, slow) < 1e-4


@pytest.mark.parametrize("grid,modes,n0,n1", [((64, 64, 64), (16, 16, 16), 2, 4), ((8, 128, 128), (4, 32, 32), 1, 3), ((10, 64, 64), (6, 12, 14), 2, 2)])
def test_fast_3d_matches_generic(cuda_device, grid, modes, n0, n1):
    """3-D: fused tcgen05 kernels on the last two dims + the generic complex table kernel on dim 0."""
    from oracle import spectral_conv_oracle as O
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, stored)
    assert plan.uses_fast_path() & 15 == 15, plan.uses_fast_path()
    torch.manual_seed(21)
    x = torch.randn(n0, n1, *grid, device=cuda_device)
    ym = torch.randn(n0, n1, *plan.kept, dtype=torch.cfloat, device=cuda_device)
    bias = torch.randn(n1, device=cuda_device)
    fast = [nb.analyze(plan, x), nb.analyze(plan, x, adjoint=True), nb.synthesize(plan, ym, bias), nb.synthesize(plan, ym, adjoint=True)]
    plan.set_fast_path(False)
    try:
        slow = [nb.analyze(plan, x), nb.analyze(plan, x, adjoint=True), nb.synthesize(plan, ym, bias), nb.synthesize(plan, ym, adjoint=True)]
    finally:
        plan.set_fast_path(True)
    for f, s_ in zip(fast, slow):
        f = torch.view_as_real(f) if f.is_complex() else f
        s_ = torch.view_as_real(s_) if s_.is_complex() else s_
        assert _rel(f, s_) < 1e-4
