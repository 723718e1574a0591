"""CPU tests of the boundary: the shared object loads, exports every symbol the header declares, and the
host-side module mirrors the reference's constructor / attribute contract. No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib
from neuraloperator_b200.build import LIB_PATH, build_library
from neuraloperator_b200.factorized import FactorizedWeight, tucker_ranks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "spectral_conv_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sc_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} not exported"
    assert b"sm_100a" in lib.sc_build_info()


def test_problem_struct_matches_header():
    # int32 ndim + 4 arrays of SC_MAX_DIMS int32 + int32 fft_norm + int32 flags
    assert ctypes.sizeof(_lib.ScProblem) == 4 * (1 + 4 * _lib.SC_MAX_DIMS + 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_plan_create_fails_loudly_without_gpu(lib):
    prob = _lib.ScProblem()
    prob.ndim = 1
    prob.grid[0] = prob.out_grid[0] = 16
    prob.n_modes[0] = prob.max_n_modes[0] = 5
    handle = ctypes.c_void_p()
    assert lib.sc_plan_create(ctypes.byref(prob), ctypes.byref(handle)) != 0
    assert lib.sc_last_error()


def test_module_contract_dense():
    conv = nb.SpectralConv(4, 6, (8, 6))
    assert conv.n_modes == [8, 4] and conv.max_n_modes == [8, 4] and conv.order == 2
    assert tuple(conv.weight.shape) == (4, 6, 8, 4) and conv.weight.tensor.dtype == torch.complex64
    assert tuple(conv.bias.shape) == (6, 1, 1)
    assert sorted(conv.state_dict().keys()) == ["bias", "weight.tensor"]
    conv.n_modes = (6, 4)           # mutable, halves the last dim (reference :404-415)
    assert conv.n_modes == [6, 3]
    x = torch.randn(2, 4, 16, 16)
    assert conv.transform(x) is x
    with pytest.raises(RuntimeError, match="no CPU path"):
        conv(x)
    cplx = nb.SpectralConv(4, 4, (8, 8), complex_data=True)     # C2C: n_modes not halved (reference :408-414)
    assert cplx.n_modes == [8, 8] and tuple(cplx.weight.shape) == (4, 4, 8, 8)
    csep = nb.SpectralConv(4, 4, (8, 8), complex_data=True, separable=True)
    assert tuple(csep.weight.shape) == (4, 8, 8)
    assert nb.SpectralConv(4, 4, (8, 8), fno_block_precision="half").fno_block_precision == "half"
    assert nb.SpectralConv(4, 4, (8, 8), fno_block_precision="mixed", factorization="tucker", rank=[2, 2, 3, 3]).implementation == "reconstructed"
    assert nb.SpectralConv(4, 4, (8, 8), fno_block_precision="half", factorization="tucker", implementation="factorized").implementation == "factorized"
    with pytest.raises(NotImplementedError):
        nb.SpectralConv(4, 4, (8, 8), fno_block_precision="half", separable=True)
    with pytest.raises(NotImplementedError):
        nb.SpectralConv(4, 4, (8, 8), fno_block_precision="mixed", complex_data=True)
    with pytest.raises(ValueError):
        nb.SpectralConv(4, 4, (8, 8), fno_block_precision="quarter")
    with pytest.raises(ValueError):
        nb.SpectralConv(4, 4, (8, 8), implementation="nope")


def test_module_contract_factorized():
    conv = nb.SpectralConv(6, 5, (8, 6), factorization="tucker", rank=[4, 3, 5, 3], implementation="factorized")
    keys = sorted(conv.state_dict().keys())
    assert keys == ["bias", "weight.core"] + [f"weight.factors.factor_{i}" for i in range(4)]
    assert conv.weight.name.lower().endswith("tucker")
    assert tuple(conv.weight.to_tensor().shape) == (6, 5, 8, 4)
    sub = conv.weight[:, :, 1:5, :3]
    assert tuple(sub.shape) == (6, 5, 4, 3)
    assert torch.allclose(sub.to_tensor(), conv.weight.to_tensor()[:, :, 1:5, :3], atol=1e-6)
    cp = nb.SpectralConv(6, 5, (8, 6), factorization="cp", rank=7)
    assert sorted(cp.state_dict().keys()) == ["bias"] + [f"weight.factors.factor_{i}" for i in range(4)] + ["weight.weights"]
    tt = nb.SpectralConv(6, 5, (8, 6), factorization="tt", rank=[1, 3, 4, 3, 1])
    assert tuple(tt.weight.to_tensor().shape) == (6, 5, 8, 4)
    assert torch.is_complex(conv.weight) and torch.is_complex(cp.weight)


def test_module_contract_separable():
    """separable=True: one channel axis (reference :346-356), equal channel counts enforced with the reference's ValueError."""
    conv = nb.SpectralConv(5, 5, (8, 6), separable=True)
    assert tuple(conv.weight.shape) == (5, 8, 4) and sorted(conv.state_dict().keys()) == ["bias", "weight.tensor"]
    tk = nb.SpectralConv(5, 5, (8, 6), separable=True, factorization="tucker", rank=[3, 5, 3])
    assert tuple(tk.weight.to_tensor().shape) == (5, 8, 4)
    with pytest.raises(ValueError, match="in_channels must be equal"):
        nb.SpectralConv(4, 6, (8, 6), separable=True)


def test_tucker_rank_rule():
    # (64,64,32,17) @ 0.1 -> (36,36,18,10)
    assert tucker_ranks((64, 64, 32, 17), 0.1) == [36, 36, 18, 10]
    w = FactorizedWeight.new((4, 4, 6, 4), rank=0.5, factorization="Tucker")
    assert w.name.lower().endswith("tucker")


def test_factorized_reconstruction_matches_oracle_layout():
    from oracle import spectral_conv_oracle as O
    torch.manual_seed(0)
    conv = nb.SpectralConv(6, 5, (8, 6), factorization="tucker", rank=[4, 3, 5, 3])
    w = O.Weight("tucker", core=conv.weight.core.detach(), factors=[f.detach() for f in conv.weight.factors])
    assert torch.allclose(w.to_dense(), conv.weight.to_tensor().detach(), atol=1e-6)
    xm = torch.randn(2, 6, 8, 4, dtype=torch.cfloat)
    assert torch.allclose(w.contract(xm), O.contract_dense(xm, w.to_dense()), atol=1e-5)


def _c_mode_bins(lib, grid, stored, max_modes, dim):
    prob = _lib.ScProblem()
    prob.ndim = len(grid)
    for j in range(len(grid)):
        prob.grid[j] = grid[j]
        prob.out_grid[j] = grid[j]
        prob.n_modes[j] = stored[j]
        prob.max_n_modes[j] = max_modes[j]
    prob.fft_norm = 0
    cap = max(grid) + 1
    kept = ctypes.c_int32(0)
    bins = (ctypes.c_int32 * cap)()
    rows = (ctypes.c_int32 * cap)()
    rc = lib.sc_problem_mode_bins(ctypes.byref(prob), dim, ctypes.byref(kept), bins, rows)
    assert rc == 0, lib.sc_last_error()
    return list(bins[:kept.value]), list(rows[:kept.value])


def _torch_slicing_bins(N, n_mode_stored, max_mode, last):
    """The reference's own slice objects (spectral_convolution.py:465-519) applied to index vectors with torch's fftshift."""
    F = N // 2 + 1 if last else N
    start = max_mode - min(F, n_mode_stored)                                    # :465-468
    if last:
        sw = slice(None, -start) if start else slice(None)                      # :486
    else:
        sw = slice(start // 2, -start // 2) if start else slice(start, None)    # :483-485
    w_rows = torch.arange(max_mode)[sw]
    kept = len(w_rows)
    spec = torch.arange(F)
    if last:
        bins = spec[slice(None, kept) if kept < F else slice(None)]             # :514-517
    else:
        spec = torch.fft.fftshift(spec)                                         # :449
        centre, neg, pos = F // 2, kept // 2, kept // 2 + kept % 2              # :507-512
        bins = spec[slice(centre - neg, centre + pos)]
    return bins.tolist(), w_rows.tolist()


def test_kept_mode_index_set_is_bit_exact(lib):
    """The library's host index math (sc_problem_mode_bins, no device needed) against the reference's slice objects and the
    oracle, swept over even/odd grids, modes below / at / above the spectrum length and max_n_modes with even / odd starts."""
    from oracle import spectral_conv_oracle as O
    checked = 0
    for N in list(range(1, 14)) + [16, 17, 31, 32, 33, 64]:
        for last in (False, True):
            F = N // 2 + 1 if last else N
            for nm in sorted({1, 2, 3, F - 1, F, F + 1, F + 3, 5, 8} - {0, -1}):
                for extra in (0, 1, 2, 3):
                    if nm < 1:
                        continue
                    mx = nm + extra
                    grid = (5, N) if last else (N, 6)          # put the dim under test in the leading or the last slot
                    dim = 1 if last else 0
                    stored = [3, nm] if last else [nm, 3]
                    maxm = [3, mx] if last else [mx, 3]
                    got = _c_mode_bins(lib, grid, stored, maxm, dim)
                    want = _torch_slicing_bins(N, nm, mx, last)
                    assert got == want, (N, last, nm, mx, got, want)
                    p = O.kept_mode_plan(grid, stored, maxm)[dim]
                    assert got == (p.in_bins, p.w_index)
                    checked += 1
    assert checked > 500


def test_complex_contraction_plans_are_valid_problems(lib):
    """complex_data=True borrows the dense mode GEMM through a real-data plan whose kept block is (k_1..k_d): grid
    (k_1, .., k_{d-1}, 2 k_d), stored modes == max modes == k.  The library's host-side plan builder (no device needed) must
    accept every such problem the complex goldens produce and keep exactly k_j modes with weight rows 0..k_j-1 (whole block)."""
    from conftest import complex_golden_index
    for name, meta in complex_golden_index().items():
        grid = meta["grid"]
        kept = [min(n, k) for n, k in zip(grid, meta["stored_n_modes"])]
        cgrid = [*kept[:-1], 2 * kept[-1]]
        for dim in range(len(kept)):
            bins, rows = _c_mode_bins(lib, cgrid, kept, kept, dim)
            assert len(bins) == kept[dim] and rows == list(range(kept[dim])), (name, dim, bins, rows)
        # the table builder runs the full host-side plan construction
        prob = _lib.ScProblem()
        prob.ndim = len(kept)
        for j in range(len(kept)):
            prob.grid[j] = prob.out_grid[j] = cgrid[j]
            prob.n_modes[j] = prob.max_n_modes[j] = kept[j]
        r, c = ctypes.c_int64(0), ctypes.c_int64(0)
        rc = lib.sc_problem_table(ctypes.byref(prob), 0, 0,   # SC_TABLE_LAST_ANALYSIS
                                  None, 0, ctypes.byref(r), ctypes.byref(c))
        assert rc == 0, (name, lib.sc_last_error())
        assert (r.value, c.value) == (cgrid[-1], 2 * kept[-1]), (name, r.value, c.value)


def test_product_package_never_uses_the_test_hooks_or_the_oracle():
    """The `sc_hostcheck_*` entry points (host execution of the layer kernels' tile functions, dry run of the C chains) and everything
    under oracle/ are test infrastructure: the package binds the symbols (header == binding == exports) but no product module calls
    them, imports the oracle, or reaches for torch.fft / cuFFT / Triton / torch.compile."""
    import glob
    import re
    pkg = os.path.join(ROOT, "neuraloperator_b200")
    for path in glob.glob(os.path.join(pkg, "*.py")):
        src = open(path).read()
        code = "\n".join(line.split("#", 1)[0] for line in src.splitlines())        # comments may mention them
        code = re.sub(r'""".*?"""', "", code, flags=re.S)
        if not path.endswith("_lib.py"):
            assert "sc_hostcheck" not in code, path
        assert not re.search(r"^\s*(from|import)\s+oracle", code, flags=re.M), path
        for banned in ("torch.fft", "cufft", "triton", "torch.compile"):
            assert banned not in code, (path, banned)
