"""The library's twiddle tables, fetched on the host (sc_problem_table: no GPU), applied in numpy and compared with the
outputs of the unmodified reference stored in tests/golden/.

The truncated transforms the kernels run ARE products with these tables (tcgen05 / SIMT kernels only differ in how they
tile the products), so this pins norms, the Hermitian rules of the C2R step (spectral_convolution.py:552-559), resampling
(`output_shape` / `resolution_scaling_factor`, :524-528), odd grids, modes above the spectrum length and `max_n_modes`
on CPU, forward (y) and backward (dx), for every golden case.  The mode-wise contraction in the middle is taken from the
oracle's einsum restatement (differentiated by torch for the backward direction)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_index, golden_weight, load_golden
from neuraloperator_b200 import _lib
from oracle import spectral_conv_oracle as O

CASES = sorted(golden_index().keys())
T_LAST_A, T_LAST_AT, T_LAST_S, T_LAST_ST, T_LEAD_A, T_LEAD_AH, T_LEAD_S, T_LEAD_SH = range(8)


def _problem(meta):
    prob = _lib.ScProblem()
    d = len(meta["grid"])
    prob.ndim = d
    for j in range(d):
        prob.grid[j] = meta["grid"][j]
        prob.out_grid[j] = meta["out_grid"][j]
        prob.n_modes[j] = meta["stored_n_modes"][j]
        prob.max_n_modes[j] = meta["max_n_modes"][j]
    prob.fft_norm = _lib.NORMS[meta["ctor"].get("fft_norm", "forward")]
    return prob


def _table(lib, prob, which, dim=0):
    rows, cols = ctypes.c_int64(0), ctypes.c_int64(0)
    assert lib.sc_problem_table(ctypes.byref(prob), which, dim, None, 0, ctypes.byref(rows), ctypes.byref(cols)) == 0, \
        lib.sc_last_error()
    buf = np.empty(rows.value * cols.value, dtype=np.float32)
    rc = lib.sc_problem_table(ctypes.byref(prob), which, dim, buf.ctypes.data_as(ctypes.c_void_p), buf.size,
                              ctypes.byref(rows), ctypes.byref(cols))
    assert rc == 0, lib.sc_last_error()
    t = buf.reshape(rows.value, cols.value).astype(np.float64)
    if which >= T_LEAD_A:
        t = t[:, 0::2] + 1j * t[:, 1::2]
    return t


def _apply(t, x, axis):
    """out[..., p, ...] = sum_q t[p, q] x[..., q, ...] along `axis`."""
    return np.moveaxis(np.tensordot(t, x, axes=([1], [axis])), 0, axis)


def _interleave(z):
    out = np.empty(z.shape[:-1] + (2 * z.shape[-1],), dtype=np.float64)
    out[..., 0::2], out[..., 1::2] = z.real, z.imag
    return out


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("name", CASES)
def test_tables_reproduce_reference_outputs(lib, name):
    meta, arr = load_golden(name)
    prob = _problem(meta)
    d = prob.ndim
    x = arr["x"].numpy().astype(np.float64)
    gy = arr["gy"].numpy().astype(np.float64)

    # ---- forward analysis: last dim (real table), then the leading dims (complex tables)
    r = x @ _table(lib, prob, T_LAST_A)
    xm = r[..., 0::2] + 1j * r[..., 1::2]
    for j in range(d - 1):
        xm = _apply(_table(lib, prob, T_LEAD_A, j), xm, 2 + j)

    # ---- contraction (oracle einsum; torch records its backward)
    plans = O.kept_mode_plan(meta["grid"], meta["stored_n_modes"], meta["max_n_modes"])
    assert list(xm.shape[2:]) == [p.kept for p in plans]
    w = golden_weight(meta, arr).sliced(plans)
    xm_t = torch.from_numpy(xm).to(torch.complex128).requires_grad_(True)
    w64 = O.Weight(w.kind, tensor=None if w.tensor is None else w.tensor.to(torch.complex128),
                   core=None if w.core is None else w.core.to(torch.complex128),
                   weights=None if w.weights is None else w.weights.to(torch.complex128),
                   factors=None if w.factors is None else [f.to(torch.complex128) for f in w.factors],
                   separable=w.separable)
    ym_t = w64.contract(xm_t)
    ym = ym_t.detach().numpy()

    # ---- forward synthesis
    u = ym
    for j in range(d - 1):
        u = _apply(_table(lib, prob, T_LEAD_S, j), u, 2 + j)
    y = _interleave(u) @ _table(lib, prob, T_LAST_S)
    if "p__bias" in arr:
        y = y + arr["p__bias"].numpy().astype(np.float64)
    y_ref = arr["y"].numpy().astype(np.float64)
    assert list(y.shape) == list(y_ref.shape)
    assert np.abs(y - y_ref).max() <= 2e-5 * np.abs(y_ref).max(), "y"

    # ---- backward: adjoint of the synthesis, conjugate contraction, adjoint of the analysis
    g = gy @ _table(lib, prob, T_LAST_ST)
    gm = g[..., 0::2] + 1j * g[..., 1::2]
    for j in range(d - 1):
        gm = _apply(_table(lib, prob, T_LEAD_SH, j), gm, 2 + j)
    ym_t.backward(torch.from_numpy(gm).to(torch.complex128))
    dxm = xm_t.grad.numpy()
    v = dxm
    for j in range(d - 1):
        v = _apply(_table(lib, prob, T_LEAD_AH, j), v, 2 + j)
    dx = _interleave(v) @ _table(lib, prob, T_LAST_AT)
    dx_ref = arr["dx"].numpy().astype(np.float64)
    assert np.abs(dx - dx_ref).max() <= 2e-5 * np.abs(dx_ref).max(), "dx"


def _bf16_round(a):
    """float32 -> nearest-even bfloat16, returned as float32 (the rounding `cvt.rn.bf16x2.f32` applies in the kernels)."""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_bf16x3_operand_split_meets_the_tolerance(lib):
    """Arithmetic model of the tensor-core path on CPU: x = x_hi + x_lo and T = T1 + T2 in bf16, products x_hi*T1 + x_lo*T1 +
    x_hi*T2 accumulated in fp32.  With the library's own cfg-2 analysis table the result stays within 1e-5 of
    the float64 product -- two orders below the 1e-3 contract; a plain bf16 product would miss it."""
    prob = _lib.ScProblem()
    prob.ndim = 2
    for j, (n, k) in enumerate(((128, 32), (128, 17))):
        prob.grid[j] = prob.out_grid[j] = n
        prob.n_modes[j] = prob.max_n_modes[j] = k
    prob.fft_norm = 0
    t64 = _table(lib, prob, T_LAST_A)                       # [128 x 34], forward scale folded in
    t = t64.astype(np.float32)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((512, 128)).astype(np.float32)
    exact = x.astype(np.float64) @ t64
    x_hi = _bf16_round(x)
    x_lo = _bf16_round(x - x_hi)
    t1 = _bf16_round(t)
    t2 = _bf16_round(t - t1)
    acc = (x_hi @ t1).astype(np.float32) + (x_lo @ t1).astype(np.float32) + (x_hi @ t2).astype(np.float32)
    scale = np.abs(exact).max()
    assert np.abs(acc - exact).max() / scale < 1e-5
    assert np.abs((x_hi @ t1) - exact).max() / scale > 1e-3    # single bf16 product: not good enough, hence the split


@pytest.mark.parametrize("grid,out", [((8, 8, 8), (12, 12, 12)), ((12, 8, 10), (8, 8, 6)), ((8, 6, 10), (8, 12, 16)), ((8, 8, 8, 8), (4, 8, 12, 6))])
def test_resample_tables_reproduce_the_reference_resample(lib, grid, out):
    """SpectralConv.transform for 3-D+ = analysis and synthesis of a SC_FLAG_RESAMPLE plan with nothing in between.  Applied in
    numpy, the plan's host tables must reproduce the reference's spectral `resample` (resample.py:52-69, restated in the oracle
    and pinned bit-exactly to the live function), forward and adjoint (the backward of transform)."""
    d = len(grid)
    stored = [min(n, m) for n, m in zip(grid[:-1], out[:-1])] + [min(grid[-1] // 2 + 1, out[-1] // 2 + 1)]
    assert all(k % 2 == 0 for k in stored[:-1])
    prob = _lib.ScProblem()
    prob.ndim = d
    for j in range(d):
        prob.grid[j], prob.out_grid[j] = grid[j], out[j]
        prob.n_modes[j] = prob.max_n_modes[j] = stored[j]
    prob.fft_norm = 0
    prob.flags = _lib.FLAG_RESAMPLE
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, *grid))
    r = x @ _table(lib, prob, T_LAST_A)
    xm = r[..., 0::2] + 1j * r[..., 1::2]
    for j in range(d - 1):
        xm = _apply(_table(lib, prob, T_LEAD_A, j), xm, 2 + j)
    u = xm
    for j in range(d - 1):
        u = _apply(_table(lib, prob, T_LEAD_S, j), u, 2 + j)
    y = _interleave(u) @ _table(lib, prob, T_LAST_S)
    xt = torch.from_numpy(x).float().requires_grad_(True)
    ref = O.resample_restated(xt, out)
    assert list(y.shape) == list(ref.shape)
    assert np.abs(y - ref.detach().numpy()).max() <= 2e-5 * np.abs(ref.detach().numpy()).max()
    # adjoint: what transform's backward applies to the upstream gradient
    gy = rng.standard_normal(y.shape)
    ref.backward(torch.from_numpy(gy).float())
    g = gy @ _table(lib, prob, T_LAST_ST)
    gm = g[..., 0::2] + 1j * g[..., 1::2]
    for j in range(d - 1):
        gm = _apply(_table(lib, prob, T_LEAD_SH, j), gm, 2 + j)
    v = gm
    for j in range(d - 1):
        v = _apply(_table(lib, prob, T_LEAD_AH, j), v, 2 + j)
    dx = _interleave(v) @ _table(lib, prob, T_LAST_AT)
    dx_ref = xt.grad.numpy()
    assert np.abs(dx - dx_ref).max() <= 2e-5 * np.abs(dx_ref).max()
