"""CPU tests: the oracle against the golden vectors minted from the unmodified reference, the
closed-form float64 statement against the torch.fft statement, and the kept-mode index math against
a literal fftshift+slice emulation (reference spectral_convolution.py:449,500-519)."""
import itertools

import numpy as np
import pytest
import torch

from conftest import forward_kwargs, golden_grads, golden_index, golden_weight, load_golden
from oracle import spectral_conv_oracle as O

CASES = sorted(golden_index().keys())
# CPU FFT libraries may differ between the minting container and the box running the tests
ATOL = 2e-5


def _close(a, b, what):
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= ATOL * max(scale, 1.0), f"{what}: max|d|={err:.3e} scale={scale:.3e}"


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    meta, arr = load_golden(name)
    w = golden_weight(meta, arr)
    bias = arr.get("p__bias")
    y, dx, dws, db = O.spectral_conv_fwd_bwd(arr["x"], w, bias, arr["gy"], meta["n_modes"], **forward_kwargs(meta))
    _close(y, arr["y"], "y")
    _close(dx, arr["dx"], "dx")
    gws, gb = golden_grads(meta, arr)
    assert len(gws) == len(dws)
    for i, (a, b) in enumerate(zip(dws, gws)):
        _close(a, b, f"dparam{i}")
    if bias is not None:
        _close(db, gb, "dbias")


@pytest.mark.parametrize("name", [c for c in CASES if golden_index()[c]["weight_kind"] == "dense"
                                  and not golden_index()[c]["ctor"].get("separable")])
def test_closed_form_f64_matches_golden(name):
    meta, arr = load_golden(name)
    kw = forward_kwargs(meta)
    plans = O.kept_mode_plan(meta["grid"], meta["stored_n_modes"], meta["max_n_modes"])
    og = O.resolve_output_grid(meta["grid"], kw.get("resolution_scaling_factor"), kw.get("output_shape"))
    assert og == meta["out_grid"]
    w = golden_weight(meta, arr).sliced(plans).tensor.numpy().astype(np.complex128)
    bias = arr["p__bias"].numpy().astype(np.float64) if "p__bias" in arr else None
    norm = kw.get("fft_norm", "forward")
    x64 = arr["x"].numpy().astype(np.float64)
    y, _, _ = O.spectral_conv_forward_f64(x64, w, bias, plans, og, norm)
    dx, dW, db = O.spectral_conv_backward_closed_form_f64(x64, w, arr["gy"].numpy().astype(np.float64), plans, og, norm)
    assert np.abs(y - arr["y"].numpy()).max() < 5e-5
    assert np.abs(dx - arr["dx"].numpy()).max() < 5e-5
    gW = arr["g__weight__tensor"]
    kept = gW
    for j, p in enumerate(plans):
        kept = kept.narrow(2 + j, p.w_index[0], p.kept)
    assert np.abs(dW - kept.numpy()).max() < 1e-4
    # gradient support == kept weight block, exactly (the kept-mode index set is bit-exact)
    mask = torch.zeros(gW.shape, dtype=torch.bool)
    sl = mask
    for j, p in enumerate(plans):
        sl = sl.narrow(2 + j, p.w_index[0], p.kept)
    sl.fill_(True)
    assert (gW[~mask] == 0).all()
    if bias is not None:
        assert np.abs(db - arr["g__bias"].numpy().reshape(-1)).max() < 1e-4


def _literal_shift_slice(F, k, last):
    """What `x[slices_x]` reads along one dim, by literally shifting an index array."""
    idx = np.arange(F)
    if last:
        return list(idx[:k]) if k < F else list(idx)
    shifted = np.fft.fftshift(idx)
    c = F // 2
    return list(shifted[c - k // 2: c + k // 2 + k % 2])


@pytest.mark.parametrize("N,n_mode", list(itertools.product([1, 2, 5, 8, 9, 12, 16, 17, 64], [1, 2, 3, 4, 7, 8, 16, 33, 64, 80])))
def test_kept_bins_match_literal_fftshift(N, n_mode):
    # leading dim and last dim of a 2-D problem with the same size/modes
    stored = O.stored_n_modes([n_mode, n_mode])
    plans = O.kept_mode_plan([N, N], stored)
    lead, last = plans
    assert lead.in_bins == _literal_shift_slice(N, min(N, stored[0]), last=False)
    assert last.in_bins == _literal_shift_slice(N // 2 + 1, min(N // 2 + 1, stored[1]), last=True)
    # signed-frequency statement of SURVEY.md App. A.1
    k = lead.kept
    assert lead.freqs == list(range(-(k // 2), k // 2 + k % 2))
    assert all((f % N) == b for f, b in zip(lead.freqs, lead.in_bins))


@pytest.mark.parametrize("maxm,n_mode,N", [(8, 6, 16), (8, 5, 16), (8, 8, 16), (9, 4, 32), (10, 3, 6), (7, 7, 4)])
def test_weight_rows_match_python_slices(maxm, n_mode, N):
    # leading dim: slice(start//2, -start//2) if start else slice(start, None); last: slice(None, -start)
    for last in (False, True):
        F = N // 2 + 1 if last else N
        k = min(F, n_mode)
        start = maxm - k
        rows = list(range(maxm))
        if last:
            expect = rows[slice(None, -start)] if start else rows
        else:
            expect = rows[slice(start // 2, -start // 2)] if start else rows[slice(start, None)]
        grid = [4, N] if last else [N, 4]
        stored = [2, n_mode] if last else [n_mode, 2]
        maxes = [2, maxm] if last else [maxm, 3]
        plan = O.kept_mode_plan(grid, stored, maxes)[1 if last else 0]
        assert plan.w_index == expect


@pytest.mark.parametrize("shape,ranks", [((3, 5, 6, 4), (4, 3, 5, 2)), ((2, 6, 5, 4, 3), (3, 4, 2, 3, 2)), ((2, 4, 9), (3, 2, 4))])
def test_tucker_pairwise_evaluation_equals_the_single_einsum(shape, ranks):
    """Large Tucker problems (BASELINE config 3) are contracted pairwise in the oracle -- the path opt_einsum picks for the
    reference's einsum (:76-103) -- because torch.einsum alone would first form the outer product of x and the core."""
    from oracle import spectral_conv_oracle as O
    torch.manual_seed(0)
    B, Ci, *kept = shape
    Co = 7
    xm = torch.randn(B, Ci, *kept, dtype=torch.cfloat)
    core = torch.randn(*ranks, dtype=torch.cfloat)
    factors = [torch.randn(n, r, dtype=torch.cfloat) for n, r in zip([Ci, Co, *kept], ranks)]
    a = O.contract_tucker(xm, core, factors)
    b = O.contract_tucker_pairwise(xm, core, factors)
    assert (a - b).abs().max().item() < 1e-5 * a.abs().max().item()


from conftest import complex_golden_index, load_complex_golden  # noqa: E402

COMPLEX_CASES = sorted(complex_golden_index().keys())


def _complex_kwargs(meta):
    kw = {"max_n_modes": meta["max_n_modes"], "fft_norm": meta["ctor"].get("fft_norm", "forward"),
          "separable": bool(meta["ctor"].get("separable", False))}
    if "output_shape" in meta["forward"]:
        kw["output_shape"] = meta["forward"]["output_shape"]
    elif "resolution_scaling_factor" in meta["ctor"]:
        kw["output_shape"] = meta["out_grid"]
    return kw


def _complex_dense_weight(meta, arr):
    if meta["weight_kind"].endswith("dense"):
        return arr["p__weight__tensor"]
    n_axes = len(meta["grid"]) + (1 if meta["ctor"].get("separable") else 2)      # separable: one channel axis
    factors = [arr[f"p__weight__factors__{i}"] for i in range(n_axes)]
    return O.tucker_to_dense(arr["p__weight__core"], factors)


@pytest.mark.parametrize("name", COMPLEX_CASES)
def test_complex_oracle_matches_reference_golden(name):
    """complex_data=True: the oracle's gather / scatter restatement against what the unmodified reference returned (y, dx)."""
    meta, arr = load_complex_golden(name)
    x = arr["x"].clone().requires_grad_(True)
    w = _complex_dense_weight(meta, arr)
    y = O.spectral_conv_forward_complex(x, w, arr.get("p__bias"), meta["n_modes"], **_complex_kwargs(meta))
    assert list(y.shape[2:]) == meta["out_grid"]
    assert (y - arr["y"]).abs().max() <= 2e-5 * arr["y"].abs().max()
    y.backward(arr["gy"])
    assert (x.grad - arr["dx"]).abs().max() <= 2e-5 * arr["dx"].abs().max()
