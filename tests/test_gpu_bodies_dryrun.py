"""The bodies of the GPU tiers that were written WITHOUT hardware access, executed on the CPU.

tests/test_gpu_zz_complex.py, tests/test_gpu_zzz_layer.py (layer epilogue, blocks, model stacks, reduced precision) first run on a
GPU at the round-end tier.  So that a mistake in a test itself (reference computation, argument order, tolerance logic, parameter
names) or in the product's host logic cannot be what fails there, this file calls those very test functions with a CPU device and
the device primitives emulated: the layer kernels by their host checks (`sc_hostcheck_*`), the spectral conv by the CPU oracle, the
transforms by the oracle's torch.fft statements and their autograd adjoints, the contractions by einsums.  What remains for the GPU
run is the device code itself."""
import contextlib
import ctypes

import pytest
import torch

from neuraloperator_b200 import _lib, spectral_conv as sc
from oracle import spectral_conv_oracle as O
from test_block_host_logic import host  # noqa: F401  (fixture: layer kernels -> host checks, conv -> oracle)
from test_factorized_host_logic import _LIB as _CHAIN_LIB, _table_contract
from test_reference_suite_cpu import FakePlan

import test_gpu_zz_complex as G_COMPLEX
import test_gpu_zzz_layer as G_LAYER

CPU = torch.device("cpu")


@pytest.fixture
def no_cuda_calls(monkeypatch):
    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


@pytest.mark.parametrize("B,Ci,Co,grid", [(2, 3, 5, (7,)), (2, 17, 65, (129,)), (3, 70, 130, (3, 11)), (2, 32, 16, (6, 6, 6))])
@pytest.mark.parametrize("opts", ["plain", "all"])
def test_body_channel_mix(host, no_cuda_calls, monkeypatch, B, Ci, Co, grid, opts):  # noqa: F811
    monkeypatch.setattr(_lib, "launch_count", lambda: len(host.calls))
    G_LAYER.test_channel_mix_forward_backward_vs_float64(CPU, B, Ci, Co, grid, opts)


def test_body_small_layer_pieces(host, no_cuda_calls):  # noqa: F811
    G_LAYER.test_channel_mix_no_grad_and_empty(CPU)
    G_LAYER.test_tanh_stabilizer_kernel(CPU)


@pytest.mark.parametrize("name", G_LAYER.CASES)
def test_body_block_goldens(host, no_cuda_calls, name):  # noqa: F811
    G_LAYER.test_block_module_matches_reference_golden(CPU, name)


@pytest.mark.parametrize("name", ["fno_d1_small", "fno_d2_small", "tfno_d2_small"])
def test_body_model_stacks(host, no_cuda_calls, name):  # noqa: F811
    G_LAYER.test_stacked_drop_ins_match_reference_fno_golden(CPU, name)


# ---- real-data conv with emulated primitives (transforms = the oracle's torch.fft statements, adjoints through autograd) ----------
def _analyze(plan, x, adjoint=False):
    if not adjoint:
        return O.analyze_modes(x, plan.dims, plan.fft_norm)
    with torch.enable_grad():
        m0 = torch.zeros(x.shape[0], x.shape[1], *plan.kept, dtype=torch.cfloat, requires_grad=True)
        return torch.autograd.grad(O.synthesize_modes(m0, plan.dims, plan.out_grid, plan.fft_norm), m0, x)[0]


def _synthesize(plan, m, bias=None, adjoint=False):
    if not adjoint:
        return O.synthesize_modes(m, plan.dims, plan.out_grid, plan.fft_norm) + (bias.reshape(1, -1, *[1] * plan.ndim) if bias is not None else 0)
    with torch.enable_grad():
        x0 = torch.zeros(m.shape[0], m.shape[1], *plan.grid, requires_grad=True)
        return torch.autograd.grad(O.analyze_modes(x0, plan.dims, plan.fft_norm), x0, m)[0]


def _contract_dense_backward(plan, xm, gm, w, need_dxm=True, need_dweight=True, need_dbias=True):
    dxm = torch.einsum("bo...,io...->bi...", gm, plan.cut(w).conj()).contiguous() if need_dxm else None
    dw = None
    if need_dweight:
        dw = torch.zeros_like(w)
        view = dw
        for j in range(plan.ndim):
            lo, hi = plan.weight_row_range(j)
            view = view.narrow(2 + j, lo, hi - lo)
        view.copy_(torch.einsum("bi...,bo...->io...", xm.conj(), gm))
    db = None
    if need_dbias:      # fft_norm "forward": the DC slot of gm is sum_n gy
        db = gm[(slice(None), slice(None)) + tuple(p.in_bins.index(0) for p in plan.dims)].real.sum(0)
    return dxm, dw, db


@pytest.fixture
def conv_primitives(monkeypatch, no_cuda_calls):
    real = _lib.load()

    class Host:
        def sc_pointwise(self, op, a, b, out, n, st):
            return real.sc_hostcheck_pointwise(op, a, b, out, n)

        def sc_last_error(self):
            return real.sc_last_error()

    monkeypatch.setattr(sc._lib, "load", lambda: Host())
    monkeypatch.setattr(sc, "_stream_ptr", lambda dev: None)
    monkeypatch.setattr(sc, "_ptr", lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None)
    monkeypatch.setattr(sc, "get_plan", lambda dev, grid, out, nm, mx, norm="forward", flags=0: FakePlan(dev, grid, out, nm, mx, norm, flags))
    monkeypatch.setattr(sc, "analyze", _analyze)
    monkeypatch.setattr(sc, "synthesize", _synthesize)
    monkeypatch.setattr(sc, "contract_dense", lambda plan, xm, w: torch.einsum("bi...,io...->bo...", xm, plan.cut(w)).contiguous())
    monkeypatch.setattr(sc, "contract_dense_backward", _contract_dense_backward)


@pytest.mark.parametrize("precision", ["mixed", "half"])
@pytest.mark.parametrize("B,Ci,Co,grid,modes", [(2, 4, 4, (16, 12), (8, 6)), (2, 3, 5, (64,), (16,)), (1, 4, 4, (8, 8, 8), (4, 4, 4))])
def test_body_reduced_precision(conv_primitives, precision, B, Ci, Co, grid, modes):
    G_LAYER.test_reduced_precision_matches_rounding_point_oracle(CPU, precision, B, Ci, Co, grid, modes)


# ---- complex data -------------------------------------------------------------------------------------------------------------
@pytest.fixture
def complex_primitives(monkeypatch, no_cuda_calls):
    monkeypatch.setattr(sc._lib, "load", lambda: _CHAIN_LIB)
    monkeypatch.setattr(sc._lib, "check", lambda rc, what: None)
    monkeypatch.setattr(sc, "_ptr", lambda t: t)
    monkeypatch.setattr(sc, "_stream_ptr", lambda dev: None)
    monkeypatch.setattr(sc, "_table_contract", _table_contract)
    monkeypatch.setattr(sc, "get_plan", lambda dev, grid, out, nm, mx, norm="forward", flags=0: FakePlan(dev, grid, out, nm, mx, norm, flags))
    monkeypatch.setattr(sc, "get_complex_plan", lambda dev, grid, out, nm, mx, norm: sc.ComplexPlan(torch.device("cpu"), grid, out, nm, mx, norm))
    monkeypatch.setattr(sc, "contract_dense", lambda plan, xm, w: torch.einsum("bi...,io...->bo...", xm, w).contiguous())
    monkeypatch.setattr(sc, "contract_dense_backward", lambda plan, xm, gm, w, **kw: (
        torch.einsum("bo...,io...->bi...", gm, w.conj()).contiguous(), torch.einsum("bi...,bo...->io...", xm.conj(), gm).contiguous(), None))


@pytest.mark.parametrize("name", G_COMPLEX.CASES)
def test_body_complex_goldens(complex_primitives, name):
    G_COMPLEX.test_complex_module_matches_golden(CPU, name)


# ---- CP / TT: Python-orchestrated chain vs the one-C-call entry points (tests/test_gpu_zzzz_factorized_c.py) ---------------------
import test_gpu_zzzz_factorized_c as G_CHAINS  # noqa: E402
from test_factorized_host_logic import _pair_reduce  # noqa: E402


class _ChainCalls(type(_CHAIN_LIB)):
    """The per-factor primitives as in test_factorized_host_logic; the one-call entry points with REAL transforms around the oracle's
    einsum statement of the contraction (their C orchestration is replayed from the launch log in tests/test_factorized_chain_log.py)."""

    @staticmethod
    def _run(plan, x, bias, contract):
        m = O.analyze_modes(x, plan.dims, plan.fft_norm)
        y = O.synthesize_modes(contract(m), plan.dims, plan.out_grid, plan.fft_norm)
        return y + (bias.reshape(1, -1, *[1] * plan.ndim) if bias is not None else 0)

    def sc_pointwise(self, *a):
        raise AssertionError("not used here")

    def sc_bias_grad(self, plan, gm, db, B, Co, st):                   # fft_norm "forward": the DC slot of gm is sum_n gy
        db.copy_(gm[(slice(None), slice(None)) + tuple(p.in_bins.index(0) for p in plan.dims)].real.sum(0))
        return 0

    def sc_forward_cp(self, plan, x, lam, u_in, u_out, u_modes, bias, y, saved, B, Ci, Co, R, ws, n, st):
        self._x = x.detach().clone()
        y.copy_(self._run(plan, x, bias, lambda m: O.contract_cp(m, lam, [u_in, u_out, *u_modes])))
        return 0

    def sc_backward_cp(self, plan, gy, lam, u_in, u_out, u_modes, saved, dx, d_lam, d_u_in, d_u_out, d_u_modes, dbias, B, Ci, Co, R, ws, n, st):
        with torch.enable_grad():
            ps = [t.detach().clone().requires_grad_(True) for t in (self._x, lam, u_in, u_out, *u_modes)]
            self._run(plan, ps[0], None, lambda m: O.contract_cp(m, ps[1], ps[2:])).backward(gy)
        for dst, src in zip([dx, d_lam, d_u_in, d_u_out, *d_u_modes], ps):
            dst.copy_(src.grad)
        if dbias is not None:
            dbias.copy_(gy.sum(dim=[0] + list(range(2, gy.ndim))))
        return 0

    def sc_forward_tt(self, plan, plan_kept, x, g0, g1, cores, bias, y, saved, B, Ci, Co, ranks, ws, n, st):
        self._x = x.detach().clone()
        y.copy_(self._run(plan, x, bias, lambda m: O.contract_tt(m, [g0, g1, *cores])))
        return 0

    def sc_backward_tt(self, plan, plan_kept, gy, g0, g1, cores, saved, dx, d_g0, d_g1, d_cores, dbias, B, Ci, Co, ranks, ws, n, st):
        with torch.enable_grad():
            ps = [t.detach().clone().requires_grad_(True) for t in (self._x, g0, g1, *cores)]
            self._run(plan, ps[0], None, lambda m: O.contract_tt(m, ps[1:])).backward(gy)
        for dst, src in zip([dx, d_g0, d_g1, *d_cores], ps):
            dst.copy_(src.grad)
        if dbias is not None:
            dbias.copy_(gy.sum(dim=[0] + list(range(2, gy.ndim))))
        return 0


@pytest.fixture
def chain_primitives(monkeypatch, conv_primitives):
    lib = _ChainCalls()
    monkeypatch.setattr(sc._lib, "load", lambda: lib)
    monkeypatch.setattr(sc._lib, "check", lambda rc, what: None)
    monkeypatch.setattr(sc._lib, "launch_count", lambda: 0)
    monkeypatch.setattr(sc, "_ptr", lambda t: t)
    monkeypatch.setattr(sc, "_ptr_array", lambda ts: list(ts))
    monkeypatch.setattr(sc, "_table_contract", _table_contract)
    monkeypatch.setattr(sc, "_pair_reduce", _pair_reduce)
    monkeypatch.setattr(sc, "_cp_factor_args", lambda us, kept: (list(us), list(kept), len(us)))
    monkeypatch.setattr(G_CHAINS._lib, "launch_count", iter(range(10 ** 6)).__next__)       # "kernels were launched": any increasing counter
    return lib


@pytest.mark.parametrize("fact,rank,Ci,Co,grid,modes,max_modes", [("cp", 6, 4, 5, (12, 10, 16), (6, 6, 8), None), ("cp", 8, 8, 8, (32, 32), (8, 8), (16, 12)),
                                                                 ("tt", 0.5, 6, 7, (64,), (20,), None), ("tt", 0.3, 8, 8, (32, 32), (8, 8), (16, 12))])
def test_body_c_chain_vs_python_chain(chain_primitives, fact, rank, Ci, Co, grid, modes, max_modes):
    G_CHAINS.test_c_chain_matches_python_chain(CPU, fact, rank, Ci, Co, grid, modes, max_modes)


@pytest.mark.parametrize("name", ["d2_cp", "d2_tt"])
def test_body_c_chain_goldens(chain_primitives, name):
    old = sc.FACTORIZED_CHAINS_IN_C
    sc.FACTORIZED_CHAINS_IN_C = True
    try:
        G_CHAINS.test_c_chain_matches_reference_golden(CPU, None, name)
    finally:
        sc.FACTORIZED_CHAINS_IN_C = old


# ---- the tensor-core mixing tier: its test LOGIC only (on the CPU both settings run the SIMT tile functions) ------------------------
import test_gpu_zzzzz_mix_tc as G_TC  # noqa: E402


@pytest.mark.parametrize("B,Ci,Co,grid", [(2, 64, 32, (50, 30)), (3, 5, 7, (129,)), (1, 200, 100, (40,))])
def test_body_tensor_core_tier_logic(host, no_cuda_calls, monkeypatch, B, Ci, Co, grid):  # noqa: F811
    flag = {"on": False}
    monkeypatch.setattr(G_TC.nb, "set_tensor_core_mixing", lambda enable: flag.__setitem__("on", bool(enable)))
    monkeypatch.setattr(G_TC.nb, "uses_tensor_core_mixing", lambda: flag["on"])
    flag["on"] = True                                           # what the `tensor_cores` fixture of that file does
    G_TC.test_tensor_core_mix_matches_float64_and_simt(CPU, None, B, Ci, Co, grid)
    G_TC.test_tensor_core_mix_through_the_block(CPU, None)


def test_body_mixed_precision_block_runs(host, no_cuda_calls):  # noqa: F811
    """(the conv is the full-precision oracle here: this only exercises the test's own logic and the block wiring with tanh)"""
    G_LAYER.test_block_with_mixed_precision_and_tanh_runs(CPU)
