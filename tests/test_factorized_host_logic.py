"""Host-side logic of the factor-by-factor contractions (Tucker / CP / TT autograd functions) without a GPU.

The device primitives behind the C ABI (`sc_table_contract`, `sc_pair_reduce`, `sc_cp_*`, the transforms and the dense mode
GEMM) are replaced by small torch restatements of their documented contracts (include/spectral_conv_b200.h), so what is
checked here is everything the Python side decides: operand strides, factor slicing, the order of the chains and which
gradient goes where.  Reference for the numbers: the oracle's einsum restatements of `_contract_tucker/_cp/_tt`
(neuralop/layers/spectral_convolution.py:55-127) differentiated by PyTorch.  The kernels themselves are checked on the GPU
(tests/test_gpu_parity.py)."""
import contextlib
import math

import pytest
import torch

from neuraloperator_b200 import spectral_conv as sc
from oracle import spectral_conv_oracle as O


def _table(table, s_p, s_q, conj, P, Q):
    idx = torch.arange(P)[:, None] * s_p + torch.arange(Q)[None, :] * s_q
    T = table.reshape(-1)[idx]
    return T.conj() if conj else T


def _table_contract(table, s_p, s_q, conj, src, n_outer, P, Q, n_inner):
    return torch.einsum("pq,oqi->opi", _table(table, s_p, s_q, conj, P, Q), src.reshape(n_outer, Q, n_inner)).reshape(-1)


def _pair_reduce(a, b, out, s_p, s_q, n_outer, P, Q, n_inner):
    r = torch.einsum("opi,oqi->pq", a.reshape(n_outer, P, n_inner).conj(), b.reshape(n_outer, Q, n_inner))
    idx = torch.arange(P)[:, None] * s_p + torch.arange(Q)[None, :] * s_q
    out.view(-1)[idx.reshape(-1)] = r.reshape(-1)
    return out


def _cp_scale(lam, us, ks):
    d, R = len(us), lam.shape[0]
    s = lam.reshape(R, *[1] * d)
    for j, u in enumerate(us):
        shp = [R] + [1] * d
        shp[1 + j] = ks[j]
        s = s * u.t().reshape(shp)
    return s.reshape(-1)


class _Lib:
    def sc_bias_grad(self, plan, gm, db, B, Co, st):
        db.copy_(gm.reshape(B, Co, -1).real.sum(dim=(0, 2)))
        return 0

    def sc_cp_scale(self, us, ks, d, lam, scale, R, st):
        scale.copy_(_cp_scale(lam, us, ks))
        return 0

    def sc_cp_apply(self, a, scale, out, conj, B, per, st):
        sc_ = scale.reshape(-1)
        out.copy_((a.reshape(B, per) * (sc_.conj() if conj else sc_)[None]).reshape(out.shape))
        return 0

    def sc_cp_dscale(self, t, g, ds, B, per, st):
        ds.copy_((t.reshape(B, per).conj() * g.reshape(B, per)).sum(0).reshape(ds.shape))
        return 0

    def sc_cp_factor_grad(self, us, ks, d, lam, ds, out, which, R, st):
        with torch.enable_grad():
            lam_ = lam.detach().clone().requires_grad_(True)
            us_ = [u.detach().clone().requires_grad_(True) for u in us]
            _cp_scale(lam_, us_, ks).backward(ds)
        out.copy_((lam_.grad if which < 0 else us_[which].grad).reshape(out.shape))
        return 0


    # the Tucker chain is one C call per direction: emulate the documented contract of the two entry points (the chain itself is
    # checked on the GPU); what is checked here is the Python plumbing -- argument order, the saved buffer, which gradient goes where
    def sc_tucker_workspace_bytes(self, plan, B, Ci, Co, ranks):
        return 16

    def sc_tucker_saved_elems(self, plan, B, Ci, Co, ranks):
        return 0

    def sc_forward_tucker(self, plan, plan_kept, x, core, u_in, u_out, u_modes, bias, y, saved, B, Ci, Co, ranks, ws, n, st):
        self._tucker_x = x.detach().clone()
        out = O.contract_tucker(x.to(torch.complex64), core, [u_in, u_out, *u_modes]).real
        y.copy_(out + (bias.reshape(1, -1, *[1] * (out.ndim - 2)) if bias is not None else 0))
        return 0

    def sc_backward_tucker(self, plan, plan_kept, gy, core, u_in, u_out, u_modes, saved, dx, d_core, d_u_in, d_u_out, d_u_modes, dbias,
                           B, Ci, Co, ranks, ws, n, st):
        with torch.enable_grad():
            ps = [t.detach().clone().requires_grad_(True) for t in (self._tucker_x, core, u_in, u_out, *u_modes)]
            O.contract_tucker(ps[0].to(torch.complex64), ps[1], ps[2:]).real.backward(gy)
        for dst, src in zip([dx, d_core, d_u_in, d_u_out, *d_u_modes], ps):
            dst.copy_(src.grad)
        if dbias is not None:
            dbias.copy_(gy.sum(dim=[0] + list(range(2, gy.ndim))))
        return 0


    # CP / TT behind one C call per direction (sc_forward_cp / sc_backward_cp, sc_forward_tt / sc_backward_tt): same treatment
    def sc_cp_saved_elems(self, plan, B, Ci, Co, R):
        return 0

    def sc_cp_workspace_bytes(self, plan, B, Ci, Co, R):
        return 16

    def sc_forward_cp(self, plan, x, lam, u_in, u_out, u_modes, bias, y, saved, B, Ci, Co, R, ws, n, st):
        assert (B, Ci, Co, R) == (x.shape[0], u_in.shape[0], u_out.shape[0], lam.shape[0])
        self._cp_x = x.detach().clone()
        out = O.contract_cp(x.to(torch.complex64), lam, [u_in, u_out, *u_modes]).real
        y.copy_(out + (bias.reshape(1, -1, *[1] * (out.ndim - 2)) if bias is not None else 0))
        return 0

    def sc_backward_cp(self, plan, gy, lam, u_in, u_out, u_modes, saved, dx, d_lam, d_u_in, d_u_out, d_u_modes, dbias, B, Ci, Co, R, ws, n, st):
        with torch.enable_grad():
            ps = [t.detach().clone().requires_grad_(True) for t in (self._cp_x, lam, u_in, u_out, *u_modes)]
            O.contract_cp(ps[0].to(torch.complex64), ps[1], ps[2:]).real.backward(gy)
        for dst, src in zip([dx, d_lam, d_u_in, d_u_out, *d_u_modes], ps):
            dst.copy_(src.grad)
        if dbias is not None:
            dbias.copy_(gy.sum(dim=[0] + list(range(2, gy.ndim))))
        return 0

    def sc_tt_saved_elems(self, plan, B, Ci, Co, ranks):
        return 0

    def sc_tt_workspace_bytes(self, plan, B, Ci, Co, ranks):
        return 16

    def sc_forward_tt(self, plan, plan_kept, x, g0, g1, cores, bias, y, saved, B, Ci, Co, ranks, ws, n, st):
        assert list(ranks) == [g1.shape[0]] + [c.shape[0] for c in cores]          # {r1, r_0 .. r_{d-1}}
        self._tt_x = x.detach().clone()
        out = O.contract_tt(x.to(torch.complex64), [g0, g1, *cores]).real
        y.copy_(out + (bias.reshape(1, -1, *[1] * (out.ndim - 2)) if bias is not None else 0))
        return 0

    def sc_backward_tt(self, plan, plan_kept, gy, g0, g1, cores, saved, dx, d_g0, d_g1, d_cores, dbias, B, Ci, Co, ranks, ws, n, st):
        with torch.enable_grad():
            ps = [t.detach().clone().requires_grad_(True) for t in (self._tt_x, g0, g1, *cores)]
            O.contract_tt(ps[0].to(torch.complex64), ps[1:]).real.backward(gy)
        for dst, src in zip([dx, d_g0, d_g1, *d_cores], ps):
            dst.copy_(src.grad)
        if dbias is not None:
            dbias.copy_(gy.sum(dim=[0] + list(range(2, gy.ndim))))
        return 0


_LIB = _Lib()


class _Plan:
    def __init__(self, kept):
        self.kept, self.ndim, self.n_modes_total, self.handle = list(kept), len(kept), math.prod(kept), None
        self.grid = self.out_grid = list(kept)


@pytest.fixture
def emulated(monkeypatch):
    """analysis = real -> complex embedding, synthesis = real part (+ bias): a valid adjoint pair, so the transforms drop out."""
    monkeypatch.setattr(sc._lib, "load", lambda: _LIB)
    monkeypatch.setattr(sc, "_ptr_array", lambda ts: list(ts))
    monkeypatch.setattr(sc, "_rank_array", lambda core: [int(r) for r in core.shape])
    monkeypatch.setattr(sc._lib, "check", lambda rc, what: None)
    monkeypatch.setattr(sc, "_ptr", lambda t: t)
    monkeypatch.setattr(sc, "_stream_ptr", lambda dev: None)
    monkeypatch.setattr(sc, "_table_contract", _table_contract)
    monkeypatch.setattr(sc, "_pair_reduce", _pair_reduce)
    monkeypatch.setattr(sc, "_cp_factor_args", lambda us, kept: (list(us), list(kept), len(us)))
    monkeypatch.setattr(sc, "analyze", lambda plan, x, adjoint=False: x.to(torch.complex64))
    monkeypatch.setattr(sc, "synthesize", lambda plan, m, bias=None, adjoint=False:
                        (m.real + (bias.reshape(1, -1, *[1] * (m.ndim - 2)) if bias is not None else 0)).contiguous())
    monkeypatch.setattr(sc, "contract_dense", lambda plan, xm, w: torch.einsum("bi...,io...->bo...", xm, w).contiguous())
    monkeypatch.setattr(sc, "contract_dense_backward", lambda plan, xm, gm, w, **kw: (
        torch.einsum("bo...,io...->bi...", gm, w.conj()).contiguous(),
        torch.einsum("bi...,bo...->io...", xm.conj(), gm).contiguous(), None))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _compare(fn_apply, fn_ref, params, gy):
    p1 = [p.clone().requires_grad_(True) for p in params]
    y1 = fn_apply(*p1)
    y1.backward(gy)
    p2 = [p.clone().requires_grad_(True) for p in params]
    y2 = fn_ref(*p2)
    y2.backward(gy)
    assert _rel(y1, y2) < 1e-5
    for i, (a, b) in enumerate(zip(p1, p2)):
        assert a.grad is not None, f"input {i} got no gradient"
        assert _rel(a.grad, b.grad) < 1e-5, f"gradient of input {i}"


def _c(*shape):
    return torch.randn(*shape, dtype=torch.complex64)


KEPT = [(5,), (4, 3), (3, 4, 2), (2, 3, 2, 2)]


@pytest.mark.parametrize("kept", KEPT)
def test_cp_chain(emulated, kept):
    d, B, Ci, Co, R = len(kept), 2, 3, 4, 5
    torch.manual_seed(0)
    params = [torch.randn(B, Ci, *kept), torch.randn(Co, *[1] * d), _c(R), _c(Ci, R), _c(Co, R), *[_c(k, R) for k in kept]]
    _compare(lambda x, b, lam, ui, uo, *um: sc._SpectralConvCP.apply(x, b, _Plan(kept), lam, ui, uo, *um),
             lambda x, b, lam, ui, uo, *um: O.contract_cp(x.to(torch.complex64), lam, [ui, uo, *um]).real + b,
             params, torch.randn(B, Co, *kept))


@pytest.mark.parametrize("kept", KEPT)
def test_tt_chain(emulated, kept):
    d, B, Ci, Co = len(kept), 2, 3, 4
    torch.manual_seed(1)
    r = [1, 3, 4] + [2 + j for j in range(d - 1)] + [1]
    params = [torch.randn(B, Ci, *kept), torch.randn(Co, *[1] * d), _c(1, Ci, r[1]), _c(r[1], Co, r[2]),
              *[_c(r[2 + j], kept[j], r[3 + j]) for j in range(d)]]
    _compare(lambda x, b, *cores: sc._SpectralConvTT.apply(x, b, _Plan(kept), None, *cores),
             lambda x, b, *cores: O.contract_tt(x.to(torch.complex64), list(cores)).real + b,
             params, torch.randn(B, Co, *kept))


@pytest.mark.parametrize("kept", KEPT)
def test_tucker_chain(emulated, kept):
    d, B, Ci, Co = len(kept), 2, 3, 4
    torch.manual_seed(2)
    ranks = [2, 3] + [2 + (j % 2) for j in range(d)]
    params = [torch.randn(B, Ci, *kept), torch.randn(Co, *[1] * d), _c(*ranks), _c(Ci, ranks[0]), _c(Co, ranks[1]),
              *[_c(k, r) for k, r in zip(kept, ranks[2:])]]
    _compare(lambda x, b, core, ui, uo, *um: sc._SpectralConvTucker.apply(x, b, _Plan(kept), _Plan(kept), core, ui, uo, *um),
             lambda x, b, core, ui, uo, *um: O.contract_tucker(x.to(torch.complex64), core, [ui, uo, *um]).real + b,
             params, torch.randn(B, Co, *kept))


@pytest.mark.parametrize("kept", KEPT)
def test_separable_chain(emulated, kept):
    d, B, C = len(kept), 2, 3
    torch.manual_seed(3)
    params = [torch.randn(B, C, *kept), _c(C, *kept), torch.randn(C, *[1] * d)]
    _compare(lambda x, w, b: sc._SpectralConvSeparable.apply(x, w, b, _Plan(kept)),
             lambda x, w, b: (x.to(torch.complex64) * w).real + b,
             params, torch.randn(B, C, *kept))


def test_double_backward_is_refused_loudly(emulated):
    """The kernels implement first-order gradients only; differentiating the backward pass must not silently return zeros."""
    kept = (4, 3)
    x = torch.randn(2, 3, *kept, requires_grad=True)
    w = _c(3, *kept).requires_grad_(True)
    y = sc._SpectralConvSeparable.apply(x, w, None, _Plan(kept))
    (gx,) = torch.autograd.grad(y.square().sum(), x, create_graph=True)
    with pytest.raises(RuntimeError, match="once_differentiable|differentiate twice"):
        gx.sum().backward()


@pytest.mark.parametrize("grid,modes,kw", [
    ((16, 12), (8, 6), {}), ((16,), (6,), {}), ((8, 6, 10), (4, 4, 6), {}), ((9, 11), (4, 5), {}), ((12, 12), (16, 16), {}),
    ((16, 12), (6, 4), {"max_n_modes": (8, 6)}), ((16, 12), (5, 3), {"max_n_modes": (8, 6)}),
    ((12, 12), (10, 8), {"out": (24, 24)}), ((12, 12), (10, 8), {"out": (6, 6)}), ((12, 13), (6, 6), {"out": (9, 16)}),
    ((12, 12), (10, 8), {"fft_norm": "ortho"}), ((12, 12), (10, 8), {"fft_norm": "backward"}), ((6, 6, 6, 6), (4, 4, 4, 4), {}),
])
def test_complex_data_chain(emulated, monkeypatch, grid, modes, kw):
    """complex_data=True is composed on the host from the complex table kernel and the dense mode GEMM: with those two emulated,
    the tables, the index rules (shift on every dim, central weight cut, first-k rule on the shifted last dim, no un-shift of
    the last dim) and the adjoint chain of the backward pass are checked against the oracle's restatement of the reference
    (pinned bit-exactly to the live module in tests/test_oracle_vs_reference.py), differentiated by torch."""
    monkeypatch.setattr(sc, "get_plan", lambda *a, **k: _Plan(a[3]))
    torch.manual_seed(5)
    B, Ci, Co = 2, 3, 4
    maxm = list(kw.get("max_n_modes", modes))
    out = list(kw.get("out", grid))
    norm = kw.get("fft_norm", "forward")
    plan = sc.ComplexPlan(torch.device("cpu"), list(grid), out, list(modes), maxm, norm)
    x = torch.randn(B, Ci, *grid, dtype=torch.complex64)
    w = _c(Ci, Co, *maxm)
    bias = torch.randn(Co, *[1] * len(grid))
    gy = torch.randn(B, Co, *out, dtype=torch.complex64)

    def ours(x_, w_, b_, lead=2):
        wk = w_
        for j in range(len(grid)):
            wk = wk.narrow(lead + j, plan.w_start[j], plan.kept[j])
        return sc._SpectralConvComplex.apply(x_, wk.contiguous(), plan, lead == 1) + b_

    def ref(x_, w_, b_, separable=False):
        return O.spectral_conv_forward_complex(x_, w_, b_, list(modes), max_n_modes=maxm, output_shape=out, fft_norm=norm,
                                               separable=separable)

    _compare(ours, ref, [x, w, bias], gy)
    # separable=True: one channel axis, mode-wise product (sc_cp_apply / sc_cp_dscale in place of the mode GEMM)
    ws = _c(Ci, *maxm)
    _compare(lambda x_, w_, b_: ours(x_, w_, b_, 1), lambda x_, w_, b_: ref(x_, w_, b_, True),
             [x, ws, torch.randn(Ci, *[1] * len(grid))], torch.randn(B, Ci, *out, dtype=torch.complex64))


from conftest import complex_golden_index, load_complex_golden  # noqa: E402


@pytest.mark.parametrize("name", sorted(complex_golden_index().keys()))
def test_complex_module_forward_against_reference_goldens(emulated, monkeypatch, name):
    """The whole complex_data module path (weight slicing, reconstructed Tucker weights, resampling, bias) with the two device
    primitives emulated, against what the unmodified reference returned: y, dx and every parameter gradient."""
    import neuraloperator_b200 as nb
    monkeypatch.setattr(sc, "get_plan", lambda *a, **k: _Plan(a[3]))
    monkeypatch.setattr(sc, "get_complex_plan", lambda dev, grid, out, nm, mx, norm: sc.ComplexPlan(torch.device("cpu"), grid, out, nm, mx, norm))
    meta, arr = load_complex_golden(name)
    conv = nb.SpectralConv(meta["in_channels"], meta["out_channels"], tuple(meta["n_modes"]), complex_data=True, **dict(meta["ctor"]))
    assert conv.n_modes == meta["stored_n_modes"] and list(conv.max_n_modes) == meta["max_n_modes"]
    params = dict(conv.named_parameters())
    with torch.no_grad():
        for pname in meta["params"]:
            ours = pname.replace("weight.factors.", "weight.factors.factor_")
            params[ours].copy_(arr["p__" + pname.replace(".", "__")])
    x = arr["x"].clone().requires_grad_(True)
    y = conv._forward_complex(x, meta["forward"].get("output_shape"))
    assert list(y.shape[2:]) == meta["out_grid"]
    y.backward(arr["gy"])
    assert _rel(y.detach(), arr["y"]) < 2e-5 and _rel(x.grad, arr["dx"]) < 2e-5
    for pname in meta["params"]:
        ours = pname.replace("weight.factors.", "weight.factors.factor_")
        assert _rel(params[ours].grad, arr["g__" + pname.replace(".", "__")]) < 5e-5, pname


def test_complex_plan_cache_does_not_deadlock(monkeypatch):
    """get_complex_plan builds a ComplexPlan under the plan-cache lock, and the ComplexPlan asks get_plan (same lock) for its
    contraction plan: with only the C plan object faked, the real cache code must return (run on a watchdog thread, so that
    a regression fails instead of hanging the suite) and must hand back the cached object the second time."""
    import threading

    class _FakeCPlan:
        def __init__(self, device, grid, out_grid, n_modes, max_n_modes, fft_norm, flags=0):
            self.args = (tuple(grid), tuple(out_grid), tuple(n_modes), tuple(max_n_modes), fft_norm, flags)

    real = sc.ComplexPlan
    monkeypatch.setattr(sc, "Plan", _FakeCPlan)
    monkeypatch.setattr(sc, "ComplexPlan", lambda dev, *a: real(dev, *a, table_device=torch.device("cpu")))
    monkeypatch.setattr(sc, "_PLAN_CACHE", type(sc._PLAN_CACHE)())
    monkeypatch.setattr(sc, "_COMPLEX_PLANS", type(sc._COMPLEX_PLANS)())
    out = {}

    def work():
        dev = torch.device("cuda", 0)
        out["a"] = sc.get_complex_plan(dev, [16, 12], [16, 12], [8, 6], [8, 6], "forward")
        out["b"] = sc.get_complex_plan(dev, [16, 12], [16, 12], [8, 6], [8, 6], "forward")

    t = threading.Thread(target=work, daemon=True)
    t.start()
    t.join(20)
    assert not t.is_alive(), "get_complex_plan deadlocked on the plan-cache lock"
    assert out["a"] is out["b"] and out["a"].kept == (8, 6)
    # the contraction plan: kept block (8, 6) of a real-data problem whose last dim is twice as long
    assert out["a"].contract_plan.args == ((8, 12), (8, 12), (8, 6), (8, 6), "forward", 0)


@pytest.mark.parametrize("kept", KEPT)
def test_cp_chain_single_c_call_plumbing(emulated, kept):
    """`_SpectralConvCPCall`: argument order, output / gradient buffers and which gradient goes where (the chain itself lives in C)."""
    d, B, Ci, Co, R = len(kept), 2, 3, 4, 5
    torch.manual_seed(0)
    params = [torch.randn(B, Ci, *kept), torch.randn(Co, *[1] * d), _c(R), _c(Ci, R), _c(Co, R), *[_c(k, R) for k in kept]]
    _compare(lambda x, b, lam, ui, uo, *um: sc._SpectralConvCPCall.apply(x, b, _Plan(kept), lam, ui, uo, *um),
             lambda x, b, lam, ui, uo, *um: O.contract_cp(x.to(torch.complex64), lam, [ui, uo, *um]).real + b,
             params, torch.randn(B, Co, *kept))


@pytest.mark.parametrize("kept", KEPT)
def test_tt_chain_single_c_call_plumbing(emulated, kept):
    d, B, Ci, Co = len(kept), 2, 3, 4
    torch.manual_seed(1)
    r = [1, 3, 4] + [2 + j for j in range(d - 1)] + [1]
    params = [torch.randn(B, Ci, *kept), torch.randn(Co, *[1] * d), _c(1, Ci, r[1]), _c(r[1], Co, r[2]),
              *[_c(r[2 + j], kept[j], r[3 + j]) for j in range(d)]]
    _compare(lambda x, b, *cores: sc._SpectralConvTTCall.apply(x, b, _Plan(kept), _Plan(kept), *cores),
             lambda x, b, *cores: O.contract_tt(x.to(torch.complex64), list(cores)).real + b,
             params, torch.randn(B, Co, *kept))
