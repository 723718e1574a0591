"""Fourier-layer epilogue kernels (csrc/sc_layer.cu) checked WITHOUT a GPU.

The four kernels are sequences of `__host__ __device__` tile functions; the `sc_hostcheck_*` entry points of the library run exactly
those functions thread by thread, block by block, on host buffers.  What is checked here is therefore the index arithmetic, the tile
edges (extents that are not multiples of the tiles), the argument options and the activation formulas of the very code the GPU
executes -- against plain torch restatements of the reference ops (conv1d with a 1x1 kernel, add, F.gelu, soft gating;
neuralop/layers/fno_block.py:377-414, channel_mlp.py:92-116, skip_connections.py:85-130).  What a CPU cannot check: launch
configuration, shared-memory races, atomics -- the kernels keep those trivial (two __syncthreads per stage, atomicAdd epilogues)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from neuraloperator_b200 import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def host_channel_mix(x, w, s_o, s_i, bias, add, gate, gated, act, B, Ci, Co, P, want_pre=True):
    lib = _lib.load()
    out = torch.full((B, Co, P), float("nan"))
    pre = torch.full((B, Co, P), float("nan")) if want_pre else None
    rc = lib.sc_hostcheck_channel_mix(_p(x), _p(w), s_o, s_i, _p(bias), _p(add), _p(gate), _p(gated), act, _p(out), _p(pre), B, Ci, Co, P)
    assert rc == 0, lib.sc_last_error()
    return out, pre


def ref_channel_mix(x, wmat, bias, add, gate, gated, act, B, Co, P):
    pre = torch.zeros(B, Co, P, dtype=torch.float64)
    if x is not None and wmat is not None and wmat.numel():
        pre = torch.einsum("oi,bip->bop", wmat.double(), x.double())
    if bias is not None:
        pre = pre + bias.double().view(1, -1, 1)
    if add is not None:
        pre = pre + add.double()
    if gated is not None:
        pre = pre + (gate.double().view(1, -1, 1) if gate is not None else 1.0) * gated.double()
    out = F.gelu(pre) if act == _lib.ACT_GELU else pre
    return out, pre


SHAPES = [(1, 1, 1, 1), (2, 3, 5, 7), (1, 16, 64, 128), (2, 17, 65, 129), (1, 64, 64, 300), (3, 70, 130, 33), (1, 33, 8, 1000)]


@pytest.mark.parametrize("B,Ci,Co,P", SHAPES)
@pytest.mark.parametrize("opts", ["plain", "bias_add_gelu", "gate", "all", "identity_skip"])
def test_channel_mix_tiles_match_torch(B, Ci, Co, P, opts):
    g = torch.Generator().manual_seed(B * 1000 + Ci * 100 + Co * 10 + P)
    x = torch.randn(B, Ci, P, generator=g)
    w = torch.randn(Co, Ci, generator=g) / max(Ci, 1) ** 0.5
    bias = torch.randn(Co, generator=g) if opts in ("bias_add_gelu", "all") else None
    add = torch.randn(B, Co, P, generator=g) if opts in ("bias_add_gelu", "all") else None
    gate = torch.randn(Co, generator=g) if opts in ("gate", "all") else None
    gated = torch.randn(B, Co, P, generator=g) if opts in ("gate", "all", "identity_skip") else None
    act = _lib.ACT_GELU if opts in ("bias_add_gelu", "all") else _lib.ACT_IDENTITY
    out, pre = host_channel_mix(x, w, Ci, 1, bias, add, gate, gated, act, B, Ci, Co, P)
    ref_out, ref_pre = ref_channel_mix(x, w, bias, add, gate, gated, act, B, Co, P)
    assert not torch.isnan(out).any() and not torch.isnan(pre).any()        # every output element was written
    assert (pre.double() - ref_pre).abs().max() < 2e-5 * max(1.0, ref_pre.abs().max().item())
    assert (out.double() - ref_out).abs().max() < 2e-5 * max(1.0, ref_out.abs().max().item())


def test_channel_mix_transposed_strides_and_no_pre():
    """The input gradient is the same kernel with the weight read through transposed strides."""
    B, Ci, Co, P = 2, 37, 21, 150
    g = torch.Generator().manual_seed(5)
    w = torch.randn(Ci, Co, generator=g)            # stored (Ci, Co); used as W[o, i] = w[i, o]
    x = torch.randn(B, Ci, P, generator=g)
    out, pre = host_channel_mix(x, w, 1, Co, None, None, None, None, _lib.ACT_IDENTITY, B, Ci, Co, P, want_pre=False)
    assert pre is None
    ref = torch.einsum("io,bip->bop", w.double(), x.double())
    assert (out.double() - ref).abs().max() < 2e-5 * ref.abs().max()


def test_channel_mix_without_mixing_term():
    """in_channels == 0: out = act(bias + add + gate * gated) -- the add + activation after a skip that had to be resampled."""
    B, Co, P = 2, 9, 77
    g = torch.Generator().manual_seed(6)
    add, gated, gate, bias = torch.randn(B, Co, P, generator=g), torch.randn(B, Co, P, generator=g), torch.randn(Co, generator=g), torch.randn(Co, generator=g)
    out, pre = host_channel_mix(None, None, 0, 0, bias, add, gate, gated, _lib.ACT_GELU, B, 0, Co, P)
    ref_out, ref_pre = ref_channel_mix(None, None, bias, add, gate, gated, _lib.ACT_GELU, B, Co, P)
    assert (out.double() - ref_out).abs().max() < 1e-5 and (pre.double() - ref_pre).abs().max() < 1e-5


TORCH_ACT = {_lib.ACT_IDENTITY: lambda t: t, _lib.ACT_GELU: F.gelu, _lib.ACT_RELU: F.relu, _lib.ACT_SILU: F.silu, _lib.ACT_TANH: torch.tanh}


@pytest.mark.parametrize("act", [_lib.ACT_RELU, _lib.ACT_SILU, _lib.ACT_TANH, _lib.ACT_GELU])
def test_every_activation_forward_and_derivative(act):
    """out = act(pre) in the mixing kernel and gpre = gout * act'(pre) in the backward kernel against torch's function and autograd."""
    lib = _lib.load()
    B, C, P = 2, 5, 333
    g = torch.Generator().manual_seed(act)
    add = torch.randn(B, C, P, generator=g) * 2
    out, pre = host_channel_mix(None, None, 0, 0, None, add, None, None, act, B, 0, C, P)
    ref_in = add.double().requires_grad_(True)
    ref = TORCH_ACT[act](ref_in)
    assert (out.double() - ref.detach()).abs().max() < 1e-6
    assert torch.equal(pre, add)
    gout = torch.randn(B, C, P, generator=g)
    ref.backward(gout.double())
    gpre = torch.full((B, C, P), float("nan"))
    assert lib.sc_hostcheck_channel_mix_act_backward(_p(gout), _p(pre), act, None, None, _p(gpre), None, None, None, B, C, P) == 0
    assert (gpre.double() - ref_in.grad).abs().max() < 2e-6


def test_channel_mix_rejects_bad_arguments():
    lib = _lib.load()
    out = torch.empty(1, 1, 1)
    x = torch.empty(1, 1, 1)
    assert lib.sc_hostcheck_channel_mix(_p(x), _p(None), 1, 1, None, None, None, None, 0, _p(out), None, 1, 1, 1, 1) != 0     # w missing
    assert lib.sc_hostcheck_channel_mix(_p(x), _p(x), 1, 1, None, None, _p(x), None, 0, _p(out), None, 1, 1, 1, 1) != 0       # gate without gated
    assert lib.sc_hostcheck_channel_mix(_p(x), _p(x), 1, 1, None, None, None, None, 7, _p(out), None, 1, 1, 1, 1) != 0        # unknown act
    assert b"activation" in lib.sc_last_error()


@pytest.mark.parametrize("B,C,P", [(1, 1, 1), (2, 3, 50), (1, 5, 4096), (2, 7, 4097), (1, 2, 10000)])
@pytest.mark.parametrize("act", [_lib.ACT_IDENTITY, _lib.ACT_GELU])
@pytest.mark.parametrize("with_gate", [False, True])
def test_act_backward_matches_autograd(B, C, P, act, with_gate):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + C + P)
    pre = torch.randn(B, C, P, generator=g, dtype=torch.float64).requires_grad_(True)
    gate = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_(True) if with_gate else None
    gated = torch.randn(B, C, P, generator=g, dtype=torch.float64).requires_grad_(True) if with_gate else None
    gout = torch.randn(B, C, P, generator=g, dtype=torch.float64)
    # reference: out = act(pre), pre = stuff + bias + gate * gated  => d(pre) = gout * act'(pre), dbias = sum, dgate = sum(gpre * gated)
    base = torch.zeros(B, C, P, dtype=torch.float64, requires_grad=True)
    bias = torch.zeros(C, dtype=torch.float64, requires_grad=True)
    full = base + bias.view(1, -1, 1) + pre.detach() + ((gate.view(1, -1, 1) * gated) if with_gate else 0)
    (F.gelu(full) if act == _lib.ACT_GELU else full).backward(gout)
    pre_val = full.detach().float().contiguous()
    gpre = torch.full((B, C, P), float("nan"))
    dgated = torch.full((B, C, P), float("nan")) if with_gate else None
    dbias = torch.full((C,), float("nan"))
    dgate = torch.full((C,), float("nan")) if with_gate else None
    gout32 = gout.float().contiguous()
    gate32 = gate.detach().float().contiguous() if with_gate else None          # (kept alive across the call: _p takes raw pointers)
    gated32 = gated.detach().float().contiguous() if with_gate else None
    rc = lib.sc_hostcheck_channel_mix_act_backward(_p(gout32), _p(pre_val), act, _p(gate32), _p(gated32), _p(gpre), _p(dgated),
                                                   _p(dbias), _p(dgate), B, C, P)
    assert rc == 0, lib.sc_last_error()
    tol = 3e-5
    assert (gpre.double() - base.grad).abs().max() < tol * max(1.0, base.grad.abs().max().item())
    assert (dbias.double() - bias.grad).abs().max() < tol * max(1.0, bias.grad.abs().max().item()) * (B * P) ** 0.5
    if with_gate:
        assert (dgated.double() - gated.grad).abs().max() < tol * max(1.0, gated.grad.abs().max().item())
        assert (dgate.double() - gate.grad).abs().max() < tol * max(1.0, gate.grad.abs().max().item()) * (B * P) ** 0.5


def test_act_backward_in_place_and_reductions_only():
    """gpre may alias gout; with the identity activation and no gpre output the call is just the dbias reduction."""
    lib = _lib.load()
    B, C, P = 2, 4, 5000
    g = torch.Generator().manual_seed(11)
    gout = torch.randn(B, C, P, generator=g)
    pre = torch.randn(B, C, P, generator=g)
    want = gout.double() * torch.autograd.functional.jacobian(lambda t: F.gelu(t).sum(), pre.double())
    buf = gout.clone()
    assert lib.sc_hostcheck_channel_mix_act_backward(_p(buf), _p(pre), _lib.ACT_GELU, None, None, _p(buf), None, None, None, B, C, P) == 0
    assert (buf.double() - want).abs().max() < 3e-5
    dbias = torch.full((C,), float("nan"))
    assert lib.sc_hostcheck_channel_mix_act_backward(_p(gout), None, _lib.ACT_IDENTITY, None, None, None, None, _p(dbias), None, B, C, P) == 0
    assert (dbias.double() - gout.double().sum(dim=(0, 2))).abs().max() < 1e-3


@pytest.mark.parametrize("B,Ci,Co,P", [(1, 1, 1, 1), (2, 3, 5, 40), (1, 64, 64, 2048), (2, 65, 70, 2049), (1, 130, 17, 100), (3, 8, 4, 4500)])
def test_weight_grad_tiles_match_torch(B, Ci, Co, P):
    lib = _lib.load()
    g = torch.Generator().manual_seed(Ci * 7 + Co)
    gp = torch.randn(B, Co, P, generator=g)
    x = torch.randn(B, Ci, P, generator=g)
    dw = torch.full((Co, Ci), float("nan"))
    assert lib.sc_hostcheck_channel_mix_weight_grad(_p(gp), _p(x), _p(dw), B, Ci, Co, P) == 0, lib.sc_last_error()
    ref = torch.einsum("bop,bip->oi", gp.double(), x.double())
    assert (dw.double() - ref).abs().max() < 3e-5 * max(1.0, ref.abs().max().item()) * (B * P) ** 0.5


@pytest.mark.parametrize("n", [0, 1, 255, 2048, 2049, 10001])
def test_pointwise_ops(n):
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    a = torch.randn(n, generator=g) * 3
    b = torch.tanh(torch.randn(n, generator=g))
    out = torch.full((n,), float("nan"))
    assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_TANH, _p(a), None, _p(out), n) == 0
    assert n == 0 or (out - torch.tanh(a)).abs().max() < 1e-6
    assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_TANH_BACKWARD, _p(a), _p(b), _p(out), n) == 0
    assert n == 0 or (out - a * (1 - b * b)).abs().max() < 1e-6
    big = torch.cat([a * 1e3, torch.tensor([65504.0, 1e-8, -0.0, 70000.0])])[: max(n, 0)] if n else a
    out2 = torch.full((big.numel(),), float("nan"))
    assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_ROUND_HALF, _p(big), None, _p(out2), big.numel()) == 0
    assert torch.equal(out2, big.half().float())           # bit-exact: round-to-nearest-even fp16, as torch's .half()
    assert lib.sc_hostcheck_pointwise(9, _p(a), None, _p(out), n) != 0
    assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_MUL, _p(a), _p(b), _p(out), n) == 0
    assert n == 0 or torch.equal(out, a * b)
    if n % 2 == 0 and n:                                     # the complex-pair ops: out = a + 1j * b, out = -1j * a on interleaved pairs
        ac, bc = torch.view_as_complex(a.view(-1, 2)), torch.view_as_complex(b.view(-1, 2).contiguous())
        assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_ADD_I_TIMES, _p(a), _p(b), _p(out), n) == 0
        assert torch.equal(torch.view_as_complex(out.view(-1, 2)), torch.complex(ac.real - bc.imag, ac.imag + bc.real))
        assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_MUL_NEG_I, _p(a), None, _p(out), n) == 0
        assert torch.equal(torch.view_as_complex(out.view(-1, 2)), torch.complex(ac.imag, -ac.real))
        assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_ADD_I_TIMES, _p(a), _p(b), _p(a), n) != 0       # in place: refused
    elif n % 2 == 1:
        assert lib.sc_hostcheck_pointwise(_lib.POINTWISE_MUL_NEG_I, _p(a), None, _p(out), n) != 0        # odd count: refused


def test_whole_layer_epilogue_on_the_host_tiles_matches_reference_ops():
    """f1 + f2 composed from the kernels' host checks == the torch ops FNOBlocks.forward_with_postactivation runs after the conv
    (fno_block.py:392-412) -- forward and every gradient, with the conv output as a given tensor."""
    lib = _lib.load()
    B, C, H, P = 2, 12, 6, 90
    g = torch.Generator().manual_seed(3)
    dd = dict(dtype=torch.float64)
    x = torch.randn(B, C, P, generator=g, **dd).requires_grad_(True)
    x_fno = torch.randn(B, C, P, generator=g, **dd).requires_grad_(True)
    w_skip = (torch.randn(C, C, 1, generator=g, **dd) / C ** 0.5).requires_grad_(True)
    w1 = (torch.randn(H, C, 1, generator=g, **dd) / C ** 0.5).requires_grad_(True)
    b1 = torch.randn(H, generator=g, **dd).requires_grad_(True)
    w2 = (torch.randn(C, H, 1, generator=g, **dd) / H ** 0.5).requires_grad_(True)
    b2 = torch.randn(C, generator=g, **dd).requires_grad_(True)
    gate = torch.randn(1, C, 1, generator=g, **dd).requires_grad_(True)
    x1 = F.gelu(x_fno + F.conv1d(x, w_skip))
    out = F.gelu(F.conv1d(F.gelu(F.conv1d(x1, w1, b1)), w2, b2) + gate * x)
    gout = torch.randn(B, C, P, generator=g, **dd)
    out.backward(gout)

    f = lambda t: t.detach().float().contiguous() if t is not None else None
    X, XF = f(x), f(x_fno)
    x1_k, pre1 = host_channel_mix(X, f(w_skip), C, 1, None, XF, None, None, _lib.ACT_GELU, B, C, C, P)
    h_k, pre2 = host_channel_mix(x1_k, f(w1), C, 1, f(b1), None, None, None, _lib.ACT_GELU, B, C, H, P)
    out_k, pre3 = host_channel_mix(h_k, f(w2), H, 1, f(b2), None, f(gate).view(-1), X, _lib.ACT_GELU, B, H, C, P)
    assert (out_k.double() - out.detach()).abs().max() < 1e-5

    def act_bwd(go, pre, gate_, gated_, Cc):
        gp = torch.empty(B, Cc, P)
        dgd = torch.empty(B, Cc, P) if gated_ is not None else None
        db = torch.empty(Cc)
        dg = torch.empty(Cc) if gated_ is not None else None
        assert lib.sc_hostcheck_channel_mix_act_backward(_p(go), _p(pre), _lib.ACT_GELU, _p(gate_), _p(gated_), _p(gp), _p(dgd), _p(db), _p(dg),
                                                         B, Cc, P) == 0
        return gp, dgd, db, dg

    def wgrad(gp, inp, Ci, Co):
        dw = torch.empty(Co, Ci)
        assert lib.sc_hostcheck_channel_mix_weight_grad(_p(gp), _p(inp), _p(dw), B, Ci, Co, P) == 0
        return dw

    def din(gp, w, Ci, Co):       # gradient w.r.t. the mixed input: W^T gp  (w stored (Co, Ci))
        o, _ = host_channel_mix(gp, w, 1, Ci, None, None, None, None, _lib.ACT_IDENTITY, B, Co, Ci, P, want_pre=False)
        return o

    gp3, dx_gate, db2_k, dgate_k = act_bwd(f(gout), pre3, f(gate).view(-1), X, C)
    dw2_k = wgrad(gp3, h_k, H, C)
    gp2, _, db1_k, _ = act_bwd(din(gp3, f(w2), H, C), pre2, None, None, H)
    dw1_k = wgrad(gp2, x1_k, C, H)
    gp1, _, _, _ = act_bwd(din(gp2, f(w1), C, H), pre1, None, None, C)
    dwskip_k = wgrad(gp1, X, C, C)
    dx_k = din(gp1, f(w_skip), C, C) + dx_gate
    for name, got, ref in [("dx", dx_k, x.grad), ("dx_fno", gp1, x_fno.grad), ("dw_skip", dwskip_k, w_skip.grad[..., 0]),
                           ("dw1", dw1_k, w1.grad[..., 0]), ("db1", db1_k, b1.grad), ("dw2", dw2_k, w2.grad[..., 0]),
                           ("db2", db2_k, b2.grad), ("dgate", dgate_k, gate.grad.view(-1))]:
        err = (got.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)
        assert err < 2e-5, (name, err)


# ---- randomised extents (hypothesis): every tile edge combination, every option subset ------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=80, deadline=None, derandomize=True)
@given(B=st.integers(1, 3), Ci=st.integers(0, 70), Co=st.integers(1, 70), P=st.integers(1, 300), has_bias=st.booleans(), has_add=st.booleans(),
       gate_kind=st.sampled_from(["none", "gated", "gate+gated"]), gelu=st.booleans(), seed=st.integers(0, 10_000))
def test_channel_mix_random_extents_and_options(B, Ci, Co, P, has_bias, has_add, gate_kind, gelu, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Ci, P, generator=g) if Ci else None
    w = torch.randn(Co, Ci, generator=g) / max(Ci, 1) ** 0.5 if Ci else None
    bias = torch.randn(Co, generator=g) if has_bias else None
    add = torch.randn(B, Co, P, generator=g) if has_add else None
    gated = torch.randn(B, Co, P, generator=g) if gate_kind != "none" else None
    gate = torch.randn(Co, generator=g) if gate_kind == "gate+gated" else None
    act = _lib.ACT_GELU if gelu else _lib.ACT_IDENTITY
    out, pre = host_channel_mix(x, w, Ci, 1, bias, add, gate, gated, act, B, Ci, Co, P)
    ref_out, ref_pre = ref_channel_mix(x, w, bias, add, gate, gated, act, B, Co, P)
    assert not torch.isnan(out).any() and not torch.isnan(pre).any()
    assert (pre.double() - ref_pre).abs().max() < 2e-5 * max(1.0, ref_pre.abs().max().item())
    assert (out.double() - ref_out).abs().max() < 2e-5 * max(1.0, ref_out.abs().max().item())


@settings(max_examples=40, deadline=None, derandomize=True)
@given(B=st.integers(1, 3), Ci=st.integers(1, 70), Co=st.integers(1, 70), P=st.integers(1, 5000), seed=st.integers(0, 10_000))
def test_weight_grad_random_extents(B, Ci, Co, P, seed):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    gp = torch.randn(B, Co, P, generator=g)
    x = torch.randn(B, Ci, P, generator=g)
    dw = torch.full((Co, Ci), float("nan"))
    assert lib.sc_hostcheck_channel_mix_weight_grad(_p(gp), _p(x), _p(dw), B, Ci, Co, P) == 0
    ref = torch.einsum("bop,bip->oi", gp.double(), x.double())
    assert (dw.double() - ref).abs().max() < 3e-5 * max(1.0, ref.abs().max().item()) * (B * P) ** 0.5
