"""The CP / TT chains behind one C call per direction (sc_forward_cp / sc_backward_cp, sc_forward_tt / sc_backward_tt), checked WITHOUT
a GPU: `sc_hostcheck_chain_log` runs the very C code of a chain on a host-only plan with a recorder installed, so that every primitive
launch -- opcode, buffers (synthetic addresses), strides, extents -- is logged instead of executed.  This file replays the log on host
arrays, each primitive by its documented contract (include/spectral_conv_b200.h; the primitives themselves are validated on
hardware), and compares y, dx and every factor gradient with the oracle's einsum statement of the reference contraction
(`_contract_cp` :55-73, `_contract_tt` :106-127) differentiated by autograd.  What is checked is everything the C chain decides:
the order of the launches, which offset of the saved buffer / workspace is which operand, every stride and extent."""
import ctypes
import math

import numpy as np
import pytest
import torch

from neuraloperator_b200 import _lib
from oracle import spectral_conv_oracle as O

(CH_ANALYZE, CH_SYNTHESIZE, CH_TABLE, CH_PAIR, CH_CP_SCALE, CH_CP_APPLY, CH_CP_DSCALE, CH_CP_FACTOR_GRAD, CH_BIAS_GRAD, CH_CONTRACT_FWD,
 CH_CONTRACT_BWD) = range(1, 12)
R_X, R_Y, R_SAVED, R_WS, R_LAM, R_UIN, R_UOUT, R_MODE0, R_BIAS, R_DLAM, R_DUIN, R_DUOUT, R_DMODE0 = 1, 2, 3, 4, 5, 6, 7, 8, 12, 15, 16, 17, 18


def chain_log(grid, modes_user, kind, direction, B, Ci, Co, ranks):
    lib = _lib.load()
    stored = O.stored_n_modes(modes_user)
    prob = _lib.ScProblem()
    prob.ndim = len(grid)
    for j in range(len(grid)):
        prob.grid[j] = prob.out_grid[j] = grid[j]
        prob.n_modes[j] = prob.max_n_modes[j] = stored[j]
    prob.fft_norm, prob.flags = 0, 0
    rk = (ctypes.c_int32 * len(ranks))(*ranks)
    n = ctypes.c_int64(0)
    assert lib.sc_hostcheck_chain_log(ctypes.byref(prob), kind, direction, B, Ci, Co, rk, None, 0, ctypes.byref(n)) == 0, lib.sc_last_error()
    buf = (ctypes.c_int64 * n.value)()
    assert lib.sc_hostcheck_chain_log(ctypes.byref(prob), kind, direction, B, Ci, Co, rk, buf, n.value, ctypes.byref(n)) == 0
    words, ops, i = list(buf), [], 0
    while i < len(words):
        ops.append((words[i], words[i + 2: i + 2 + words[i + 1]]))
        i += 2 + words[i + 1]
    return ops


class Memory:
    """One byte array per synthetic region; typed views at (region << 40) + offset."""

    def __init__(self):
        self.regions = {}
        self.limits = {}             # bytes the Python side allocates for a region (saved buffer, workspace), from record 0 of a log

    def _buf(self, region):
        if region not in self.regions:
            self.regions[region] = np.zeros(1 << 22, dtype=np.uint8)
            self.regions[region][:] = 0x7F                              # poison: reading something never written shows up as garbage
        return self.regions[region]

    def view(self, addr, n, dtype):
        region, off = addr >> 40, addr & ((1 << 40) - 1)
        item = np.dtype(dtype).itemsize
        assert off % item == 0 and off + n * item <= self.limits.get(region, 1 << 22), ("out of bounds", region, off, n * item)
        return self._buf(region)[off: off + n * item].view(dtype)

    def put(self, region, array):
        a = np.ascontiguousarray(array)
        self.view(region << 40, a.size, a.dtype)[:] = a.reshape(-1)

    def get(self, region, shape, dtype):
        return self.view(region << 40, math.prod(shape), dtype).reshape(shape).copy()


def replay(ops, mem, plans, grid, B_, fft_norm="forward"):
    c64, f32 = np.complex64, np.float32
    kept = [p.kept for p in plans]
    M = math.prod(kept)
    t = torch.from_numpy
    for op, a in ops:
        if op == 0:                  # allocation sizes of the saved buffer (elements) and the workspace (bytes)
            mem.limits[R_SAVED], mem.limits[R_WS] = a[0] * 8, a[1]
            assert 0 < mem.limits[R_SAVED] <= 1 << 22 and 0 < mem.limits[R_WS] <= 1 << 22
        elif op == CH_ANALYZE:
            images, n_images, out, adjoint = a
            if not adjoint:
                x = t(mem.view(images, n_images * math.prod(grid), f32).reshape(1, n_images, *grid).copy())
                mem.view(out, n_images * M, c64)[:] = O.analyze_modes(x, plans, fft_norm).numpy().reshape(-1)
            else:                                                      # adjoint of the synthesis: what autograd applies to gy
                gy = t(mem.view(images, n_images * math.prod(grid), f32).reshape(1, n_images, *grid).copy())
                m0 = torch.zeros(1, n_images, *kept, dtype=torch.cfloat, requires_grad=True)
                gm = torch.autograd.grad(O.synthesize_modes(m0, plans, grid, fft_norm), m0, gy)[0]
                mem.view(out, n_images * M, c64)[:] = gm.numpy().reshape(-1)
        elif op == CH_SYNTHESIZE:
            modes, n_images, n_channels, bias, images, adjoint = a
            m = t(mem.view(modes, n_images * M, c64).reshape(1, n_images, *kept).copy())
            if not adjoint:
                y = O.synthesize_modes(m, plans, grid, fft_norm)
                if bias:
                    bvec = t(mem.view(bias, n_channels, f32).copy())
                    y = y + bvec.repeat(n_images // n_channels).view(1, -1, *[1] * len(grid))   # image n uses bias[n % n_channels]
            else:
                assert not bias
                x0 = torch.zeros(1, n_images, *grid, requires_grad=True)
                y = torch.autograd.grad(O.analyze_modes(x0, plans, fft_norm), x0, m)[0]
            mem.view(images, n_images * math.prod(grid), f32)[:] = y.detach().numpy().reshape(-1)
        elif op == CH_TABLE:
            T, s_p, s_q, conj, src, dst, n_outer, P, Q, n_inner = a
            span = (P - 1) * s_p + (Q - 1) * s_q + 1
            tab = mem.view(T, span, c64)
            idx = np.arange(P)[:, None] * s_p + np.arange(Q)[None, :] * s_q
            Tm = tab[idx].conj() if conj else tab[idx]
            x = mem.view(src, n_outer * Q * n_inner, c64).reshape(n_outer, Q, n_inner)
            mem.view(dst, n_outer * P * n_inner, c64)[:] = np.einsum("pq,oqi->opi", Tm, x).reshape(-1)
        elif op == CH_PAIR:
            A, Bp, out, s_p, s_q, n_outer, P, Q, n_inner = a
            am = mem.view(A, n_outer * P * n_inner, c64).reshape(n_outer, P, n_inner)
            bm = mem.view(Bp, n_outer * Q * n_inner, c64).reshape(n_outer, Q, n_inner)
            r = np.einsum("opi,oqi->pq", am.conj(), bm)
            span = (P - 1) * s_p + (Q - 1) * s_q + 1
            o = mem.view(out, span, c64)
            o[(np.arange(P)[:, None] * s_p + np.arange(Q)[None, :] * s_q).reshape(-1)] = r.reshape(-1)
        elif op in (CH_CP_SCALE, CH_CP_FACTOR_GRAD):
            d, us, ks = a[0], a[1:5], a[5:9]
            rest = a[9:]
            if op == CH_CP_SCALE:
                lam_p, scale_p, R, Mm = rest
            else:
                lam_p, dscale_p, out_p, which, R, Mm = rest
            assert Mm == M and list(ks[:d]) == kept
            lam = t(mem.view(lam_p, R, c64).copy()).requires_grad_(True)
            facs = [t(mem.view(us[j], ks[j] * R, c64).reshape(ks[j], R).copy()).requires_grad_(True) for j in range(d)]
            s = lam.reshape(R, *[1] * d)
            for j, u in enumerate(facs):
                shp = [R] + [1] * d
                shp[1 + j] = ks[j]
                s = s * u.t().reshape(shp)                           # scale[e, m] = lambda[e] prod_j U_j[m_j, e]
            if op == CH_CP_SCALE:
                mem.view(scale_p, R * M, c64)[:] = s.detach().numpy().reshape(-1)
            else:
                ds = t(mem.view(dscale_p, R * M, c64).reshape(s.shape).copy())
                s.backward(ds)
                g = lam.grad if which < 0 else facs[which].grad
                mem.view(out_p, g.numel(), c64)[:] = g.numpy().reshape(-1)
        elif op == CH_CP_APPLY:
            src, scale, dst, conj, batch, per = a
            sc_ = mem.view(scale, per, c64)
            mem.view(dst, batch * per, c64)[:] = (mem.view(src, batch * per, c64).reshape(batch, per) * (sc_.conj() if conj else sc_)[None]).reshape(-1)
        elif op == CH_CP_DSCALE:
            tt_, g, ds, batch, per = a
            mem.view(ds, per, c64)[:] = (mem.view(tt_, batch * per, c64).reshape(batch, per).conj() * mem.view(g, batch * per, c64).reshape(batch, per)).sum(0)
        elif op == CH_BIAS_GRAD:
            gm, db, batch, Co, n_modes, dc_slot, inv_scale_bits = a
            inv_scale = np.frombuffer(np.int64(inv_scale_bits).tobytes()[:4], dtype=np.float32)[0]
            g = mem.view(gm, batch * Co * n_modes, c64).reshape(batch, Co, n_modes)
            mem.view(db, Co, f32)[:] = g[:, :, dc_slot].real.sum(0) * inv_scale
        elif op == CH_CONTRACT_FWD:
            xm, w, ym, Bn, Cin, Cout = a
            x = mem.view(xm, Bn * Cin * M, c64).reshape(Bn, Cin, M)
            wm = mem.view(w, Cin * Cout * M, c64).reshape(Cin, Cout, M)
            mem.view(ym, Bn * Cout * M, c64)[:] = np.einsum("bim,iom->bom", x, wm).reshape(-1)
        elif op == CH_CONTRACT_BWD:
            xm, gm, w, dxm, dw, Bn, Cin, Cout = a
            x = mem.view(xm, Bn * Cin * M, c64).reshape(Bn, Cin, M)
            g = mem.view(gm, Bn * Cout * M, c64).reshape(Bn, Cout, M)
            wm = mem.view(w, Cin * Cout * M, c64).reshape(Cin, Cout, M)
            mem.view(dxm, Bn * Cin * M, c64)[:] = np.einsum("bom,iom->bim", g, wm.conj()).reshape(-1)
            mem.view(dw, Cin * Cout * M, c64)[:] = np.einsum("bim,bom->iom", x.conj(), g).reshape(-1)
        else:
            raise AssertionError(f"unknown opcode {op}")


def _c(*shape, g=None):
    return torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g))


def rel(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


CASES = [((24,), (10,)), ((12, 10), (6, 8)), ((8, 6, 10), (4, 4, 6)), ((6, 6, 4, 6), (4, 2, 4, 4)), ((9, 11), (4, 5))]


@pytest.mark.parametrize("grid,modes", CASES)
def test_cp_chain_in_c_replayed_on_the_host(grid, modes):
    B, Ci, Co, R = 2, 3, 4, 5
    g = torch.Generator().manual_seed(len(grid))
    stored = O.stored_n_modes(modes)
    plans = O.kept_mode_plan(list(grid), stored, stored)
    kept = [p.kept for p in plans]
    d = len(grid)
    x, gy = torch.randn(B, Ci, *grid, generator=g), torch.randn(B, Co, *grid, generator=g)
    lam, u_in, u_out = _c(R, g=g), _c(Ci, R, g=g), _c(Co, R, g=g)
    u_modes = [_c(k, R, g=g) for k in kept]
    bias = torch.randn(Co, *[1] * d, generator=g)
    w = O.Weight("cp", weights=lam, factors=[u_in, u_out, *u_modes])
    y_ref, dx_ref, dws, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes, max_n_modes=stored)

    mem = Memory()
    for region, val in [(R_X, x), (R_LAM, lam), (R_UIN, u_in), (R_UOUT, u_out), (R_BIAS, bias.reshape(-1))] + [(R_MODE0 + j, u) for j, u in enumerate(u_modes)]:
        mem.put(region, val.numpy())
    replay(chain_log(grid, modes, 0, 0, B, Ci, Co, [R]), mem, plans, list(grid), B)
    assert rel(mem.get(R_Y, (B, Co, *grid), np.float32), y_ref) < 2e-5
    mem.put(R_X, gy.numpy())                                            # region 1 carries gy on the way back, region 2 receives dx
    replay(chain_log(grid, modes, 0, 1, B, Ci, Co, [R]), mem, plans, list(grid), B)
    assert rel(mem.get(R_Y, (B, Ci, *grid), np.float32), dx_ref) < 3e-5
    got = [mem.get(R_DLAM, (R,), np.complex64), mem.get(R_DUIN, (Ci, R), np.complex64), mem.get(R_DUOUT, (Co, R), np.complex64)] + \
          [mem.get(R_DMODE0 + j, (kept[j], R), np.complex64) for j in range(d)]
    for i, (a, b) in enumerate(zip(got, dws)):
        assert rel(a, b) < 5e-5, f"factor gradient {i}"
    assert rel(mem.get(R_BIAS, (Co,), np.float32), db_ref.reshape(-1)) < 3e-5


@pytest.mark.parametrize("grid,modes", CASES)
def test_tt_chain_in_c_replayed_on_the_host(grid, modes):
    B, Ci, Co = 2, 3, 4
    g = torch.Generator().manual_seed(10 + len(grid))
    stored = O.stored_n_modes(modes)
    plans = O.kept_mode_plan(list(grid), stored, stored)
    kept = [p.kept for p in plans]
    d = len(grid)
    r1 = 3
    rk = [4] + [2 + j for j in range(d - 1)] + [1]                      # r_0 .. r_{d-1}, r_d = 1
    x, gy = torch.randn(B, Ci, *grid, generator=g), torch.randn(B, Co, *grid, generator=g)
    g0, g1 = _c(1, Ci, r1, g=g), _c(r1, Co, rk[0], g=g)
    cores = [_c(rk[j], kept[j], rk[j + 1], g=g) for j in range(d)]
    bias = torch.randn(Co, *[1] * d, generator=g)
    w = O.Weight("tt", factors=[g0, g1, *cores])
    y_ref, dx_ref, dws, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes, max_n_modes=stored)

    ranks = [r1] + rk[:d]
    mem = Memory()
    for region, val in [(R_X, x), (R_UIN, g0), (R_UOUT, g1), (R_BIAS, bias.reshape(-1))] + [(R_MODE0 + j, c) for j, c in enumerate(cores)]:
        mem.put(region, val.numpy())
    replay(chain_log(grid, modes, 1, 0, B, Ci, Co, ranks), mem, plans, list(grid), B)
    assert rel(mem.get(R_Y, (B, Co, *grid), np.float32), y_ref) < 2e-5
    mem.put(R_X, gy.numpy())
    replay(chain_log(grid, modes, 1, 1, B, Ci, Co, ranks), mem, plans, list(grid), B)
    assert rel(mem.get(R_Y, (B, Ci, *grid), np.float32), dx_ref) < 3e-5
    got = [mem.get(R_DUIN, (1, Ci, r1), np.complex64), mem.get(R_DUOUT, (r1, Co, rk[0]), np.complex64)] + \
          [mem.get(R_DMODE0 + j, (rk[j], kept[j], rk[j + 1]), np.complex64) for j in range(d)]
    for i, (a, b) in enumerate(zip(got, dws)):
        assert rel(a, b) < 5e-5, f"core gradient {i}"
    assert rel(mem.get(R_BIAS, (Co,), np.float32), db_ref.reshape(-1)) < 3e-5
