"""Tensor-core mode-wise contraction (`k_mode_gemm_quad2`) against float64 einsums, over extents on both sides of the
64-row / 64-column tile and of the 8-k chunk / 32-k slab granularity, and the two kernel generations against each other."""
import pytest
import torch

import neuraloperator_b200 as nb
from oracle import spectral_conv_oracle as O

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def rel_err(a, ref):
    a, ref = a.detach().double(), ref.detach().double()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,Ci,Co,grid,modes", [
    (32, 64, 64, (128, 128), (32, 32)),     # headline extents: 136 quads, K = 64 / 32
    (96, 64, 64, (64, 64), (16, 16)),       # batch > 64: two column tiles (forward / dxm), three K slabs (dweight)
    (8, 128, 128, (64, 64), (16, 16)),      # hidden_channels 128: two row tiles, four K slabs, ring recycling
    (3, 5, 7, (32, 32), (8, 8)),            # nothing is a multiple of anything: K tails, row / column padding
    (4, 40, 24, (32, 32), (8, 8)),          # K = 40: one full slab + one chunk
    (70, 72, 130, (16, 16), (8, 8)),        # every extent above 64, none a multiple of 8
    (16, 32, 32, (1024,), (16,)),           # 1-D: 9 modes -> not a multiple of 4: single-mode kernel
    (2, 16, 16, (16, 16, 16), (8, 8, 8)),   # 3-D mode block
])
def test_contraction_matches_float64(cuda_device, B, Ci, Co, grid, modes):
    dev = cuda_device
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(dev, grid, grid, stored, stored)
    torch.manual_seed(B * 1000 + Ci)
    cplx = dict(dtype=torch.complex64, device=dev)
    xm = torch.randn(B, Ci, *plan.kept, **cplx)
    gm = torch.randn(B, Co, *plan.kept, **cplx)
    w = torch.randn(Ci, Co, *plan.kept, **cplx)
    ym = nb.contract_dense(plan, xm, w)
    dxm, dw, db = nb.contract_dense_backward(plan, xm, gm, w)
    torch.cuda.synchronize()
    X, G, Wd = xm.to(torch.complex128), gm.to(torch.complex128), w.to(torch.complex128)
    flat = lambda t: t.reshape(*t.shape[:2], -1)
    ym_ref = torch.einsum("bim,iom->bom", flat(X), flat(Wd))
    dxm_ref = torch.einsum("bom,iom->bim", flat(G), flat(Wd).conj())
    dw_ref = torch.einsum("bim,bom->iom", flat(X).conj(), flat(G))
    assert rel_err(torch.view_as_real(flat(ym)), torch.view_as_real(ym_ref)) < REL_TOL, "ym"
    assert rel_err(torch.view_as_real(flat(dxm)), torch.view_as_real(dxm_ref)) < REL_TOL, "dxm"
    assert rel_err(torch.view_as_real(flat(dw)), torch.view_as_real(dw_ref)) < REL_TOL, "dweight"
    # dbias: sum over the batch of the real part of the all-zero-frequency slot of gm, undoing the synthesis scale (1 here: "forward" norm)
    dc = [k // 2 for k in plan.kept[:-1]] + [0]
    db_ref = gm[(slice(None), slice(None), *dc)].real.double().sum(0)
    assert rel_err(db, db_ref) < REL_TOL, "dbias"


def test_contraction_is_deterministic(cuda_device):
    dev = cuda_device
    plan = nb.get_plan(dev, (128, 128), (128, 128), [32, 17], [32, 17])
    torch.manual_seed(0)
    xm = torch.randn(32, 64, 32, 17, dtype=torch.complex64, device=dev)
    w = torch.randn(64, 64, 32, 17, dtype=torch.complex64, device=dev)
    a = nb.contract_dense(plan, xm, w)
    b = nb.contract_dense(plan, xm, w)
    assert torch.equal(torch.view_as_real(a), torch.view_as_real(b))
