"""Last-dim ("rows") tensor-core kernels for grids the fused 2-D kernels do not cover (cfg-5 sizes, 1-D, big 3-D planes)."""
import pytest
import torch

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib

pytestmark = pytest.mark.gpu


REL_TOL = 1e-4


def rel_err(a, ref):
    a, ref = a.detach().cpu().double(), ref.detach().cpu().double()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("grid,modes,n0,n1", [
    ((256, 256), (64, 64), 1, 2),        # cfg-5a shape: 2k = 66 -> N1 = 80, two K slabs in the synthesis
    ((1024,), (16,), 16, 32),            # cfg-1: 1-D, 16 slabs
    ((256, 64), (32, 32), 2, 2),         # tall image, one slab
    ((8, 16, 128), (4, 8, 48), 2, 2),    # 3-D with a wide last dim
    ((384, 192), (20, 12), 1, 2),        # non power-of-two multiples of 64
])
@pytest.mark.parametrize("adjoint", [False, True])
def test_rows_kernels_match_the_generic_chain(cuda_device, grid, modes, n0, n1, adjoint):
    from oracle import spectral_conv_oracle as O
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, stored)
    mask = plan.uses_fast_path()
    assert mask & (64 if adjoint else 16) and mask & (128 if adjoint else 32), mask
    torch.manual_seed(7)
    x = torch.randn(n0, n1, *grid, device=cuda_device)
    fast = nb.analyze(plan, x, adjoint=adjoint)
    plan.set_fast_path(False)
    try:
        slow = nb.analyze(plan, x, adjoint=adjoint)
    finally:
        plan.set_fast_path(True)
    assert rel_err(torch.view_as_real(fast), torch.view_as_real(slow)) < REL_TOL
    bias = None if adjoint else torch.randn(n1, device=cuda_device)
    m = torch.randn(n0, n1, *plan.kept, dtype=torch.complex64, device=cuda_device)
    yf = nb.synthesize(plan, m, bias, adjoint=adjoint)
    plan.set_fast_path(False)
    try:
        ys = nb.synthesize(plan, m, bias, adjoint=adjoint)
    finally:
        plan.set_fast_path(True)
    assert rel_err(yf, ys) < REL_TOL


def test_cfg5a_forward_backward_against_the_oracle(cuda_device):
    from oracle import spectral_conv_oracle as O
    B, C, grid, modes = 1, 4, (256, 256), (64, 64)
    x, w, bias, gy = O.make_inputs(B, C, C, grid, modes, seed=0)
    y_ref, dx_ref, dws_ref, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes)
    conv = nb.SpectralConv(C, C, modes).to(cuda_device)
    with torch.no_grad():
        conv.weight.tensor.copy_(w.tensor.to(cuda_device))
        conv.bias.copy_(bias.to(cuda_device))
    xd = x.to(cuda_device).requires_grad_(True)
    y = conv(xd)
    y.backward(gy.to(cuda_device))
    assert rel_err(y, y_ref) < REL_TOL and rel_err(xd.grad, dx_ref) < REL_TOL
    assert rel_err(torch.view_as_real(conv.weight.tensor.grad), torch.view_as_real(dws_ref[0])) < REL_TOL
    assert rel_err(conv.bias.grad, db_ref) < REL_TOL
