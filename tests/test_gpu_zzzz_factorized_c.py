"""CP / TT contracted factor by factor behind ONE C call per direction (sc_forward_cp / sc_backward_cp, sc_forward_tt /
sc_backward_tt; opt-in through `spectral_conv.FACTORIZED_CHAINS_IN_C`) against the golden vectors minted from the unmodified
reference and against the Python-orchestrated chains (the default, validated on hardware in round 2) on the same parameters.
(Named zzzz: written after the round's GPU minutes were spent, so it runs after every other tier.)"""
import pytest
import torch

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib, spectral_conv as sc
from conftest import golden_grads, load_golden
from test_gpu_parity import _module_from_golden, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]
REL_TOL = 1e-4


@pytest.fixture
def c_chains():
    old = sc.FACTORIZED_CHAINS_IN_C
    sc.FACTORIZED_CHAINS_IN_C = True
    yield
    sc.FACTORIZED_CHAINS_IN_C = old


@pytest.mark.parametrize("name", ["d2_cp", "d2_tt"])
def test_c_chain_matches_reference_golden(cuda_device, c_chains, name):
    meta, arr = load_golden(name)
    conv = _module_from_golden(meta, arr, cuda_device)
    x = arr["x"].to(cuda_device).requires_grad_(True)
    before = _lib.launch_count()
    y = conv(x)
    y.backward(arr["gy"].to(cuda_device))
    torch.cuda.synchronize()
    assert _lib.launch_count() > before
    assert rel_err(y, arr["y"]) < REL_TOL, "y"
    assert rel_err(x.grad, arr["dx"]) < REL_TOL, "dx"
    gws, gb = golden_grads(meta, arr)
    for i, (p, g) in enumerate(zip(conv.weight.decomposition(), gws)):
        assert p.grad is not None, f"param {i} received no grad"
        assert rel_err(p.grad, g) < REL_TOL, f"dparam{i}"
    assert rel_err(conv.bias.grad, gb) < REL_TOL, "dbias"


@pytest.mark.parametrize("fact,rank,Ci,Co,grid,modes,max_modes", [
    ("cp", 6, 4, 5, (12, 10, 16), (6, 6, 8), None),
    ("cp", 8, 8, 8, (32, 32), (8, 8), (16, 12)),
    ("cp", 5, 6, 6, (64,), (20,), None),
    ("tt", 0.3, 12, 10, (40, 36), (16, 12), None),
    ("tt", 0.5, 6, 7, (64,), (20,), None),
    ("tt", 0.4, 4, 5, (12, 10, 16), (6, 6, 8), None),
    ("tt", 0.3, 8, 8, (32, 32), (8, 8), (16, 12)),
    ("tt", 0.5, 3, 3, (6, 6, 6, 6), (4, 4, 4, 4), None),
])
def test_c_chain_matches_python_chain(cuda_device, fact, rank, Ci, Co, grid, modes, max_modes):
    dev = cuda_device
    torch.manual_seed(13)
    conv = nb.SpectralConv(Ci, Co, modes, implementation="factorized", factorization=fact, rank=rank, max_n_modes=max_modes).to(dev)
    with torch.no_grad():
        for p in conv.weight.decomposition():
            p.copy_(torch.randn_like(p))
        conv.bias.normal_()
    B = 3
    x = torch.randn(B, Ci, *grid, device=dev)
    gy = torch.randn(B, Co, *grid, device=dev)
    res = {}
    for in_c in (False, True):
        sc.FACTORIZED_CHAINS_IN_C = in_c
        try:
            for p in conv.parameters():
                p.grad = None
            xx = x.clone().requires_grad_(True)
            y = conv(xx)
            y.backward(gy)
            torch.cuda.synchronize()
            res[in_c] = (y.detach(), xx.grad, conv.bias.grad.clone(), [p.grad.clone() for p in conv.weight.decomposition()])
        finally:
            sc.FACTORIZED_CHAINS_IN_C = False
    (ya, dxa, dba, dpa), (yb, dxb, dbb, dpb) = res[False], res[True]
    assert rel_err(yb, ya) < 1e-5 and rel_err(dxb, dxa) < 1e-5 and rel_err(dbb, dba) < 1e-5     # same kernels, same operands
    for i, (a, b) in enumerate(zip(dpa, dpb)):
        assert rel_err(b, a) < 1e-5, f"dparam{i}"
