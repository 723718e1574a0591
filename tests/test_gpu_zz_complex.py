"""complex_data=True on the GPU: the nn.Module (C2C chain on the complex table kernel + the dense mode GEMM) against golden vectors
minted from the unmodified reference (oracle/make_golden_complex.py): y, dx, every parameter gradient, dbias.
(Named zz so that it runs after the real-data tiers.)"""
import pytest
import torch

import neuraloperator_b200 as nb
from conftest import complex_golden_index, load_complex_golden

# This tier was written after the round's GPU minutes were spent: a hard per-test limit (stack dump + process exit) keeps a
# fault here from holding the GPU box.
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]
REL_TOL = 1e-4
CASES = sorted(complex_golden_index().keys())


def rel_err(a, ref):
    a, ref = a.detach().cpu(), ref.detach().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return (a - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)


@pytest.mark.parametrize("name", CASES)
def test_complex_module_matches_golden(cuda_device, name):
    meta, arr = load_complex_golden(name)
    ctor = dict(meta["ctor"])
    conv = nb.SpectralConv(meta["in_channels"], meta["out_channels"], tuple(meta["n_modes"]), complex_data=True, **ctor).to(cuda_device)
    assert conv.n_modes == meta["stored_n_modes"]
    params = dict(conv.named_parameters())
    with torch.no_grad():
        for pname in meta["params"]:
            key = "p__" + pname.replace(".", "__")
            ours = pname.replace("weight.factors.", "weight.factors.factor_")
            assert params[ours].shape == arr[key].shape, (pname, params[ours].shape, arr[key].shape)
            params[ours].copy_(arr[key].to(cuda_device))
    x = arr["x"].to(cuda_device).requires_grad_(True)
    kw = {"output_shape": tuple(meta["forward"]["output_shape"])} if "output_shape" in meta["forward"] else {}
    y = conv(x, **kw)
    assert y.dtype == torch.complex64 and list(y.shape[2:]) == meta["out_grid"]
    y.backward(arr["gy"].to(cuda_device))
    assert rel_err(y, arr["y"]) < REL_TOL, "y"
    assert rel_err(x.grad, arr["dx"]) < REL_TOL, "dx"
    for pname in meta["params"]:
        key = "g__" + pname.replace(".", "__")
        ours = pname.replace("weight.factors.", "weight.factors.factor_")
        assert params[ours].grad is not None, pname
        assert rel_err(params[ours].grad, arr[key]) < REL_TOL, pname
