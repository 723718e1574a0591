"""The tensor-core variant of the channel-mixing kernel (`k_channel_mix_tc`: tcgen05 bf16x3, activations as a tensor-memory A
operand) against float64 and against the exact-fp32 SIMT kernel.  OPT-IN code that has never run on hardware (written after the
round's GPU minutes were spent; it cannot be emulated on a CPU), hence the very last file of the GPU tier.  Its waits are the
library's bounded mbarrier waits: a protocol bug traps after 2 s instead of hanging the GPU."""
import pytest
import torch
import torch.nn.functional as F

import neuraloperator_b200 as nb
from neuraloperator_b200 import _lib

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120, method="thread")]


@pytest.fixture
def tensor_cores():
    nb.set_tensor_core_mixing(True)
    yield
    nb.set_tensor_core_mixing(False)


def rel_err(a, ref):
    wide = lambda t: t.detach().cpu().to(torch.complex128 if t.is_complex() else torch.float64)       # noqa: E731
    a, ref = wide(a), wide(ref)
    return (a - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)


@pytest.mark.parametrize("B,Ci,Co,grid", [(1, 64, 64, (128,)), (2, 64, 64, (32, 32)), (2, 64, 32, (50, 30)), (2, 32, 64, (33,)),
                                           (3, 5, 7, (129,)), (2, 128, 128, (24, 24)), (1, 200, 100, (40,)), (4, 64, 64, (128, 128))])
def test_tensor_core_mix_matches_float64_and_simt(cuda_device, tensor_cores, B, Ci, Co, grid):
    assert nb.uses_tensor_core_mixing()
    g = torch.Generator().manual_seed(Ci + Co)
    x = torch.randn(B, Ci, *grid, generator=g)
    w = torch.randn(Co, Ci, 1, generator=g) / Ci ** 0.5
    bias = torch.randn(Co, generator=g)
    add = torch.randn(B, Co, *grid, generator=g)
    gate = torch.randn(1, Co, *[1] * len(grid), generator=g)
    ref = F.gelu(F.conv1d(x.double().flatten(2), w.double(), bias.double()).view(B, Co, *grid) + add.double() + gate.double() * x.double()[:, :1])
    dev = [t.to(cuda_device) for t in (x, w, bias, add, gate)]
    gated = dev[0][:, :1].expand(B, Co, *grid).contiguous()
    with torch.no_grad():
        out_tc = nb.channel_mix(dev[0], dev[1], dev[2], dev[3], dev[4], gated, act=_lib.ACT_GELU)
        torch.cuda.synchronize()
        nb.set_tensor_core_mixing(False)
        out_simt = nb.channel_mix(dev[0], dev[1], dev[2], dev[3], dev[4], gated, act=_lib.ACT_GELU)
        torch.cuda.synchronize()
    assert rel_err(out_simt, ref) < 1e-5
    assert rel_err(out_tc, ref) < 1e-4              # bf16x3: ~1e-5 of max|ref|
    assert rel_err(out_tc, out_simt) < 1e-4


def test_tensor_core_mix_through_the_block(cuda_device, tensor_cores):
    """A whole layer with the mixing launches on tensor cores (forward + the input-gradient launches of backward) against the SIMT run."""
    torch.manual_seed(3)
    blk = nb.FNOBlocks(64, 64, (16, 16), n_layers=2, implementation="reconstructed").to(cuda_device)
    x = torch.randn(2, 64, 64, 64, device=cuda_device)
    gy = torch.randn(2, 64, 64, 64, device=cuda_device)
    res = {}
    for tc in (True, False):
        nb.set_tensor_core_mixing(tc)
        for p in blk.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        y = blk(xx, 0)
        y.backward(gy)
        torch.cuda.synchronize()
        res[tc] = (y.detach(), xx.grad, blk.fno_skips[0].conv.weight.grad.clone())
    for a, b in zip(res[True], res[False]):
        assert rel_err(a, b) < 1e-4
