"""The Fourier-layer oracle (oracle/fno_block_oracle.py) against golden vectors minted from the UNMODIFIED reference `FNOBlocks`
(oracle/make_golden_block.py) and, where /root/reference exists (the build container), against the live class: forward, dx and the
gradient of every parameter the layer touches.  CPU only."""
import importlib

import pytest
import torch

from conftest import block_golden_index, block_oracle_kwargs, load_block_golden
from oracle import fno_block_oracle as BO
from oracle.load_reference import load_reference_spectral_conv, reference_available

CASES = sorted(block_golden_index().keys())


def rel_err(a, ref):
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return (a - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)


@pytest.mark.parametrize("name", CASES)
def test_block_oracle_matches_golden(name):
    meta, io, params, grads = load_block_golden(name)
    if "ada_in_embedding" in io:
        params = dict(params, ada_in_embedding=io["ada_in_embedding"])
    y, dx, g = BO.fno_block_fwd_bwd(io["x"], params, meta["index"], io["gy"], **block_oracle_kwargs(meta))
    g.pop("ada_in_embedding", None)
    assert list(y.shape[2:]) == meta["out_grid"]
    assert rel_err(y, io["y"]) < 2e-5, "y"
    assert rel_err(dx, io["dx"]) < 2e-5, "dx"
    assert sorted(g.keys()) == sorted(meta["touched"])         # the layer touches exactly the parameters the reference's does
    for pname in meta["touched"]:
        assert rel_err(g[pname], grads[pname]) < 2e-5, pname


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["block_d2_default_mid", "block_d2_default_last", "block_d3_mid", "block_d2_tanh",
                                  "block_d2_preactivation_mid", "block_d2_upsample", "block_d2_no_mlp_mid", "block_d2_tucker"])
def test_block_oracle_matches_live_reference(name):
    """Fresh random parameters and inputs (not the stored ones) through the live reference class and the restatement."""
    meta, io, _, _ = load_block_golden(name)
    load_reference_spectral_conv()
    fb = importlib.import_module("neuralop.layers.fno_block")
    torch.manual_seed(99)
    blk = fb.FNOBlocks(meta["in_channels"], meta["out_channels"], tuple(meta["n_modes"]), n_layers=meta["n_layers"], **meta["ctor"])
    with torch.no_grad():
        for pname, p in blk.named_parameters():
            if "skips" in pname:
                p.add_(0.2 * torch.randn_like(p))
    x = torch.randn_like(io["x"]).requires_grad_(True)
    kw = {k: tuple(v) for k, v in meta["forward"].items()}
    y = blk(x, meta["index"], **kw)
    gy = torch.randn_like(y)
    y.backward(gy)
    params = {k: v.detach() for k, v in blk.named_parameters()}
    y2, dx2, g2 = BO.fno_block_fwd_bwd(x.detach(), params, meta["index"], gy, **block_oracle_kwargs(meta))
    assert rel_err(y2, y.detach()) < 1e-6
    assert rel_err(dx2, x.grad) < 1e-6
    for pname, p in blk.named_parameters():
        if p.grad is not None:
            assert rel_err(g2[pname], p.grad) < 1e-6, pname
        else:
            assert pname not in g2
