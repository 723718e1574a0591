"""Mint golden vectors for the Fourier layer from the UNMODIFIED reference `FNOBlocks`.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden_block.py            # rewrites tests/golden/block_*.npz + block_index.json

Each file holds seeded inputs (x, every parameter of the block, the upstream gradient) and what
`neuralop.layers.fno_block.FNOBlocks.forward(x, index, output_shape)` -- with the reference's own SpectralConv inside, imported
through `oracle/load_reference.py` -- returns for them: y, and the autograd gradients dx and d(parameter) for every parameter the
layer touches.  Parameters are stored under the reference's state-dict names.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.load_reference import load_reference_spectral_conv  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, B, Ci, Co, grid, n_modes, n_layers, index, ctor kwargs, forward kwargs
CASES = [
    ("block_d2_default_mid", 2, 8, 8, (16, 16), (8, 8), 2, 0, {}, {}),
    ("block_d2_default_last", 2, 8, 8, (16, 16), (8, 8), 2, 1, {}, {}),
    ("block_d1_single", 3, 6, 6, (40,), (12,), 1, 0, {}, {}),
    ("block_d1_mid", 2, 6, 6, (33,), (10,), 3, 1, {}, {}),
    ("block_d3_mid", 1, 4, 4, (8, 10, 12), (4, 4, 6), 2, 0, {}, {}),
    ("block_d2_odd_grid_linear_skips", 2, 6, 10, (15, 17), (6, 8), 2, 0, {"channel_mlp_skip": "linear"}, {}),
    ("block_d2_channels_change_last", 2, 5, 9, (12, 12), (6, 6), 1, 0, {"channel_mlp_skip": "linear"}, {}),
    ("block_d2_softgating_fno_skip", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"fno_skip": "soft-gating"}, {}),
    ("block_d2_identity_skips", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"fno_skip": "identity", "channel_mlp_skip": "identity"}, {}),
    ("block_d2_no_skips", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"fno_skip": None, "channel_mlp_skip": None}, {}),
    ("block_d2_no_mlp_mid", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"use_channel_mlp": False}, {}),
    ("block_d2_no_mlp_last", 2, 6, 6, (12, 14), (6, 6), 2, 1, {"use_channel_mlp": False}, {}),
    ("block_d2_tanh", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"stabilizer": "tanh"}, {}),
    ("block_d2_preactivation_mid", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"preactivation": True}, {}),
    ("block_d2_preactivation_last", 2, 6, 6, (12, 14), (6, 6), 2, 1, {"preactivation": True}, {}),
    ("block_d2_expansion_2", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"channel_mlp_expansion": 2.0}, {}),
    ("block_d2_upsample", 2, 4, 4, (10, 12), (6, 6), 2, 0, {"resolution_scaling_factor": 2}, {}),
    ("block_d1_downsample", 2, 4, 4, (32,), (8,), 2, 0, {"resolution_scaling_factor": 0.5}, {}),
    ("block_d2_output_shape", 2, 4, 4, (10, 12), (6, 6), 2, 0, {}, {"output_shape": (14, 9)}),
    ("block_d3_upsample", 1, 3, 3, (6, 6, 8), (4, 4, 4), 2, 0, {"resolution_scaling_factor": 2}, {}),
    ("block_d2_max_modes", 2, 6, 6, (16, 16), (6, 6), 2, 0, {"max_n_modes": (8, 8)}, {}),
    ("block_d2_tucker", 2, 6, 6, (16, 12), (8, 6), 2, 0,
     {"factorization": "tucker", "implementation": "factorized", "rank": [4, 3, 5, 3]}, {}),
    ("block_d2_cp_reconstructed", 2, 6, 6, (16, 12), (8, 6), 2, 0,
     {"factorization": "cp", "implementation": "reconstructed", "rank": 7}, {}),
    ("block_d2_darcy_small", 2, 16, 16, (32, 32), (16, 16), 4, 2, {}, {}),
    # appended (seeds follow the position in this list: keep the order)
    ("block_d2_instance_norm_mid", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"norm": "instance_norm"}, {}),
    ("block_d2_instance_norm_last", 2, 6, 6, (12, 14), (6, 6), 2, 1, {"norm": "instance_norm"}, {}),
    ("block_d2_group_norm_mid", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"norm": "group_norm"}, {}),
    ("block_d2_group_norm_3_groups", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"norm": "group_norm", "norm_groups": 3}, {}),
    ("block_d1_group_norm_preactivation", 2, 4, 4, (30,), (8,), 2, 0, {"norm": "group_norm", "norm_groups": 2, "preactivation": True}, {}),
    ("block_d3_instance_norm_no_mlp", 1, 4, 4, (6, 8, 10), (4, 4, 4), 2, 0, {"norm": "instance_norm", "use_channel_mlp": False}, {}),
    ("block_d2_group_norm_upsample_softgating", 2, 4, 4, (10, 12), (6, 6), 2, 0,
     {"norm": "group_norm", "resolution_scaling_factor": 2, "fno_skip": "soft-gating"}, {}),
    ("block_d2_batch_norm_mid", 3, 6, 6, (12, 14), (6, 6), 2, 0, {"norm": "batch_norm"}, {}),
    ("block_d1_batch_norm_last", 3, 4, 4, (30,), (8,), 2, 1, {"norm": "batch_norm"}, {}),
    ("block_d2_ada_in_mid", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"norm": "ada_in", "ada_in_features": 5}, {}),
    ("block_cplx_d2_mid", 2, 4, 4, (10, 12), (6, 6), 2, 0, {"complex_data": True}, {}),
    ("block_cplx_d2_last", 2, 4, 4, (10, 12), (6, 6), 2, 1, {"complex_data": True}, {}),
    ("block_cplx_d1_tanh_linear_skips", 2, 4, 6, (24,), (8,), 2, 0, {"complex_data": True, "stabilizer": "tanh", "channel_mlp_skip": "linear"}, {}),
    ("block_cplx_d2_identity_no_mlp", 2, 4, 4, (10, 12), (6, 6), 2, 0, {"complex_data": True, "fno_skip": "identity", "use_channel_mlp": False}, {}),
    ("block_cplx_d3_preactivation_softgating", 1, 3, 3, (6, 6, 8), (4, 4, 4), 2, 0,
     {"complex_data": True, "preactivation": True, "fno_skip": "soft-gating"}, {}),
    ("block_d2_relu", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"non_linearity": "relu"}, {}),
    ("block_d2_silu_preactivation", 2, 6, 6, (12, 14), (6, 6), 2, 0, {"non_linearity": "silu", "preactivation": True}, {}),
    ("block_d1_tanh_activation_no_mlp", 2, 4, 4, (30,), (8,), 2, 0, {"non_linearity": "tanh", "use_channel_mlp": False}, {}),
    ("block_d2_conv_bias_kernel_3", 2, 4, 6, (12, 14), (6, 6), 2, 0, {"conv_bias_kernel": 3, "channel_mlp_skip": "linear"}, {}),
    ("block_d1_conv_bias_kernel_5_upsample", 2, 4, 4, (24,), (8,), 2, 0, {"conv_bias_kernel": 5, "resolution_scaling_factor": 2}, {}),
]

ACTIVATIONS = {"relu": torch.nn.functional.relu, "silu": torch.nn.functional.silu, "tanh": torch.tanh}


def main():
    os.makedirs(OUT, exist_ok=True)
    load_reference_spectral_conv()
    fb = importlib.import_module("neuralop.layers.fno_block")
    index = {}
    for seed, (name, B, Ci, Co, grid, modes, n_layers, idx, ckw, fkw) in enumerate(CASES):
        torch.manual_seed(7000 + seed)
        ctor = dict(implementation="reconstructed")
        ctor.update(ckw)
        live = dict(ctor)
        if "non_linearity" in live:
            live["non_linearity"] = ACTIVATIONS[live["non_linearity"]]
        blk = fb.FNOBlocks(Ci, Co, modes, n_layers=n_layers, **live)
        with torch.no_grad():                                   # soft-gating weights start at exactly 1: make them matter
            for pname, p in blk.named_parameters():
                if "channel_mlp_skips" in pname or (pname.startswith("fno_skips") and p.ndim == len(grid) + 2) or pname.startswith("norm."):
                    p.add_(0.3 * torch.randn_like(p))
        embedding = None
        if ctor.get("norm") == "ada_in":
            embedding = torch.randn(ctor["ada_in_features"])
            blk.set_ada_in_embeddings(embedding)
        x = torch.randn(B, Ci, *grid, dtype=torch.cfloat if ctor.get("complex_data") else torch.float32, requires_grad=True)
        y = blk(x, idx, **fkw)
        gy = torch.randn_like(y)
        y.backward(gy)
        arrays = {}
        for key, val in (("x", x.detach()), ("gy", gy), ("y", y.detach()), ("dx", x.grad)):
            arrays[key + ("__c" if val.is_complex() else "")] = torch.view_as_real(val).numpy() if val.is_complex() else val.numpy()
        if embedding is not None:
            arrays["ada_in_embedding"] = embedding.numpy()
        for bname, buf in blk.named_buffers():                   # batch norm: the running statistics AFTER this forward
            arrays["b__" + bname.replace(".", "__")] = buf.detach().numpy()
        pnames, touched = [], []
        for pname, p in blk.named_parameters():
            key = pname.replace(".", "__")
            val = p.detach()
            arrays["p__" + key + ("__c" if val.is_complex() else "")] = torch.view_as_real(val).numpy() if val.is_complex() else val.numpy()
            pnames.append(pname)
            if p.grad is not None:
                gval = p.grad
                arrays["g__" + key + ("__c" if gval.is_complex() else "")] = torch.view_as_real(gval).numpy() if gval.is_complex() else gval.numpy()
                touched.append(pname)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        index[name] = {
            "B": B, "in_channels": Ci, "out_channels": Co, "grid": list(grid), "n_modes": list(modes), "n_layers": n_layers,
            "index": idx, "ctor": {k: (list(v) if isinstance(v, tuple) else v) for k, v in ctor.items()},
            "forward": {k: list(v) for k, v in fkw.items()}, "params": pnames, "touched": touched,
            "weight_kind": blk.convs[0].weight.name.lower().replace("complex", ""), "out_grid": list(y.shape[2:]),
        }
        print(f"{name:34s} y{tuple(y.shape)} touched={len(touched)}/{len(pnames)}")
    with open(os.path.join(OUT, "block_index.json"), "w") as f:
        json.dump({"reference": "neuraloperator@93d3f06 neuralop/layers/fno_block.py (FNOBlocks, with the reference SpectralConv inside)",
                   "generator": "oracle/make_golden_block.py", "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
