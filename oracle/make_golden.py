"""Mint golden vectors for the SpectralConv path from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Each file holds seeded inputs (x, weight tensors, bias, upstream grad) and what the reference class
`neuralop.layers.spectral_convolution.SpectralConv` (imported through `oracle/load_reference.py`) returns
for them: y, and the autograd gradients dx, d(weight tensors), dbias.  The reference has no golden
vectors of its own, so these are the pins for both the oracle and the CUDA path.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.load_reference import load_reference_spectral_conv  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, B, Ci, Co, grid, n_modes, extra ctor kwargs, forward kwargs
CASES = [
    ("d1_even", 2, 3, 4, (16,), (8,), {}, {}),
    ("d1_burgers_small", 2, 4, 4, (64,), (16,), {}, {}),
    ("d2_even", 2, 3, 4, (16, 12), (8, 6), {}, {}),
    ("d2_odd_grid_odd_modes", 2, 3, 4, (9, 9), (4, 5), {}, {}),
    ("d2_modes_exceed_grid", 2, 3, 2, (8, 8), (12, 12), {}, {}),
    ("d2_all_modes", 1, 2, 2, (8, 8), (8, 8), {}, {}),
    ("d2_darcy_small", 2, 8, 8, (32, 32), (16, 16), {}, {}),
    ("d3_mixed", 2, 3, 4, (8, 9, 10), (4, 5, 7), {}, {}),
    ("d3_ns_small", 1, 4, 4, (16, 16, 16), (8, 8, 8), {}, {}),
    ("d4_small", 1, 2, 2, (6, 6, 6, 6), (4, 4, 4, 4), {}, {}),
    ("d2_max_modes_even_start", 2, 3, 4, (16, 12), (6, 4), {"max_n_modes": (8, 6)}, {}),
    ("d2_max_modes_odd_start", 2, 3, 4, (16, 12), (5, 4), {"max_n_modes": (8, 6)}, {}),
    ("d2_downsample", 2, 3, 4, (12, 12), (10, 8), {"resolution_scaling_factor": 0.5}, {}),
    ("d2_upsample", 2, 3, 4, (12, 12), (10, 8), {"resolution_scaling_factor": 2}, {}),
    ("d2_upsample_all_modes", 1, 2, 3, (12, 12), (12, 12), {"resolution_scaling_factor": 2}, {}),
    ("d2_output_shape", 2, 3, 4, (12, 13), (6, 6), {}, {"output_shape": (9, 16)}),
    ("d2_norm_ortho", 2, 3, 4, (12, 12), (10, 8), {"fft_norm": "ortho"}, {}),
    ("d2_norm_backward", 2, 3, 4, (12, 12), (10, 8), {"fft_norm": "backward"}, {}),
    ("d2_no_bias", 2, 3, 4, (12, 12), (6, 6), {"bias": False}, {}),
    ("d2_tucker", 2, 6, 5, (16, 12), (8, 6),
     {"factorization": "tucker", "implementation": "factorized", "rank": [4, 3, 5, 3]}, {}),
    ("d2_tucker_reconstructed", 2, 6, 5, (16, 12), (8, 6),
     {"factorization": "tucker", "implementation": "reconstructed", "rank": [4, 3, 5, 3]}, {}),
    ("d1_tucker", 2, 6, 5, (32,), (12,),
     {"factorization": "tucker", "implementation": "factorized", "rank": [4, 3, 5]}, {}),
    ("d3_tucker", 1, 4, 4, (8, 8, 8), (4, 4, 4),
     {"factorization": "tucker", "implementation": "factorized", "rank": [3, 3, 3, 3, 2]}, {}),
    ("d2_cp", 2, 6, 5, (16, 12), (8, 6),
     {"factorization": "cp", "implementation": "factorized", "rank": 7}, {}),
    ("d2_tt", 2, 6, 5, (16, 12), (8, 6),
     {"factorization": "tt", "implementation": "factorized", "rank": [1, 3, 4, 3, 1]}, {}),
    ("d2_separable", 2, 5, 5, (16, 12), (8, 6), {"separable": True}, {}),
    ("d2_separable_max_modes", 2, 4, 4, (16, 12), (5, 4), {"separable": True, "max_n_modes": (8, 6)}, {}),
    ("d1_separable", 3, 6, 6, (32,), (12,), {"separable": True, "implementation": "factorized"}, {}),
    ("d3_separable", 1, 3, 3, (8, 8, 8), (4, 4, 4), {"separable": True}, {}),
    ("d2_separable_tucker", 2, 5, 5, (16, 12), (8, 6),
     {"separable": True, "factorization": "tucker", "implementation": "factorized", "rank": [3, 5, 3]}, {}),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference_spectral_conv()
    index = {}
    for seed, (name, B, Ci, Co, grid, modes, ckw, fkw) in enumerate(CASES):
        torch.manual_seed(1000 + seed)
        conv = ref.SpectralConv(Ci, Co, modes, **ckw)
        x = torch.randn(B, Ci, *grid, requires_grad=True)
        y = conv(x, **fkw)
        gy = torch.randn_like(y)
        y.backward(gy)
        arrays = {"x": x.detach().numpy(), "gy": gy.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
        pnames = []
        for pname, p in conv.named_parameters():
            key = pname.replace(".", "__")
            arrays["p__" + key] = p.detach().numpy()
            arrays["g__" + key] = p.grad.numpy()
            pnames.append(pname)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        index[name] = {
            "B": B, "in_channels": Ci, "out_channels": Co, "grid": list(grid), "n_modes": list(modes),
            "ctor": {k: (list(v) if isinstance(v, tuple) else v) for k, v in ckw.items()},
            "forward": {k: list(v) for k, v in fkw.items()},
            "stored_n_modes": list(conv.n_modes), "max_n_modes": list(conv.max_n_modes),
            "params": pnames, "weight_kind": conv.weight.name.lower(),
            "out_grid": list(y.shape[2:]),
        }
        print(f"{name:28s} y{tuple(y.shape)} params={pnames}")
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump({"reference": "neuraloperator@93d3f06 neuralop/layers/spectral_convolution.py",
                   "generator": "oracle/make_golden.py", "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
