"""Mint golden vectors for `SpectralConv(complex_data=True)` from the UNMODIFIED reference (test infrastructure; run in the
build container: `python oracle/make_golden_complex.py`).  Separate from make_golden.py so that the real-data files and
their index stay untouched: writes tests/golden/cplx_*.npz and tests/golden/complex_index.json."""
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
    ("cplx_d1", 2, 3, 4, (16,), (6,), {}, {}),
    ("cplx_d2_even", 2, 3, 4, (16, 12), (8, 6), {}, {}),
    ("cplx_d2_odd", 2, 3, 4, (9, 11), (4, 5), {}, {}),
    ("cplx_d2_modes_exceed_grid", 2, 3, 2, (12, 12), (16, 16), {}, {}),
    ("cplx_d2_max_modes_even_start", 2, 3, 4, (16, 12), (6, 4), {"max_n_modes": (8, 6)}, {}),
    ("cplx_d2_max_modes_odd_start", 2, 3, 4, (16, 12), (5, 3), {"max_n_modes": (8, 6)}, {}),
    ("cplx_d2_upsample", 2, 3, 4, (12, 12), (10, 8), {"resolution_scaling_factor": 2}, {}),
    ("cplx_d2_downsample", 2, 3, 4, (12, 12), (10, 8), {"resolution_scaling_factor": 0.5}, {}),
    ("cplx_d2_output_shape", 2, 3, 4, (12, 13), (6, 6), {}, {"output_shape": (9, 16)}),
    ("cplx_d2_norm_ortho", 2, 3, 4, (12, 12), (10, 8), {"fft_norm": "ortho"}, {}),
    ("cplx_d2_norm_backward", 2, 3, 4, (12, 12), (10, 8), {"fft_norm": "backward"}, {}),
    ("cplx_d2_no_bias", 2, 3, 4, (12, 12), (6, 6), {"bias": False}, {}),
    ("cplx_d3", 1, 3, 4, (8, 6, 10), (4, 4, 6), {}, {}),
    ("cplx_d4", 1, 2, 2, (6, 6, 6, 6), (4, 4, 4, 4), {}, {}),
    ("cplx_d2_tucker_reconstructed", 2, 6, 5, (16, 12), (8, 6), {"factorization": "tucker", "rank": [4, 3, 5, 3]}, {}),
    # separable=True (one channel axis, mode-wise product); appended so that the seeds of the cases above do not move
    ("cplx_d2_separable", 2, 5, 5, (16, 12), (8, 6), {"separable": True}, {}),
    ("cplx_d1_separable_upsample", 2, 4, 4, (16,), (6,), {"separable": True, "resolution_scaling_factor": 2}, {}),
    ("cplx_d3_separable_max_modes", 1, 3, 3, (8, 6, 10), (3, 4, 5), {"separable": True, "max_n_modes": (4, 4, 6)}, {}),
    ("cplx_d2_separable_tucker", 2, 5, 5, (16, 12), (8, 6), {"separable": True, "factorization": "tucker", "rank": [3, 5, 3]}, {}),
]


def _np(t):
    t = t.detach()
    return torch.view_as_real(t).numpy() if t.is_complex() else t.numpy()


def main():
    ref = load_reference_spectral_conv()
    index = {}
    for seed, (name, B, Ci, Co, grid, modes, ckw, fkw) in enumerate(CASES):
        torch.manual_seed(5000 + seed)
        conv = ref.SpectralConv(Ci, Co, modes, complex_data=True, **ckw)
        x = torch.randn(B, Ci, *grid, dtype=torch.cfloat, requires_grad=True)
        y = conv(x, **fkw)
        gy = torch.randn_like(y)
        y.backward(gy)
        # complex tensors are stored as (..., 2) real arrays (keys suffixed __c)
        arrays = {"x__c": _np(x), "gy__c": _np(gy), "y__c": _np(y), "dx__c": _np(x.grad)}
        pnames = []
        for pname, p in conv.named_parameters():
            key = pname.replace(".", "__")
            arrays["p__" + key + ("__c" if p.is_complex() else "")] = _np(p)
            arrays["g__" + key + ("__c" if p.is_complex() else "")] = _np(p.grad)
            pnames.append(pname)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        index[name] = {
            "B": B, "in_channels": Ci, "out_channels": Co, "grid": list(grid), "n_modes": list(modes),
            "ctor": {k: (list(v) if isinstance(v, tuple) else v) for k, v in ckw.items()},
            "forward": {k: list(v) for k, v in fkw.items()},
            "stored_n_modes": list(conv.n_modes), "max_n_modes": list(conv.max_n_modes),
            "params": pnames, "weight_kind": conv.weight.name.lower(), "out_grid": list(y.shape[2:]),
        }
        print(f"{name:32s} y{tuple(y.shape)} params={pnames}")
    with open(os.path.join(OUT, "complex_index.json"), "w") as f:
        json.dump({"reference": "neuraloperator@93d3f06 neuralop/layers/spectral_convolution.py (complex_data=True)",
                   "generator": "oracle/make_golden_complex.py", "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
