#!/usr/bin/env python
"""Measurement-only helper (NOT product code, not used by bench.py): times the reference's op sequence
(torch.fft + einsum, i.e. PyTorch eager + cuFFT/cuBLAS, via the oracle port) ON THE GPU for the BASELINE
configs, next to our kernels in the same process.  This is the 1.5x denominator named in BASELINE.md section 3.

    python scripts/bench_torch_gpu.py [--configs 2,4,5a] > gpurun_out/torch_gpu.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neuraloperator_b200 as nb  # noqa: E402
from oracle import spectral_conv_oracle as O  # noqa: E402

CONFIGS = {
    "1": (16, 32, (1024,), (16,)),
    "2": (32, 64, (128, 128), (32, 32)),
    "4": (8, 32, (64, 64, 64), (16, 16, 16)),
    "5a": (16, 64, (256, 256), (64, 64)),
    "5b": (16, 64, (512, 512), (64, 64)),
    "5c": (16, 64, (1024, 1024), (64, 64)),
}


def time_fn(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def tucker_config(dev, reps):
    """BASELINE config 3: TFNO2d Darcy, Tucker ranks (36, 36, 18, 10), implementation="factorized"."""
    B, C, grid, modes, ranks = 32, 64, (128, 128), (32, 32), [36, 36, 18, 10]
    x, w, bias, gy = O.make_inputs(B, C, C, grid, modes, seed=0, kind="tucker", ranks=ranks)
    xd, gd = x.to(dev), gy.to(dev)
    core = w.core.to(dev).requires_grad_(True)
    factors = [f.to(dev).requires_grad_(True) for f in w.factors]
    wd = O.Weight("tucker", core=core, factors=factors)
    bd = bias.to(dev).requires_grad_(True)

    def torch_step():
        xx = xd.detach().requires_grad_(True)
        core.grad = None
        for f in factors:
            f.grad = None
        # torch.einsum without opt_einsum contracts left to right and would materialise a 1.9 PiB intermediate for the
        # reference's 6-operand expression; the best case for the eager path is "reconstruct W, then the dense contraction"
        dense = O.Weight("dense", tensor=O.tucker_to_dense(core, factors))
        y = O.spectral_conv_forward(xx, dense, bd, modes)
        y.backward(gd)

    conv = nb.SpectralConv(C, C, modes, factorization="tucker", rank=ranks, implementation="factorized").to(dev)
    with torch.no_grad():
        for dst, src in zip(conv.weight.decomposition(), w.params()):
            dst.copy_(src.to(dev))
        conv.bias.copy_(bias.to(dev))

    def our_step():
        xx = xd.detach().requires_grad_(True)
        for prm in conv.parameters():
            prm.grad = None
        y = conv(xx)
        y.backward(gd)

    t_ref = time_fn(torch_step, reps=reps)
    t_our = time_fn(our_step, reps=reps)
    return {"shape": [B, C, *grid], "modes": list(modes), "tucker_ranks": ranks, "torch_cufft_ms": t_ref, "ours_ms": t_our,
            "torch_samples_per_s": B / t_ref * 1e3, "ours_samples_per_s": B / t_our * 1e3, "speedup": t_ref / t_our}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,5a")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {}
    for name in args.configs.split(","):
        if name == "3":
            out["3"] = tucker_config(dev, args.reps)
            print("3", json.dumps(out["3"]), flush=True)
            continue
        B, C, grid, modes = CONFIGS[name]
        x, w, bias, gy = O.make_inputs(B, C, C, grid, modes, seed=0)
        xd, gd = x.to(dev), gy.to(dev)
        wd = O.Weight("dense", tensor=w.tensor.to(dev).requires_grad_(True))
        bd = bias.to(dev).requires_grad_(True)

        def torch_step():
            xx = xd.detach().requires_grad_(True)
            wd.tensor.grad = None
            bd.grad = None
            y = O.spectral_conv_forward(xx, wd, bd, modes)
            y.backward(gd)

        conv = nb.SpectralConv(C, C, modes).to(dev)
        with torch.no_grad():
            conv.weight.tensor.copy_(w.tensor.to(dev))
            conv.bias.copy_(bias.to(dev))

        def our_step():
            xx = xd.detach().requires_grad_(True)
            conv.weight.tensor.grad = None
            conv.bias.grad = None
            y = conv(xx)
            y.backward(gd)

        t_ref = time_fn(torch_step, reps=args.reps)
        t_our = time_fn(our_step, reps=args.reps)
        kept = nb.get_plan(dev, grid, grid, conv.n_modes, conv.max_n_modes).kept
        S = 1
        for g in grid:
            S *= g
        M = 1
        for k in kept:
            M *= k
        bytes_step = 16 * B * C * S + 24 * C * C * M + 16 * B * C * M
        out[name] = {"shape": [B, C, *grid], "modes": list(modes), "torch_cufft_ms": t_ref, "ours_ms": t_our,
                     "torch_samples_per_s": B / t_ref * 1e3, "ours_samples_per_s": B / t_our * 1e3,
                     "speedup": t_ref / t_our, "algorithmic_bytes": bytes_step,
                     "ours_gbs": bytes_step / t_our / 1e6}
        print(name, json.dumps(out[name]), flush=True)
    print(json.dumps({"torch_gpu_baseline": out}))


if __name__ == "__main__":
    main()
