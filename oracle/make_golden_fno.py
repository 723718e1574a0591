"""Mint a model-level golden from the UNMODIFIED reference `neuralop.models.FNO` (positional embedding and domain padding off, so that
forward = lifting -> n_layers x FNOBlocks -> projection, fno.py:384-404): x, the whole state dict, y, dx and every parameter gradient.
TEST INFRASTRUCTURE; run in the build container:  python oracle/make_golden_fno.py"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.load_reference import REF_ROOT, load_reference_spectral_conv  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference_fno():
    load_reference_spectral_conv()
    if "neuralop.models" not in sys.modules:
        m = types.ModuleType("neuralop.models")
        m.__path__ = [os.path.join(REF_ROOT, "neuralop", "models")]
        sys.modules["neuralop.models"] = m
    return importlib.import_module("neuralop.models.fno")


CASES = [
    ("fno_d2_small", dict(n_modes=(8, 8), in_channels=2, out_channels=1, hidden_channels=8, n_layers=3), (2, 2, 16, 16)),
    ("fno_d1_small", dict(n_modes=(12,), in_channels=1, out_channels=2, hidden_channels=6, n_layers=2), (3, 1, 40)),
    ("tfno_d2_small", dict(n_modes=(8, 6), in_channels=1, out_channels=1, hidden_channels=6, n_layers=2, factorization="tucker",
                           implementation="factorized", rank=[3, 3, 4, 3]), (2, 1, 16, 12)),
]


def main():
    fno = load_reference_fno()
    index = {}
    for seed, (name, kw, shape) in enumerate(CASES):
        torch.manual_seed(9000 + seed)
        model = fno.FNO(positional_embedding=None, domain_padding=None, **kw)
        with torch.no_grad():
            for pname, p in model.named_parameters():
                if "channel_mlp_skips" in pname:
                    p.add_(0.3 * torch.randn_like(p))
        x = torch.randn(*shape, requires_grad=True)
        y = model(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        arrays = {"x": x.detach().numpy(), "gy": gy.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
        for pname, p in model.named_parameters():
            key = pname.replace(".", "__")
            for tag, val in (("p__", p.detach()), ("g__", p.grad)):
                arrays[tag + key + ("__c" if val.is_complex() else "")] = torch.view_as_real(val).numpy() if val.is_complex() else val.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        index[name] = {"ctor": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "shape": list(shape),
                       "params": [n for n, _ in model.named_parameters()]}
        print(name, tuple(y.shape), len(index[name]["params"]), "parameters")
    with open(os.path.join(OUT, "fno_index.json"), "w") as f:
        json.dump({"reference": "neuraloperator@93d3f06 neuralop/models/fno.py (FNO, positional_embedding=None, domain_padding=None)",
                   "generator": "oracle/make_golden_fno.py", "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
