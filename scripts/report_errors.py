"""Prints max|d|/max|ref| of y, dx, dW, db against the CPU oracle for the BASELINE configs (measurement helper)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloperator_b200 as nb
from oracle import spectral_conv_oracle as O
dev = torch.device("cuda:0")
out = {}
for name, (B, C, grid, modes) in {"cfg1": (16, 32, (1024,), (16,)), "cfg2": (32, 64, (128, 128), (32, 32)),
                                  "cfg4": (8, 32, (64, 64, 64), (16, 16, 16))}.items():
    x, w, bias, gy = O.make_inputs(B, C, C, grid, modes, seed=0)
    y_ref, dx_ref, dws_ref, db_ref = O.spectral_conv_fwd_bwd(x, w, bias, gy, modes)
    conv = nb.SpectralConv(C, C, modes).to(dev)
    with torch.no_grad():
        conv.weight.tensor.copy_(w.tensor.to(dev)); conv.bias.copy_(bias.to(dev))
    xd = x.to(dev).requires_grad_(True)
    y = conv(xd); y.backward(gy.to(dev))
    rel = lambda a, b: ((a.detach().cpu() - b).abs().max() / b.abs().max()).item()
    plan = nb.get_plan(dev, grid, grid, conv.n_modes, conv.max_n_modes)
    out[name] = {"fast_path_mask": plan.uses_fast_path(), "y": rel(y, y_ref), "dx": rel(xd.grad, dx_ref),
                 "dW": rel(conv.weight.tensor.grad, dws_ref[0]), "db": rel(conv.bias.grad, db_ref)}
print(json.dumps(out, indent=1))
