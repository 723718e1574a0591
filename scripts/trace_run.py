import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloperator_b200 as nb
dev = torch.device("cuda:0")
conv = nb.SpectralConv(64, 64, (32, 32)).to(dev)
x = torch.randn(32, 64, 128, 128, device=dev, requires_grad=True)
g = torch.randn(32, 64, 128, 128, device=dev)
for _ in range(2):
    y = conv(x); y.backward(g)
torch.cuda.synchronize()
