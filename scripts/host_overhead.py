"""Host-side cost of one fwd+bwd through the nn.Module (tiny tensors: GPU time negligible)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloperator_b200 as nb
dev = torch.device("cuda:0")
for shape, modes in [((2, 8, 128, 128), (32, 32)), ((2, 8, 32, 32), (16, 16))]:
    conv = nb.SpectralConv(shape[1], shape[1], modes).to(dev)
    x = torch.randn(*shape, device=dev, requires_grad=True)
    g = torch.randn(*shape, device=dev)
    for _ in range(20):
        conv(x).backward(g)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        conv.weight.tensor.grad = None; conv.bias.grad = None; x.grad = None
        conv(x).backward(g)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{shape} modes {modes}: host enqueue {1e6*(t1-t0)/n:.1f} us/step, total {1e6*(t2-t0)/n:.1f} us/step")
