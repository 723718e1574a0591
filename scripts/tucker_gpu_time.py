"""cfg-3 (TFNO2d Darcy, Tucker ranks (36,36,18,10), implementation="factorized") through this library only: eager wall/event
time per step; run it under `ncu --metrics gpu__time_duration.sum` to get the device-busy time per step next to it."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloperator_b200 as nb
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
conv = nb.SpectralConv(64, 64, (32, 32), factorization="tucker", rank=[36, 36, 18, 10], implementation="factorized").to(dev)
x = torch.randn(32, 64, 128, 128, device=dev)
g = torch.randn(32, 64, 128, 128, device=dev)
def step():
    xx = x.detach().requires_grad_(True)
    for p in conv.parameters():
        p.grad = None
    conv(xx).backward(g)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps):
    step()
e1.record(); t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"tucker cfg-3: {e0.elapsed_time(e1) / steps:.3f} ms/step (events), host issue time {1e3 * t_host / steps:.3f} ms/step")
