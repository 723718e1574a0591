"""Pure-read / pure-write / copy bandwidth of this B200 with PyTorch's own kernels (context for the transform kernels: the analysis
launch is a pure read stream, the synthesis launch a pure write stream; MEASURED_PEAKS.json's figure is a copy, half reads half writes)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29                       # 2 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
w = t(lambda: a.fill_(1.0))
r = t(lambda: a.sum())
r2 = t(lambda: a.max())
c = t(lambda: b.copy_(a))
gb = n * 4 / 1e9
print(f"write (fill_) {gb / w * 1e3:.0f} GB/s | read (sum) {gb / r * 1e3:.0f} GB/s | read (max) {gb / r2 * 1e3:.0f} GB/s | copy {2 * gb / c * 1e3:.0f} GB/s (read+write bytes)")
