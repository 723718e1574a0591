import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuraloperator_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
Ci, Co, Mt = 64, 64, 544
w = torch.randn(Ci, Co, Mt, dtype=torch.complex64, device=dev)
cyc = torch.zeros(2 * (Mt // 4), dtype=torch.int64, device=dev)
big = torch.empty(64 * 1024 * 1024, device=dev)          # 256 MB: flushes L2
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for cold in (False, True, False, True):
    if cold:
        big.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.sc_probe_tma_gather(ctypes.c_void_p(w.data_ptr()), Ci, Co, Mt, ctypes.c_void_p(cyc.data_ptr()), st), "probe")
    e1.record()
    torch.cuda.synchronize()
    c = cyc.view(-1, 2).cpu()
    print("cold" if cold else "warm", "kernel us", round(e0.elapsed_time(e1) * 1e3, 2), "issue cycles med", int(c[:, 0].median()),
          "all-landed cycles min/med/max", int(c[:, 1].min()), int(c[:, 1].median()), int(c[:, 1].max()),
          "-> B/clk/SM at median", round(128 * 1024 / float(c[:, 1].median()), 1))
