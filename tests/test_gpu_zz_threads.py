"""Runs last among the GPU tests (file name): thread / context corner cases of the C library."""
import pytest
import torch

import neuraloperator_b200 as nb

pytestmark = pytest.mark.gpu


def test_first_library_call_of_a_thread_is_a_tensor_map_encode(cuda_device):
    """PyTorch runs backward on autograd worker threads that have no CUDA context bound until their first runtime call; when
    every allocation of that backward is served from the caching allocator, the first CUDA-facing call of the thread is the
    driver-side cuTensorMapEncodeTiled of the fused kernels (cache miss for a new buffer).  It must bind the context itself."""
    import threading
    from oracle import spectral_conv_oracle as O
    grid, modes = (64, 64), (16, 16)
    stored = O.stored_n_modes(modes)
    plan = nb.get_plan(cuda_device, grid, grid, stored, stored)
    assert plan.uses_fast_path() & 1
    x = torch.randn(4, 8, *grid, device=cuda_device)
    ref = nb.analyze(plan, x)                       # main thread: also leaves `modes` / workspace sized blocks in the allocator cache
    fresh = x.clone()                               # new buffer -> tensor-map cache miss on the worker thread
    torch.cuda.synchronize()
    del_me = nb.analyze(plan, x)
    del del_me
    out = {}

    def worker():
        try:
            out["m"] = nb.analyze(plan, fresh)
        except Exception as e:                      # noqa: BLE001
            out["err"] = e

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert "err" not in out, out.get("err")
    torch.cuda.synchronize()
    assert torch.equal(out["m"], ref)
