"""Two-GPU tests (skipped on a single-GPU box): the library's own all-reduce over NVLink peer memory, and the data-parallel
backward pass that uses it -- gradients must equal the average of the per-rank gradients, and match NCCL's result."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import neuraloperator_b200 as nb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    msgs = []
    try:
        # ---- the collective alone
        red = nb.PeerGradientAllReducer([], n_ctas=8)
        for n in (4, 1000, 1 << 20, 4456512):
            buf = red.grad_buffer(n, dev)
            torch.manual_seed(100 + rank)
            local = torch.randn(n, device=dev)
            buf.copy_(local)
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            want = sum(gathered) / world
            torch.cuda.synchronize(dev)
            dist.barrier()
            red.reduce_in_backward([buf])
            red.finish()
            torch.cuda.synchronize(dev)
            err = (buf - want).abs().max().item()
            if err > 1e-6:
                ok = False
                msgs.append(f"allreduce n={n}: max err {err}")
        # ---- data-parallel backward: peer-memory reducer vs the average of the per-rank gradients
        torch.manual_seed(0)
        conv = nb.SpectralConv(8, 8, (32, 32)).to(dev)          # same parameters on every rank
        torch.manual_seed(10 + rank)
        x = torch.randn(4, 8, 128, 128, device=dev, requires_grad=True)
        g = torch.randn(4, 8, 128, 128, device=dev)
        conv(x).backward(g)
        gw_local, gb_local = conv.weight.tensor.grad.clone(), conv.bias.grad.clone()
        conv.weight.tensor.grad = None
        conv.bias.grad = None
        lw = [torch.empty_like(gw_local) for _ in range(world)]
        lb = [torch.empty_like(gb_local) for _ in range(world)]
        dist.all_gather(lw, gw_local)
        dist.all_gather(lb, gb_local)
        want_w, want_b = sum(lw) / world, sum(lb) / world
        conv.gradient_reducer = nb.PeerGradientAllReducer(conv.parameters(), n_ctas=8)
        for _ in range(2):                                       # second pass reuses the persistent buffer
            conv.weight.tensor.grad = None
            conv.bias.grad = None
            x.grad = None
            conv(x).backward(g)
            conv.gradient_reducer.finish()
            torch.cuda.synchronize(dev)
            ew = (conv.weight.tensor.grad - want_w).abs().max().item() / want_w.abs().max().item()
            eb = (conv.bias.grad - want_b).abs().max().item() / want_b.abs().max().item()
            if ew > 1e-5 or eb > 1e-5:
                ok = False
                msgs.append(f"data-parallel grads: rel err dW {ew} db {eb}")
    except Exception as exc:   # noqa: BLE001
        ok = False
        msgs.append(repr(exc)[:400])
    out[rank] = (ok, msgs)
    torch.cuda.synchronize(dev)
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_allreduce_and_data_parallel_backward(world):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        ctx = mp.spawn(_worker, args=(world, port, out), nprocs=world, join=False)
        for p in ctx.processes:
            p.join(180)
        res = dict(out)
    assert len(res) == world, f"workers did not report: {res}"
    for r in range(world):
        assert res[r][0], res[r][1]
