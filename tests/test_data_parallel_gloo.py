"""world_size-2 gloo test of the one collective on the path: the gradient all-reduce (CPU tensors)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuraloperator_b200.data_parallel import GradientAllReducer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    w = torch.nn.Parameter(torch.zeros(3, 4, 5, dtype=torch.cfloat))
    b = torch.nn.Parameter(torch.zeros(4, 1, 1))
    w.grad = torch.randn(3, 4, 5, dtype=torch.cfloat)
    b.grad = torch.randn(4, 1, 1)
    local = (w.grad.clone(), b.grad.clone())
    red = GradientAllReducer([w, b])
    red.start()
    red.finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    mean_w = sum(g[0] for g in gathered) / world
    mean_b = sum(g[1] for g in gathered) / world
    ok = torch.allclose(w.grad, mean_w, atol=1e-6) and torch.allclose(b.grad, mean_b, atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_all_reduce_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert out[0] and out[1]


def _worker_mixed(rank, world, port, out):
    """A backward pass reduces its own gradients (`reduce_in_backward`); `start()` must still reduce every OTHER parameter
    (round-1 advisor finding: one global flag used to turn `start()` into a no-op for all of them)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(200 + rank)
    w = torch.nn.Parameter(torch.zeros(3, 4, 5, dtype=torch.cfloat))
    b = torch.nn.Parameter(torch.zeros(4, 1, 1))
    extra = torch.nn.Parameter(torch.zeros(7))
    # the peer-memory reducer must behave like the base class for tensors that do not live in its symmetric buffer (CPU here)
    from neuraloperator_b200.data_parallel import PeerGradientAllReducer
    red = PeerGradientAllReducer([w, b, extra])
    ok = True
    for step in range(2):                       # two steps: the per-step bookkeeping is reset by finish()
        gw, gb = torch.randn(3, 4, 5, dtype=torch.cfloat), torch.randn(4, 1, 1)
        ge = torch.randn(7)
        local = (gw.clone(), gb.clone(), ge.clone())
        red.reduce_in_backward([gw, gb])        # what SpectralConv.backward does with dweight / dbias
        w.grad, b.grad, extra.grad = gw, gb, ge  # autograd steals the buffers
        red.start()
        red.finish()
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        means = [sum(g[i] for g in gathered) / world for i in range(3)]
        ok = ok and torch.allclose(w.grad, means[0], atol=1e-6) and torch.allclose(b.grad, means[1], atol=1e-6) \
            and torch.allclose(extra.grad, means[2], atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_backward_reduced_and_other_parameters_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_mixed, args=(world, port, out), nprocs=world, join=True)
        assert out[0] and out[1]
