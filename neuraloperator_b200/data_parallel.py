"""Batch-sharded data parallelism for the spectral conv: ONE gradient all-reduce per step.

The reference wraps the model in DDP (neuralop/training/trainer.py:203-205): samples are independent in the
forward pass and in dx; only dweight / dbias sum over the batch.  `GradientAllReducer` does the same thing for
the parameters it is given -- averages `.grad` across ranks with a single all-reduce on one flat float32
buffer (complex64 grads travel as (re, im) float pairs) -- issued on a side stream so that it overlaps
whatever the caller launches next (the dx inverse transforms, the previous layer's backward).
Backend: NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.average = average
        self._flat: Optional[torch.Tensor] = None
        self._stream: Optional[torch.cuda.Stream] = None
        self._work = None

    @staticmethod
    def _as_real(t: torch.Tensor) -> torch.Tensor:
        return torch.view_as_real(t) if t.is_complex() else t

    def _numel(self) -> int:
        return sum(self._as_real(p).numel() for p in self.params)

    def _ensure_buffer(self, device):
        n = self._numel()
        if self._flat is None or self._flat.numel() != n or self._flat.device != device:
            self._flat = torch.empty(n, dtype=torch.float32, device=device)
        if device.type == "cuda" and self._stream is None:
            self._stream = torch.cuda.Stream(device=device)

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def start(self):
        """Pack the grads and launch the all-reduce (non-blocking on CUDA). No-op for a single rank."""
        if self.world_size() == 1 or not self.params:
            return
        device = self.params[0].grad.device
        self._ensure_buffer(device)
        if device.type == "cuda":
            self._stream.wait_stream(torch.cuda.current_stream(device))
            ctx = torch.cuda.stream(self._stream)
        else:
            ctx = torch.autograd.profiler.record_function("grad_allreduce")
        with ctx:
            off = 0
            for p in self.params:
                g = self._as_real(p.grad).reshape(-1)
                self._flat[off:off + g.numel()].copy_(g)
                off += g.numel()
            self._work = dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for the all-reduce and scatter the averaged values back into `.grad`."""
        if self._work is None:
            return
        device = self._flat.device
        if device.type == "cuda":
            ctx = torch.cuda.stream(self._stream)
        else:
            ctx = torch.autograd.profiler.record_function("grad_allreduce_unpack")
        with ctx:
            self._work.wait()
            scale = 1.0 / self.world_size() if self.average else 1.0
            off = 0
            for p in self.params:
                g = self._as_real(p.grad)
                n = g.numel()
                g.copy_((self._flat[off:off + n] * scale).view_as(g))
                off += n
        if device.type == "cuda":
            torch.cuda.current_stream(device).wait_stream(self._stream)
        self._work = None

    def all_reduce(self):
        self.start()
        self.finish()
