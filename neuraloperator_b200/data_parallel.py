"""Batch-sharded data parallelism for the spectral conv: ONE gradient all-reduce per step.

The reference wraps the model in DDP (neuralop/training/trainer.py:203-205): samples are independent in the forward pass and
in dx; only dweight / dbias sum over the batch.  `GradientAllReducer` averages gradient tensors across ranks on a side
stream (complex64 grads travel as (re, im) float pairs, in place, no packing copy), so that the collective overlaps whatever
the caller launches next.  Attached to a `SpectralConv` (`conv.gradient_reducer = reducer`) the backward pass itself starts
the all-reduce of dweight / dbias as soon as the contraction backward has produced them, i.e. underneath the dx synthesis
kernel -- the DDP overlap, for a single layer.  Backend: NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter] = (), process_group=None, average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.average = average
        self._stream: Optional[torch.cuda.Stream] = None
        self._pending = []          # (work, tensor) pairs of the collectives in flight
        self._early = False         # True once a backward pass has already started this step's collectives

    @staticmethod
    def _as_real(t: torch.Tensor) -> torch.Tensor:
        return torch.view_as_real(t) if t.is_complex() else t

    @staticmethod
    def _coalesce(tensors):
        """Merge float32 tensors that sit back to back in one storage (SpectralConv.backward lays dweight and dbias out
        that way) into single flat views: one collective instead of one per tensor."""
        out, i = [], 0
        while i < len(tensors):
            t = tensors[i]
            j = i
            end = t.data_ptr() + t.numel() * t.element_size()
            total = t.numel()
            while (j + 1 < len(tensors) and t.is_contiguous() and tensors[j + 1].is_contiguous()
                   and tensors[j + 1].dtype == t.dtype and tensors[j + 1].device == t.device
                   and tensors[j + 1].untyped_storage().data_ptr() == t.untyped_storage().data_ptr()
                   and tensors[j + 1].data_ptr() == end):
                j += 1
                end += tensors[j].numel() * tensors[j].element_size()
                total += tensors[j].numel()
            if j > i:
                out.append(torch.as_strided(t, (total,), (1,), storage_offset=t.storage_offset()))
            else:
                out.append(t)
            i = j + 1
        return out

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _backend_has_avg(self) -> bool:
        return dist.get_backend(self.group) == "nccl"

    def start_tensors(self, tensors: Iterable[torch.Tensor]):
        """Launch the all-reduce of `tensors` (in place). On CUDA it runs on a side stream that first waits for the work
        already queued on the current stream (the kernels that produced the tensors)."""
        tensors = self._coalesce([self._as_real(t) for t in tensors if t is not None])
        if self.world_size() == 1 or not tensors:
            return
        device = tensors[0].device
        op = dist.ReduceOp.AVG if (self.average and self._backend_has_avg()) else dist.ReduceOp.SUM
        if device.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=device)
            self._stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(self._stream):
                for t in tensors:
                    r = self._as_real(t)
                    r.record_stream(self._stream)
                    self._pending.append((dist.all_reduce(r, op=op, group=self.group, async_op=True), r, op))
        else:
            for t in tensors:
                r = self._as_real(t)
                self._pending.append((dist.all_reduce(r, op=op, group=self.group, async_op=True), r, op))

    def start_early(self, tensors):
        """Called from inside a backward pass (SpectralConv with a reducer attached)."""
        self._early = True
        self.start_tensors(tensors)

    def start(self):
        """All-reduce the `.grad` of the registered parameters, unless a backward pass already did it for this step."""
        if self._early:
            return
        self.start_tensors([p.grad for p in self.params if p.grad is not None])

    def finish(self):
        """Wait for the collectives; the current stream then sees the averaged gradients."""
        if not self._pending:
            self._early = False
            return
        device = self._pending[0][1].device
        scale = 1.0 / self.world_size()
        if device.type == "cuda":
            with torch.cuda.stream(self._stream):
                for work, r, op in self._pending:
                    work.wait()
                    if self.average and op == dist.ReduceOp.SUM:
                        r.mul_(scale)
            torch.cuda.current_stream(device).wait_stream(self._stream)
        else:
            for work, r, op in self._pending:
                work.wait()
                if self.average and op == dist.ReduceOp.SUM:
                    r.mul_(scale)
        self._pending = []
        self._early = False

    def all_reduce(self):
        self.start()
        self.finish()
