"""Batch-sharded data parallelism for the spectral conv: ONE gradient all-reduce per step.

The reference wraps the model in DDP (neuralop/training/trainer.py:203-205): samples are independent in the forward pass and
in dx; only dweight / dbias sum over the batch.  `GradientAllReducer` averages gradient tensors across ranks on a side
stream (complex64 grads travel as (re, im) float pairs, in place, no packing copy).

Attached to a `SpectralConv` (`conv.gradient_reducer = reducer`) the backward pass itself all-reduces dweight / dbias: the
library records an event as soon as the dweight kernel has been launched (`sc_backward_dense(..., grads_ready)`), the
collective stream waits on it, and the collective runs underneath the dxm contraction and the dx synthesis kernel -- the DDP
overlap, for a single layer.  Before `backward` returns, the compute stream is made to wait for the collective, so whatever
autograd does next with the returned tensors (steal them as `.grad`, or accumulate them into an existing `.grad`) is ordered
after the in-place reduction.  Gradients reduced that way are remembered (by storage address) and skipped by `start()`, which
reduces the `.grad` of every OTHER registered parameter.  Backend: NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU
tests.
"""
import ctypes
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter] = (), process_group=None, average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.average = average
        self._stream: Optional[torch.cuda.Stream] = None
        self._pending = []            # (work, tensor, op) of the collectives in flight
        self._reduced_in_backward = set()   # data_ptr of gradient buffers a backward pass already reduced this step
        self._event = None            # sc_event handle (created on first use)

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

    def _side_stream(self, device) -> torch.cuda.Stream:
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def grads_ready_event(self) -> ctypes.c_void_p:
        """The `grads_ready` event handed to sc_backward_dense (created once per reducer)."""
        if self._event is None:
            from . import _lib
            ev = ctypes.c_void_p()
            _lib.check(_lib.load().sc_event_create(ctypes.byref(ev)), "sc_event_create")
            self._event = ev
        return self._event

    def start_tensors(self, tensors: Iterable[torch.Tensor], after_event: Optional[ctypes.c_void_p] = None):
        """Launch the all-reduce of `tensors` (in place). On CUDA it runs on a side stream that first waits for `after_event`
        (an sc_event recorded on the compute stream) or, without one, for everything already queued on the current stream."""
        tensors = self._coalesce([self._as_real(t) for t in tensors if t is not None])
        if self.world_size() == 1 or not tensors:
            return
        device = tensors[0].device
        op = dist.ReduceOp.AVG if (self.average and self._backend_has_avg()) else dist.ReduceOp.SUM
        if device.type == "cuda":
            side = self._side_stream(device)
            if after_event is not None:
                from . import _lib
                _lib.check(_lib.load().sc_stream_wait_event(ctypes.c_void_p(side.cuda_stream), after_event), "sc_stream_wait_event")
            else:
                side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for t in tensors:
                    r = self._as_real(t)
                    r.record_stream(side)
                    self._pending.append((dist.all_reduce(r, op=op, group=self.group, async_op=True), r, op))
        else:
            for t in tensors:
                r = self._as_real(t)
                self._pending.append((dist.all_reduce(r, op=op, group=self.group, async_op=True), r, op))

    def reduce_in_backward(self, tensors, after_event: Optional[ctypes.c_void_p] = None):
        """Called from inside a backward pass with the gradients it is about to return: all-reduce them (after `after_event`,
        i.e. underneath the kernels already queued behind it) and order the compute stream after the result."""
        tensors = [t for t in tensors if t is not None]
        if self.world_size() == 1 or not tensors:
            return
        self.start_tensors(tensors, after_event=after_event)
        self._wait_pending()
        for t in tensors:
            self._reduced_in_backward.add(t.data_ptr())

    def start(self):
        """All-reduce the `.grad` of every registered parameter that a backward pass has not reduced already."""
        grads = [p.grad for p in self.params if p.grad is not None and p.grad.data_ptr() not in self._reduced_in_backward]
        self.start_tensors(grads)

    def _wait_pending(self):
        if not self._pending:
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

    def finish(self):
        """Wait for the collectives; the current stream then sees the averaged gradients.  Ends the step."""
        self._wait_pending()
        self._reduced_in_backward.clear()

    def all_reduce(self):
        self.start()
        self.finish()

    def __del__(self):
        try:
            if self._event is not None:
                from . import _lib
                _lib.load().sc_event_destroy(self._event)
                self._event = None
        except Exception:   # noqa: BLE001
            pass


class PeerGradientAllReducer(GradientAllReducer):
    """`GradientAllReducer` whose collective is this library's own kernel over NVLink peer memory (`sc_allreduce_p2p`).

    The gradients of the attached `SpectralConv` (dweight and dbias, back to back) are written by the backward kernels straight
    into a CUDA symmetric-memory buffer (`torch.distributed._symmetric_memory`: every rank maps every other rank's buffer), and
    the two-shot all-reduce kernel averages them in place on the collective stream, behind the library's `grads_ready` event --
    no NCCL call, no copy, ~12 CTAs.  The buffer is persistent: the `.grad` tensors autograd hands out are views of it and are
    overwritten by the next backward pass (the usual `zero_grad(set_to_none=True)` training loop).
    Everything that does not live in that buffer (other parameters) still goes through the base class (NCCL / gloo)."""

    def __init__(self, params: Iterable[torch.nn.Parameter] = (), process_group=None, average: bool = True, n_ctas: int = 12):
        super().__init__(params, process_group, average)
        self.n_ctas = int(n_ctas)
        self._sym = {}            # (n_floats, device index) -> (tensor, padded length, peer buffer array, signal pad array)

    def grad_buffer(self, n_floats: int, device) -> torch.Tensor:
        """A persistent symmetric-memory float32 buffer of `n_floats` elements (collective on first use: every rank calls it)."""
        key = (int(n_floats), device.index)
        ent = self._sym.get(key)
        if ent is None:
            import torch.distributed._symmetric_memory as symm_mem
            padded = (int(n_floats) + 3) // 4 * 4
            t = symm_mem.empty(padded, dtype=torch.float32, device=device)
            group = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(t, group)
            world = hdl.world_size
            off = int(getattr(hdl, "offset", 0) or 0)
            bufs = (ctypes.c_void_p * world)(*[int(p) + off for p in hdl.buffer_ptrs])
            sigs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.signal_pad_ptrs])
            if int(hdl.signal_pad_size) < 4 * self.n_ctas * world:
                raise RuntimeError("symmetric-memory signal pad too small for the requested number of CTAs")
            ent = self._sym[key] = (t, padded, bufs, sigs, hdl)
        return ent[0][:n_floats]

    def _entry_of(self, tensors):
        for t in tensors:
            for ent in self._sym.values():
                if t.untyped_storage().data_ptr() == ent[0].untyped_storage().data_ptr():
                    return ent
        return None

    def reduce_in_backward(self, tensors, after_event: Optional[ctypes.c_void_p] = None):
        tensors = [t for t in tensors if t is not None]
        world = self.world_size()
        if world == 1 or not tensors:
            return
        ent = self._entry_of(tensors)
        if ent is None:                      # not in the symmetric buffer: NCCL / gloo path of the base class
            return super().reduce_in_backward(tensors, after_event)
        from . import _lib
        lib = _lib.load()
        buf, padded, bufs, sigs, hdl = ent
        device = buf.device
        side = self._side_stream(device)
        if after_event is not None:
            _lib.check(lib.sc_stream_wait_event(ctypes.c_void_p(side.cuda_stream), after_event), "sc_stream_wait_event")
        else:
            side.wait_stream(torch.cuda.current_stream(device))
        scale = 1.0 / world if self.average else 1.0
        with torch.cuda.device(device):
            _lib.check(lib.sc_allreduce_p2p(bufs, sigs, dist.get_rank(self.group), world, padded, scale, self.n_ctas,
                                            ctypes.c_void_p(side.cuda_stream)), "sc_allreduce_p2p")
        torch.cuda.current_stream(device).wait_stream(side)
        for t in tensors:
            self._reduced_in_backward.add(t.data_ptr())
