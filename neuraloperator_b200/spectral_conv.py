"""`SpectralConv`: B200-native drop-in for `neuralop.layers.spectral_convolution.SpectralConv`.

Same constructor signature (reference spectral_convolution.py:285-305), same `forward(x, output_shape=None)`,
`transform`, mutable `n_modes`, `max_n_modes`, `weight`, `bias` (:383-415) -- so it can be passed as
`conv_module=` to the reference's `FNOBlocks` / `FNO` / `TFNO` (fno_block.py:163,210-240; fno.py:208,315).
Everything between the input tensor and the output tensor runs in hand-written sm_100a kernels behind the
C ABI of `include/spectral_conv_b200.h`; there is no PyTorch / cuFFT / CPU fallback.  PyTorch is used for
device memory, the stream, and the autograd graph node.
"""
import ctypes
import threading
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _lib
from .factorized import FactorizedWeight

Number = Union[int, float]


# --------------------------------------------------------------------------------------------------
# plans (kept-mode index set + twiddle tables), cached per shape: n_modes and the grid may change between
# calls (incremental training mutates n_modes, resolution invariance changes the grid)
# --------------------------------------------------------------------------------------------------
class Plan:
    def __init__(self, device: torch.device, grid, out_grid, n_modes_stored, max_n_modes, fft_norm: str):
        lib = _lib.load()
        prob = _lib.ScProblem()
        d = len(grid)
        if not 1 <= d <= _lib.SC_MAX_DIMS:
            raise NotImplementedError(f"SpectralConv supports 1..{_lib.SC_MAX_DIMS} spatial dims, got {d}")
        if fft_norm not in _lib.NORMS:
            raise ValueError(f"unknown fft_norm {fft_norm!r}")
        prob.ndim = d
        for j in range(d):
            prob.grid[j] = int(grid[j])
            prob.out_grid[j] = int(out_grid[j])
            prob.n_modes[j] = int(n_modes_stored[j])
            prob.max_n_modes[j] = int(max_n_modes[j])
        prob.fft_norm = _lib.NORMS[fft_norm]
        self._lib = lib
        self.device = device
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.sc_plan_create(ctypes.byref(prob), ctypes.byref(self.handle)), "sc_plan_create")
        kept = (ctypes.c_int32 * _lib.SC_MAX_DIMS)()
        lib.sc_plan_kept_modes(self.handle, kept)
        self.ndim = d
        self.grid = tuple(int(g) for g in grid)
        self.out_grid = tuple(int(g) for g in out_grid)
        self.kept = tuple(int(kept[j]) for j in range(d))
        self.n_modes_total = 1
        for k in self.kept:
            self.n_modes_total *= k

    def mode_bins(self, dim: int) -> Tuple[List[int], List[int]]:
        """(unshifted spectrum bins read, weight rows used) for kept slots of `dim` -- for index-set checks."""
        k = self.kept[dim]
        bins = (ctypes.c_int32 * k)()
        rows = (ctypes.c_int32 * k)()
        _lib.check(self._lib.sc_plan_mode_bins(self.handle, dim, bins, rows), "sc_plan_mode_bins")
        return list(bins), list(rows)

    def workspace_bytes(self, n_images: int) -> int:
        return int(self._lib.sc_workspace_bytes(self.handle, n_images))

    def set_fast_path(self, enable: bool):
        _lib.check(self._lib.sc_plan_set_fast_path(self.handle, int(bool(enable))), "sc_plan_set_fast_path")

    def uses_fast_path(self) -> int:
        return int(self._lib.sc_plan_uses_fast_path(self.handle))

    def __del__(self):
        try:
            if self.handle:
                self._lib.sc_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_PLAN_CACHE: "OrderedDict[tuple, Plan]" = OrderedDict()
_PLAN_LOCK = threading.Lock()
_PLAN_CACHE_MAX = 64


def get_plan(device: torch.device, grid, out_grid, n_modes_stored, max_n_modes, fft_norm="forward") -> Plan:
    if device.type != "cuda":
        raise RuntimeError("neuraloperator_b200.SpectralConv runs on CUDA (sm_100a) only; there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, tuple(grid), tuple(out_grid), tuple(n_modes_stored), tuple(max_n_modes), fft_norm)
    with _PLAN_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
            return plan
        plan = Plan(torch.device("cuda", idx), grid, out_grid, n_modes_stored, max_n_modes, fft_norm)
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
            _PLAN_CACHE.popitem(last=False)
        return plan


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _workspace(plan: Plan, n_images: int, device) -> torch.Tensor:
    return torch.empty(max(plan.workspace_bytes(n_images), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------------
# thin functional wrappers over the C ABI (also what the parity tests call)
# --------------------------------------------------------------------------------------------------
def analyze(plan: Plan, images: torch.Tensor, adjoint: bool = False) -> torch.Tensor:
    """images (n0, n1, *grid) float32 -> kept modes (n0, n1, *kept) complex64 (adjoint: images on out_grid)."""
    lib = _lib.load()
    spatial = plan.out_grid if adjoint else plan.grid
    assert images.dtype == torch.float32 and images.is_contiguous() and tuple(images.shape[2:]) == spatial
    n_images = images.shape[0] * images.shape[1]
    modes = torch.empty((*images.shape[:2], *plan.kept), dtype=torch.complex64, device=images.device)
    ws = _workspace(plan, n_images, images.device)
    with torch.cuda.device(images.device):
        _lib.check(lib.sc_analyze(plan.handle, _ptr(images), n_images, _ptr(modes), int(adjoint), _ptr(ws),
                                  ws.numel(), _stream_ptr(images.device)), "sc_analyze")
    return modes


def synthesize(plan: Plan, modes: torch.Tensor, bias: Optional[torch.Tensor] = None, adjoint: bool = False) -> torch.Tensor:
    lib = _lib.load()
    assert modes.dtype == torch.complex64 and modes.is_contiguous() and tuple(modes.shape[2:]) == plan.kept
    n_images = modes.shape[0] * modes.shape[1]
    spatial = plan.grid if adjoint else plan.out_grid
    out = torch.empty((*modes.shape[:2], *spatial), dtype=torch.float32, device=modes.device)
    ws = _workspace(plan, n_images, modes.device)
    b = bias.reshape(-1).contiguous() if bias is not None else None
    with torch.cuda.device(modes.device):
        _lib.check(lib.sc_synthesize(plan.handle, _ptr(modes), n_images, modes.shape[1], _ptr(b), _ptr(out),
                                     int(adjoint), _ptr(ws), ws.numel(), _stream_ptr(modes.device)), "sc_synthesize")
    return out


def contract_dense(plan: Plan, xm: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    B, Ci = xm.shape[:2]
    Co = weight.shape[1]
    assert weight.dtype == torch.complex64 and weight.is_contiguous() and weight.shape[0] == Ci
    ym = torch.empty((B, Co, *plan.kept), dtype=torch.complex64, device=xm.device)
    with torch.cuda.device(xm.device):
        _lib.check(lib.sc_contract_dense(plan.handle, _ptr(xm), _ptr(weight), _ptr(ym), B, Ci, Co,
                                         _stream_ptr(xm.device)), "sc_contract_dense")
    return ym


def contract_dense_backward(plan: Plan, xm, gm, weight, need_dxm=True, need_dweight=True, need_dbias=True):
    lib = _lib.load()
    B, Co = gm.shape[:2]
    Ci = weight.shape[0]
    dxm = torch.empty((B, Ci, *plan.kept), dtype=torch.complex64, device=gm.device) if need_dxm else None
    dw = torch.empty_like(weight) if need_dweight else None
    db = torch.empty(Co, dtype=torch.float32, device=gm.device) if need_dbias else None
    with torch.cuda.device(gm.device):
        _lib.check(lib.sc_contract_dense_backward(plan.handle, _ptr(xm), _ptr(gm), _ptr(weight), _ptr(dxm), _ptr(dw),
                                                  _ptr(db), B, Ci, Co, _stream_ptr(gm.device)),
                   "sc_contract_dense_backward")
    return dxm, dw, db


class _SpectralConvDense(torch.autograd.Function):
    """y = SpectralConv.forward(x) with a dense weight; saves only the kept input modes (B,Ci,*kept)."""

    @staticmethod
    def forward(ctx, x, weight, bias, plan: Plan, reducer=None):
        lib = _lib.load()
        B, Ci = x.shape[:2]
        Co = weight.shape[1]
        dev = x.device
        y = torch.empty((B, Co, *plan.out_grid), dtype=torch.float32, device=dev)
        xm = torch.empty((B, Ci, *plan.kept), dtype=torch.complex64, device=dev)
        n_max = B * max(Ci, Co)
        ws = _workspace(plan, n_max, dev)
        b = bias.reshape(-1) if bias is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.sc_forward_dense(plan.handle, _ptr(x), _ptr(weight), _ptr(b), _ptr(y), _ptr(xm), B, Ci, Co,
                                            _ptr(ws), ws.numel(), _stream_ptr(dev)), "sc_forward_dense")
        ctx.plan = plan
        ctx.reducer = reducer
        ctx.has_bias = bias is not None
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.save_for_backward(xm, weight)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        xm, weight = ctx.saved_tensors
        plan = ctx.plan
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        gy = gy.contiguous()
        if gy.dtype != torch.float32:
            gy = gy.float()
        B, Ci = xm.shape[:2]
        Co = weight.shape[1]
        dev = gy.device
        dx = torch.empty((B, Ci, *plan.grid), dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty_like(weight) if need_dw else None
        db = torch.empty(Co, dtype=torch.float32, device=dev) if need_db else None
        ws = _workspace(plan, B * max(Ci, Co), dev)
        reducer = ctx.reducer
        with torch.cuda.device(dev):
            if reducer is None or reducer.world_size() == 1:
                _lib.check(lib.sc_backward_dense(plan.handle, _ptr(gy), _ptr(weight), _ptr(xm), _ptr(dx), _ptr(dw), _ptr(db),
                                                 B, Ci, Co, _ptr(ws), ws.numel(), _stream_ptr(dev)), "sc_backward_dense")
            else:
                # data parallel: same kernels, but the gradient all-reduce of dweight / dbias is launched on the reducer's
                # side stream right after the contraction backward, underneath the dx synthesis kernel
                st = _stream_ptr(dev)
                gm = torch.empty((B, Co, *plan.kept), dtype=torch.complex64, device=dev)
                dxm = torch.empty((B, Ci, *plan.kept), dtype=torch.complex64, device=dev) if need_dx else None
                _lib.check(lib.sc_analyze(plan.handle, _ptr(gy), B * Co, _ptr(gm), 1, _ptr(ws), ws.numel(), st), "sc_analyze")
                _lib.check(lib.sc_contract_dense_backward(plan.handle, _ptr(xm), _ptr(gm), _ptr(weight), _ptr(dxm), _ptr(dw),
                                                          _ptr(db), B, Ci, Co, st), "sc_contract_dense_backward")
                reducer.start_early([dw, db])
                if need_dx:
                    _lib.check(lib.sc_synthesize(plan.handle, _ptr(dxm), B * Ci, 0, _ptr(None), _ptr(dx), 1, _ptr(ws),
                                                 ws.numel(), st), "sc_synthesize")
        if db is not None:
            db = db.reshape(ctx.bias_shape)
        return dx, dw, db, None, None


def spectral_conv_dense(x, weight, bias, plan: Plan, reducer=None):
    return _SpectralConvDense.apply(x, weight, bias, plan, reducer)


# --------------------------------------------------------------------------------------------------
# the module
# --------------------------------------------------------------------------------------------------
def _validate_scaling_factor(factor, n_dim) -> Optional[List[float]]:
    """Single-layer case of neuralop/utils.py:151-197 (`validate_scaling_factor(..., n_layers=None)`)."""
    if factor is None:
        return None
    if isinstance(factor, (int, float)):
        return [float(factor)] * n_dim
    if isinstance(factor, (list, tuple)) and len(factor) > 0 and all(isinstance(s, (int, float)) for s in factor):
        if len(factor) == n_dim:
            return [float(s) for s in factor]
        return [[float(s)] * n_dim for s in factor]
    return None


class BaseSpectralConv(nn.Module):
    """Plugin contract of the reference (`neuralop/layers/base_spectral_conv.py:4-27`)."""

    def __init__(self, device=None, dtype=None):
        super().__init__()
        self.dtype = dtype
        self.device = device

    def transform(self, x):
        return x


class SpectralConv(BaseSpectralConv):
    """Fourier-layer spectral convolution (real data, full precision) on hand-written sm_100a kernels.

    Parameters: identical to the reference class (spectral_convolution.py:183-305). Variants the kernels do
    not cover raise `NotImplementedError` at construction: `complex_data=True`, `separable=True`,
    `fno_block_precision != "full"`.
    """

    def __init__(
        self,
        in_channels,
        out_channels,
        n_modes,
        complex_data=False,
        max_n_modes=None,
        bias=True,
        separable=False,
        resolution_scaling_factor: Optional[Union[Number, List[Number]]] = None,
        fno_block_precision="full",
        rank=1.0,
        factorization=None,
        implementation="reconstructed",
        enforce_hermitian_symmetry=True,
        fixed_rank_modes=False,
        decomposition_kwargs: Optional[dict] = None,
        init_std="auto",
        fft_norm="forward",
        device=None,
    ):
        super().__init__(device=device)
        if complex_data:
            raise NotImplementedError("complex_data=True (C2C transforms) is not covered by the B200 kernels yet")
        if separable:
            raise NotImplementedError("separable=True is not covered by the B200 kernels yet")
        if fno_block_precision != "full":
            raise NotImplementedError("fno_block_precision must be 'full' (half/mixed spectral precision not built yet)")
        if implementation not in ("reconstructed", "factorized"):
            raise ValueError(f'Got implementation={implementation}, expected "reconstructed" or "factorized"')
        if fft_norm not in _lib.NORMS:
            raise ValueError(f"Got fft_norm={fft_norm}, expected one of {sorted(_lib.NORMS)}")

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.complex_data = complex_data
        self.n_modes = n_modes
        self.order = len(self.n_modes)
        if max_n_modes is None:
            max_n_modes = self.n_modes
        elif isinstance(max_n_modes, int):
            max_n_modes = [max_n_modes]
        self.max_n_modes = max_n_modes
        self.fno_block_precision = fno_block_precision
        self.rank = rank
        self.factorization = factorization
        self.implementation = implementation
        # the kernels always apply the Hermitian rules of the reference's default path (:547-559); with the flag
        # off the reference calls irfftn, whose C2R step ignores the same imaginary parts (identical on CPU)
        self.enforce_hermitian_symmetry = enforce_hermitian_symmetry
        self.resolution_scaling_factor = _validate_scaling_factor(resolution_scaling_factor, self.order)
        if init_std == "auto":
            init_std = (2 / (in_channels + out_channels)) ** 0.5
        if isinstance(fixed_rank_modes, bool):
            fixed_rank_modes = [0] if fixed_rank_modes else None
        self.fft_norm = fft_norm
        self.separable = separable
        # optional neuraloperator_b200.GradientAllReducer: backward then overlaps the dweight/dbias all-reduce with dx
        self.gradient_reducer = None

        weight_shape = (in_channels, out_channels, *self.max_n_modes)
        tensor_kwargs = decomposition_kwargs if decomposition_kwargs is not None else {}
        self.weight = FactorizedWeight.new(weight_shape, rank=self.rank, factorization=factorization or "Dense",
                                           fixed_rank_modes=fixed_rank_modes, dtype=torch.cfloat, device=device,
                                           **tensor_kwargs)
        self.weight.normal_(0, init_std)
        if bias:
            self.bias = nn.Parameter(init_std * torch.randn(*((self.out_channels,) + (1,) * self.order), device=device))
        else:
            self.bias = None

    # -- n_modes: stored with the last dim already halved, mutable at run time (:400-415) -------------
    @property
    def n_modes(self):
        return self._n_modes

    @n_modes.setter
    def n_modes(self, n_modes):
        n_modes = [n_modes] if isinstance(n_modes, int) else list(n_modes)
        if not self.complex_data:
            n_modes[-1] = n_modes[-1] // 2 + 1
        self._n_modes = n_modes

    def _output_grid(self, in_grid, output_shape):
        if output_shape is not None:
            return [int(s) for s in output_shape]
        if self.resolution_scaling_factor is not None:
            return [round(s * r) for s, r in zip(in_grid, self.resolution_scaling_factor)]
        return list(in_grid)

    def transform(self, x, output_shape=None):
        """Skip-connection transform (:383-398): identity unless the conv changes resolution."""
        in_shape = list(x.shape[2:])
        out_shape = self._output_grid(in_shape, output_shape)
        if in_shape == list(out_shape):
            return x
        raise NotImplementedError("SpectralConv.transform with a resolution change needs the reference's "
                                  "`resample` (neuralop/layers/resample.py), which is outside the spectral-conv path")

    def forward(self, x: torch.Tensor, output_shape: Optional[Tuple[int]] = None):
        if x.ndim != self.order + 2:
            raise ValueError(f"expected input of shape (batch, channels, {self.order} spatial dims), got {tuple(x.shape)}")
        if x.shape[1] != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}")
        if not x.is_cuda:
            raise RuntimeError("neuraloperator_b200.SpectralConv has no CPU path: move the module and input to a B200")
        if x.dtype != torch.float32:
            raise TypeError(f"SpectralConv (full precision, real data) expects float32 input, got {x.dtype}")
        grid = list(x.shape[2:])
        out_grid = self._output_grid(grid, output_shape)
        plan = get_plan(x.device, grid, out_grid, self.n_modes, self.max_n_modes, self.fft_norm)
        x = x.contiguous()
        # dense weight goes straight to the kernels; a factorized one is reconstructed first (differentiably)
        w = self.weight.to_tensor()
        if not w.is_contiguous():
            w = w.contiguous()
        return spectral_conv_dense(x, w, self.bias, plan, self.gradient_reducer if self.factorization is None else None)
