"""Complex spectral weights in dense or factorized (Tucker / CP / TT) form.

Host-side mirror of what the reference obtains from `tltorch.FactorizedTensor.new(...)`
(neuralop/layers/spectral_convolution.py:362-370): same factor layouts (pinned by the reference's einsum
strings, :63-68, :86-98, :117-127), same attribute names the reference reads (`.name`, `.core`, `.factors`,
`.weights`, `.to_tensor()`, `.shape`, slicing) and the parameter names tltorch registers
(`tensor` | `core`, `factors.factor_k` | `weights`, `factors.factor_k`), so state-dicts line up.

tltorch itself is an unpinned third-party dependency that is not part of the reference tree; its
rank-from-float rule and init-std split are restated from its published behaviour and
are not parity-pinned -- pass integer ranks for reproducible shapes.
"""
import math
from typing import List, Optional, Sequence

import torch
from torch import nn

_SYMS = "abcdefghijklmnopqrstuvwxyz"


def _complex_normal_(t: torch.Tensor, std: float):
    with torch.no_grad():
        re = torch.randn(t.shape, dtype=torch.float32, device=t.device)
        im = torch.randn(t.shape, dtype=torch.float32, device=t.device)
        t.copy_(torch.complex(re, im) * (std / math.sqrt(2.0)))
    return t


class FactorList(nn.Module):
    """Parameters registered as factor_0, factor_1, ... (tltorch's FactorList naming)."""

    def __init__(self, factors: Sequence[torch.Tensor]):
        super().__init__()
        self._n = len(factors)
        for i, f in enumerate(factors):
            self.register_parameter(f"factor_{i}", f if isinstance(f, nn.Parameter) else nn.Parameter(f))

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        return getattr(self, f"factor_{i}")

    def __iter__(self):
        return (self[i] for i in range(self._n))


def _bisect(fun, lo=0.0, hi=1.0):
    while fun(hi) < 0:
        hi *= 2.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if fun(mid) < 0:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def tucker_ranks(shape, rank, fixed_rank_modes=None) -> List[int]:
    if isinstance(rank, (list, tuple)):
        return [int(r) for r in rank]
    if isinstance(rank, int) and not isinstance(rank, bool):
        return [min(int(rank), s) for s in shape]
    if rank == "same":
        rank = 1.0
    fixed = set(fixed_rank_modes or [])
    target = float(rank) * math.prod(shape)
    free = [s for i, s in enumerate(shape) if i not in fixed]
    fixed_prod = math.prod(s for i, s in enumerate(shape) if i in fixed) if fixed else 1

    def count(f):
        return fixed_prod * math.prod(s * f for s in free) + sum(s * s * f for s in free) - target

    f = _bisect(count)
    return [s if i in fixed else max(int(round(s * f)), 1) for i, s in enumerate(shape)]


def cp_rank(shape, rank) -> int:
    if isinstance(rank, int) and not isinstance(rank, bool):
        return int(rank)
    if rank == "same":
        rank = 1.0
    return max(int(round(float(rank) * math.prod(shape) / sum(shape))), 1)


def tt_ranks(shape, rank) -> List[int]:
    n = len(shape)
    if isinstance(rank, (list, tuple)):
        r = [int(v) for v in rank]
        if len(r) != n + 1 or r[0] != 1 or r[-1] != 1:
            raise ValueError("TT rank list must have len(shape)+1 entries with boundary ranks 1")
        return r
    if isinstance(rank, int) and not isinstance(rank, bool):
        return [1] + [int(rank)] * (n - 1) + [1]
    if rank == "same":
        rank = 1.0
    target = float(rank) * math.prod(shape)
    if n == 1:
        return [1, 1]
    a = sum(shape[1:-1])
    b = shape[0] + shape[-1]
    r = (-b + math.sqrt(b * b + 4 * a * target)) / (2 * a) if a > 0 else target / b
    r = max(int(round(r)), 1)
    return [1] + [r] * (n - 1) + [1]


class FactorizedWeight(nn.Module):
    """Base class; `name` ends with the lowercase kind like tltorch's (`spectral_convolution.py:160-166`)."""

    kind = "base"

    @property
    def name(self):
        return "Complex" + self.kind.capitalize() if self.kind != "tt" else "ComplexTT"

    @staticmethod
    def new(shape, rank=1.0, factorization="Dense", fixed_rank_modes=None, dtype=torch.cfloat, device=None, **kw):
        if dtype not in (torch.cfloat, torch.complex64):
            raise NotImplementedError("spectral weights are complex64 (the reference creates them with dtype=cfloat)")
        kind = (factorization or "Dense").lower().replace("complex", "")
        shape = tuple(int(s) for s in shape)
        if kind == "dense":
            return DenseWeight(torch.empty(shape, dtype=dtype, device=device))
        if kind == "tucker":
            ranks = tucker_ranks(shape, rank, fixed_rank_modes)
            return TuckerWeight(torch.empty(ranks, dtype=dtype, device=device),
                                [torch.empty((s, r), dtype=dtype, device=device) for s, r in zip(shape, ranks)])
        if kind == "cp":
            r = cp_rank(shape, rank)
            return CPWeight(torch.ones(r, dtype=dtype, device=device),
                            [torch.empty((s, r), dtype=dtype, device=device) for s in shape])
        if kind == "tt":
            ranks = tt_ranks(shape, rank)
            return TTWeight([torch.empty((ranks[i], s, ranks[i + 1]), dtype=dtype, device=device)
                             for i, s in enumerate(shape)])
        raise ValueError(f"Got unexpected factorization {factorization!r}; expected Dense, Tucker, CP or TT")

    @staticmethod
    def from_tensor(tensor, rank=None, factorization="ComplexDense", **kw):
        kind = factorization.lower().replace("complex", "")
        if kind != "dense":
            raise NotImplementedError("from_tensor: only a dense weight can be built from a full tensor (no "
                                      "decomposition routine is part of the SpectralConv path)")
        return DenseWeight(tensor.detach().clone())

    # -- interface the reference reads ----------------------------------------------------------
    def to_tensor(self) -> torch.Tensor:
        raise NotImplementedError

    @property
    def shape(self):
        raise NotImplementedError

    @property
    def ndim(self):
        return len(self.shape)

    def is_complex(self):
        return True

    @property
    def dtype(self):
        return torch.cfloat

    def normal_(self, mean=0.0, std=1.0):
        raise NotImplementedError

    def decomposition(self) -> List[torch.Tensor]:
        raise NotImplementedError

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = [a.to_tensor() if isinstance(a, FactorizedWeight) else a for a in args]
        return func(*args, **kwargs)


class DenseWeight(FactorizedWeight):
    kind = "dense"

    def __init__(self, tensor):
        super().__init__()
        self.tensor = nn.Parameter(tensor)

    @property
    def shape(self):
        return self.tensor.shape

    def normal_(self, mean=0.0, std=1.0):
        _complex_normal_(self.tensor.data, std)
        return self

    def to_tensor(self):
        return self.tensor

    def decomposition(self):
        return [self.tensor]

    def __getitem__(self, idx):
        return self.tensor[idx]


class _View(FactorizedWeight):
    """Result of slicing a factorized weight: same kind, sliced factors, shared core (not registered)."""


def _slice_idx(idx, n):
    if not isinstance(idx, tuple):
        idx = (idx,)
    return list(idx) + [slice(None)] * (n - len(idx))


class TuckerWeight(FactorizedWeight):
    kind = "tucker"

    def __init__(self, core, factors, register=True):
        super().__init__()
        if register:
            self.core = nn.Parameter(core)
            self.factors = FactorList(factors)
        else:
            self.__dict__["core"] = core
            self.__dict__["factors"] = list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[0] for f in self.factors])

    @property
    def rank(self):
        return tuple(self.core.shape)

    def normal_(self, mean=0.0, std=1.0):
        r = math.prod(self.core.shape)
        std_f = (std / math.sqrt(r)) ** (1.0 / (len(self.factors) + 1))
        _complex_normal_(self.core.data, std_f)
        for f in self.factors:
            _complex_normal_(f.data, std_f)
        return self

    def to_tensor(self):
        n = len(self.factors)
        cs, os_ = _SYMS[:n], _SYMS[n:2 * n]
        eq = cs + "," + ",".join(o + c for o, c in zip(os_, cs)) + "->" + os_
        return torch.einsum(eq, self.core, *list(self.factors))

    def decomposition(self):
        return [self.core, *list(self.factors)]

    def __getitem__(self, idx):
        idx = _slice_idx(idx, len(self.factors))
        return TuckerWeight(self.core, [f[i, :] for f, i in zip(self.factors, idx)], register=False)


class CPWeight(FactorizedWeight):
    kind = "cp"

    def __init__(self, weights, factors, register=True):
        super().__init__()
        if register:
            self.weights = nn.Parameter(weights)
            self.factors = FactorList(factors)
        else:
            self.__dict__["weights"] = weights
            self.__dict__["factors"] = list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[0] for f in self.factors])

    @property
    def rank(self):
        return int(self.weights.shape[0])

    def normal_(self, mean=0.0, std=1.0):
        std_f = (std / math.sqrt(self.rank)) ** (1.0 / len(self.factors))
        with torch.no_grad():
            self.weights.fill_(1)
        for f in self.factors:
            _complex_normal_(f.data, std_f)
        return self

    def to_tensor(self):
        n = len(self.factors)
        eq = "z," + ",".join(s + "z" for s in _SYMS[:n]) + "->" + _SYMS[:n]
        return torch.einsum(eq, self.weights, *list(self.factors))

    def decomposition(self):
        return [self.weights, *list(self.factors)]

    def __getitem__(self, idx):
        idx = _slice_idx(idx, len(self.factors))
        return CPWeight(self.weights, [f[i, :] for f, i in zip(self.factors, idx)], register=False)


class TTWeight(FactorizedWeight):
    kind = "tt"

    def __init__(self, factors, register=True):
        super().__init__()
        if register:
            self.factors = FactorList(factors)
        else:
            self.__dict__["factors"] = list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[1] for f in self.factors])

    @property
    def rank(self):
        fs = list(self.factors)
        return tuple([f.shape[0] for f in fs] + [fs[-1].shape[2]])

    def normal_(self, mean=0.0, std=1.0):
        r = math.prod(f.shape[0] for f in self.factors)
        std_f = (std / math.sqrt(r)) ** (1.0 / len(self.factors))
        for f in self.factors:
            _complex_normal_(f.data, std_f)
        return self

    def to_tensor(self):
        fs = list(self.factors)
        out = fs[0]
        for f in fs[1:]:
            out = torch.tensordot(out, f, dims=([-1], [0]))
        return out.squeeze(0).squeeze(-1)

    def decomposition(self):
        return list(self.factors)

    def __getitem__(self, idx):
        idx = _slice_idx(idx, len(self.factors))
        return TTWeight([f[:, i, :] for f, i in zip(self.factors, idx)], register=False)
