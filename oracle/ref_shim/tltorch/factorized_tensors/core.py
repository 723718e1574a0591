"""Minimal stand-in for `tltorch.factorized_tensors.core.FactorizedTensor`.

TEST INFRASTRUCTURE ONLY (tensorly-torch is an unpinned third-party dependency of the reference,
`pyproject.toml:12-13`, absent from this image). It restates only what the reference call sites
need (`spectral_convolution.py:362-370,489,39-40,71-73,101-103,130-132,160-166`):

* factor layout: Tucker core `(r_0..r_n)`, factors `(dim_k, r_k)`; CP `weights (R,)`, factors
  `(dim_k, R)`; TT cores `(r_k, dim_k, r_{k+1})` -- pinned by the reference einsum strings
  (`spectral_convolution.py:63-68,86-98,117-127`).
* indexing with slices returns the same kind with sliced factors (Tucker/CP) -- what
  `self.weight[slices_w]` relies on.

Rank-from-float and the init-std split across factors are NOT pinned by anything in the reference
tree; this shim takes explicit integer ranks (a float `rank` falls back to the tensorly rule restated
in SURVEY.md App. B) and the golden vectors carry the factor tensors explicitly.
"""
import math

import torch
from torch import nn


def _tucker_ranks_from_float(shape, rank, fixed_rank_modes=None):
    if isinstance(rank, (list, tuple)):
        return [int(r) for r in rank]
    if isinstance(rank, int):
        return [min(rank, s) for s in shape]
    fixed = set(fixed_rank_modes or [])
    free = [s for i, s in enumerate(shape) if i not in fixed]
    n_param = rank * math.prod(shape)
    fixed_prod = math.prod(s for i, s in enumerate(shape) if i in fixed) if fixed else 1
    # solve fixed_prod * prod(free) * f^n + sum(free s^2) f  [+ fixed s^2] = n_param by bisection
    lo, hi = 0.0, 1.0

    def count(f):
        return fixed_prod * math.prod(s * f for s in free) + sum(s * s * f for s in free)

    while count(hi) < n_param:
        hi *= 2
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if count(mid) < n_param:
            lo = mid
        else:
            hi = mid
    f = 0.5 * (lo + hi)
    return [s if i in fixed else max(int(round(s * f)), 1) for i, s in enumerate(shape)]


class FactorizedTensor(nn.Module):
    _name = "FactorizedTensor"

    @property
    def name(self):
        return self._name

    @classmethod
    def new(cls, shape, rank="same", factorization="Dense", fixed_rank_modes=None,
            dtype=None, device=None, **kwargs):
        kind = factorization.lower().replace("complex", "")
        shape = tuple(int(s) for s in shape)
        if kind == "dense":
            return DenseTensor(torch.empty(shape, dtype=dtype, device=device))
        if kind == "tucker":
            ranks = _tucker_ranks_from_float(shape, rank, fixed_rank_modes)
            core = torch.empty(ranks, dtype=dtype, device=device)
            factors = [torch.empty((s, r), dtype=dtype, device=device) for s, r in zip(shape, ranks)]
            return TuckerTensor(core, factors)
        if kind == "cp":
            if isinstance(rank, float):
                rank = max(int(round(rank * math.prod(shape) / sum(shape))), 1)
            weights = torch.ones(int(rank), dtype=dtype, device=device)
            factors = [torch.empty((s, int(rank)), dtype=dtype, device=device) for s in shape]
            return CPTensor(weights, factors)
        if kind == "tt":
            if isinstance(rank, (int, float)):
                r = int(rank) if isinstance(rank, int) else max(int(round(rank * min(shape))), 1)
                ranks = [1] + [r] * (len(shape) - 1) + [1]
            else:
                ranks = list(rank)
            factors = [torch.empty((ranks[i], s, ranks[i + 1]), dtype=dtype, device=device)
                       for i, s in enumerate(shape)]
            return TTTensor(factors)
        raise ValueError(f"unknown factorization {factorization}")

    @classmethod
    def from_tensor(cls, tensor, rank="same", factorization="Dense", **kwargs):
        kind = factorization.lower().replace("complex", "")
        if kind != "dense":
            raise NotImplementedError("shim: from_tensor only for Dense")
        return DenseTensor(tensor.detach().clone())

    def normal_(self, mean=0.0, std=1.0):
        raise NotImplementedError

    def to_tensor(self):
        raise NotImplementedError

    @property
    def shape(self):
        raise NotImplementedError

    def is_complex(self):
        return True

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = [a.to_tensor() if isinstance(a, FactorizedTensor) else a for a in args]
        return func(*args, **kwargs)


def _cnormal_(t, std):
    with torch.no_grad():
        if t.is_complex():
            t.copy_(torch.randn(t.shape, dtype=t.dtype) * std)  # complex normal: var split re/im
        else:
            t.normal_(0, std)
    return t


class DenseTensor(FactorizedTensor):
    _name = "Dense"

    def __init__(self, tensor):
        super().__init__()
        self.tensor = nn.Parameter(tensor)

    @property
    def shape(self):
        return self.tensor.shape

    def normal_(self, mean=0.0, std=1.0):
        _cnormal_(self.tensor.data, std)
        return self

    def to_tensor(self):
        return self.tensor

    def __getitem__(self, idx):
        return self.tensor[idx]


class TuckerTensor(FactorizedTensor):
    _name = "Tucker"

    def __init__(self, core, factors, as_param=True):
        super().__init__()
        if as_param:
            self.core = nn.Parameter(core)
            self.factors = nn.ParameterList([nn.Parameter(f) for f in factors])
        else:
            self.core, self.factors = core, list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[0] for f in self.factors])

    def normal_(self, mean=0.0, std=1.0):
        r = math.prod(self.core.shape)
        std_f = (std / math.sqrt(r)) ** (1.0 / (len(self.factors) + 1))
        _cnormal_(self.core.data, std_f)
        for f in self.factors:
            _cnormal_(f.data, std_f)
        return self

    def to_tensor(self):
        syms = "abcdefghij"
        n = len(self.factors)
        core_s = syms[:n]
        out_s = syms[n:2 * n]
        eq = core_s + "," + ",".join(o + c for o, c in zip(out_s, core_s)) + "->" + out_s
        return torch.einsum(eq, self.core, *self.factors)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = list(idx) + [slice(None)] * (len(self.factors) - len(idx))
        return TuckerTensor(self.core, [f[i, :] for f, i in zip(self.factors, idx)], as_param=False)


class CPTensor(FactorizedTensor):
    _name = "CP"

    def __init__(self, weights, factors, as_param=True):
        super().__init__()
        if as_param:
            self.weights = nn.Parameter(weights)
            self.factors = nn.ParameterList([nn.Parameter(f) for f in factors])
        else:
            self.weights, self.factors = weights, list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[0] for f in self.factors])

    def normal_(self, mean=0.0, std=1.0):
        rank = self.weights.shape[0]
        std_f = (std / math.sqrt(rank)) ** (1.0 / len(self.factors))
        with torch.no_grad():
            self.weights.data.fill_(1)
        for f in self.factors:
            _cnormal_(f.data, std_f)
        return self

    def to_tensor(self):
        syms = "abcdefghij"
        n = len(self.factors)
        eq = "z," + ",".join(s + "z" for s in syms[:n]) + "->" + syms[:n]
        return torch.einsum(eq, self.weights, *self.factors)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = list(idx) + [slice(None)] * (len(self.factors) - len(idx))
        return CPTensor(self.weights, [f[i, :] for f, i in zip(self.factors, idx)], as_param=False)


class TTTensor(FactorizedTensor):
    _name = "TT"

    def __init__(self, factors, as_param=True):
        super().__init__()
        if as_param:
            self.factors = nn.ParameterList([nn.Parameter(f) for f in factors])
        else:
            self.factors = list(factors)

    @property
    def shape(self):
        return torch.Size([f.shape[1] for f in self.factors])

    def normal_(self, mean=0.0, std=1.0):
        r = math.prod(f.shape[0] for f in self.factors)
        std_f = (std / math.sqrt(r)) ** (1.0 / len(self.factors))
        for f in self.factors:
            _cnormal_(f.data, std_f)
        return self

    def to_tensor(self):
        out = self.factors[0]
        for f in self.factors[1:]:
            out = torch.tensordot(out, f, dims=([-1], [0]))
        return out.squeeze(0).squeeze(-1)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = list(idx) + [slice(None)] * (len(self.factors) - len(idx))
        return TTTensor([f[:, i, :] for f, i in zip(self.factors, idx)], as_param=False)
