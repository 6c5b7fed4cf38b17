"""Additive secret sharing over Z_{2^64} with fixed-point encoding."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch

PRECISION_FRACTIONAL = 3
BASE = 10 ** PRECISION_FRACTIONAL


def fix_precision(x: torch.Tensor, precision_fractional: int = PRECISION_FRACTIONAL) -> torch.Tensor:
    """``round(x * 10**p)`` as int64 (reference ``cf.py:265``, ``fc.py:438,444``).  CUDA tensors use the
    ``fix_precision_kernel`` of ``csrc/elementwise.cu``."""
    if x.is_cuda:
        from ..ops import _ext
        return _ext.require().fix_precision(x.contiguous().float(), float(10 ** precision_fractional))
    return torch.round(x.double() * (10 ** precision_fractional)).to(torch.int64)


def float_precision(x: torch.Tensor, precision_fractional: int = PRECISION_FRACTIONAL) -> torch.Tensor:
    if x.is_cuda:
        from ..ops import _ext
        return _ext.require().float_precision(x.contiguous(), float(10 ** precision_fractional))
    return x.to(torch.float64).div(10 ** precision_fractional).float()


def _rand_ring(shape, gen: torch.Generator, device=None) -> torch.Tensor:
    r = torch.randint(-(2 ** 62), 2 ** 62, tuple(shape), dtype=torch.int64, generator=gen)
    return r.to(device) if device is not None else r


def ring_matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """int64 matmul with wrap-around (== arithmetic mod 2^64).  torch has no integer GEMM on CUDA; CUDA
    tensors use the tiled ``ring_matmul_kernel``."""
    if a.is_cuda and a.dim() == 2 and b.dim() == 2:
        from ..ops import _ext
        return _ext.require().ring_matmul(a.contiguous(), b.contiguous())
    return (a.unsqueeze(-1) * b.unsqueeze(-3)).sum(-2) if a.dim() > 1 else a @ b


class CryptoProvider:
    """Deals Beaver triples and helps with comparisons (the reference's ``crypto_provider``
    VirtualWorker, ``fc.py:428``).  Counts what it dealt so tests can assert on traffic."""

    def __init__(self, seed: int = 0) -> None:
        self.gen = torch.Generator().manual_seed(seed)
        self.triples_dealt = 0
        self.comparisons = 0

    def matmul_triple(self, a_shape, b_shape, device=None) -> Tuple["SharedTensor", "SharedTensor", "SharedTensor"]:
        a, b = _rand_ring(a_shape, self.gen, device), _rand_ring(b_shape, self.gen, device)
        c = ring_matmul(a, b)
        self.triples_dealt += 1
        return share(a, self), share(b, self), share(c, self)

    def mul_triple(self, shape, device=None) -> Tuple["SharedTensor", "SharedTensor", "SharedTensor"]:
        a, b = _rand_ring(shape, self.gen, device), _rand_ring(shape, self.gen, device)
        self.triples_dealt += 1
        return share(a, self), share(b, self), share(a * b, self)

    def positive_bit(self, blinded0: torch.Tensor, blinded1: torch.Tensor) -> "SharedTensor":
        """Shares of ``[x > 0]`` from the two workers' blinded shares of ``t*x`` (t > 0 unknown here)."""
        self.comparisons += 1
        return share(((blinded0 + blinded1) > 0).to(torch.int64), self)


@dataclass
class SharedTensor:
    """Two additive shares; ``shares[0] + shares[1] == secret (mod 2^64)``.  In a deployment each
    share lives on a different worker; here both sit in one process like PySyft VirtualWorkers."""

    shares: List[torch.Tensor]
    provider: CryptoProvider
    scale: int = 1  # BASE**k bookkeeping is explicit: values carry one factor of BASE unless noted

    @property
    def shape(self):
        return self.shares[0].shape

    def get(self) -> torch.Tensor:
        return self.shares[0] + self.shares[1]

    def refresh(self) -> "SharedTensor":
        r = _rand_ring(self.shape, self.provider.gen, self.shares[0].device)
        return SharedTensor([self.shares[0] + r, self.shares[1] - r], self.provider)

    # -- linear ops are local -----------------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, SharedTensor):
            return SharedTensor([self.shares[0] + other.shares[0], self.shares[1] + other.shares[1]], self.provider)
        return SharedTensor([self.shares[0] + other, self.shares[1]], self.provider)

    def __sub__(self, other):
        if isinstance(other, SharedTensor):
            return SharedTensor([self.shares[0] - other.shares[0], self.shares[1] - other.shares[1]], self.provider)
        return SharedTensor([self.shares[0] - other, self.shares[1]], self.provider)

    def __neg__(self):
        return SharedTensor([-self.shares[0], -self.shares[1]], self.provider)

    def mul_public(self, c) -> "SharedTensor":
        return SharedTensor([self.shares[0] * c, self.shares[1] * c], self.provider)

    def t(self) -> "SharedTensor":
        return SharedTensor([self.shares[0].t(), self.shares[1].t()], self.provider)

    def sum(self, dim=None) -> "SharedTensor":
        f = (lambda s: s.sum()) if dim is None else (lambda s: s.sum(dim))
        return SharedTensor([f(self.shares[0]), f(self.shares[1])], self.provider)

    def truncate(self, divisor: int = BASE) -> "SharedTensor":
        """SecureML local truncation: error <= 1 ulp with overwhelming probability."""
        s0 = torch.div(self.shares[0], divisor, rounding_mode="floor")
        s1 = -torch.div(-self.shares[1], divisor, rounding_mode="floor")
        return SharedTensor([s0, s1], self.provider)

    # -- Beaver multiplication --------------------------------------------------------------------
    def matmul(self, other: "SharedTensor") -> "SharedTensor":
        a, b, c = self.provider.matmul_triple(self.shape, other.shape, self.shares[0].device)
        d = (self - a).get()   # opened
        e = (other - b).get()  # opened
        z0 = c.shares[0] + ring_matmul(d, b.shares[0]) + ring_matmul(a.shares[0], e) + ring_matmul(d, e)
        z1 = c.shares[1] + ring_matmul(d, b.shares[1]) + ring_matmul(a.shares[1], e)
        return SharedTensor([z0, z1], self.provider)

    def mul(self, other: "SharedTensor") -> "SharedTensor":
        a, b, c = self.provider.mul_triple(self.shape, self.shares[0].device)
        d, e = (self - a).get(), (other - b).get()
        z0 = c.shares[0] + d * b.shares[0] + a.shares[0] * e + d * e
        z1 = c.shares[1] + d * b.shares[1] + a.shares[1] * e
        return SharedTensor([z0, z1], self.provider)

    def positive_bit(self) -> "SharedTensor":
        """Shares of the 0/1 indicator ``[x > 0]`` (helper-aided, see package docstring)."""
        t = int(torch.randint(1, 2 ** 16, (1,), generator=self.provider.gen))  # common secret of the 2 workers
        return self.provider.positive_bit(self.shares[0] * t, self.shares[1] * t)


def share(secret: torch.Tensor, provider) -> "SharedTensor":
    if hasattr(provider, "deal"):     # smpc.dist.PartyContext: the dealer's value is shared out, the other ranks pass the shape
        return provider.share(secret if provider.is_dealer else None, tuple(secret.shape))
    r = _rand_ring(secret.shape, provider.gen, secret.device)
    return SharedTensor([r, secret - r], provider)
