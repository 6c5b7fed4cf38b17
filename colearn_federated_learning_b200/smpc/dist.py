"""The encrypted demo across the ranks of a process group (secret-share scatter + per-multiplication opens).

The reference runs its SMPC demo on two ``VirtualWorker`` share holders plus a ``crypto_provider`` worker
(``federated_coordinator.py:428-438``).  In box mode those three are RANKS: the **dealer** (the coordinator, which owns the
model and the data and also plays the crypto provider: it deals the Beaver triples and helps with comparisons) and two
**parties** that hold one additive share of every tensor each and never see a secret.  With only two ranks the dealer shares
rank 0 with party 0 (documented: then only party 1 is blind).

Every rank executes the same program (:class:`~.mlp.SharedMLP` runs unchanged on :class:`DistShared` tensors); what differs
per role is which share it holds:

* ``ctx.share(secret)``   dealer: ``r`` -> party 0, ``secret - r`` -> party 1 (point-to-point sends);
* linear ops              local on the parties (public constants are added by party 0 only);
* ``matmul`` / ``mul``    the dealer deals a Beaver triple, the parties **open** ``x - a`` and ``y - b`` — the one operation that
                          moves data between the two parties.  On GPUs the exchange is this repo's own transport: each party
                          writes its share straight into the peer's symmetric buffer with ``p2p_copy_kernel`` (NVLink store +
                          ``st.release.sys`` flag), waits for the peer's flag and adds; elsewhere ``dist.all_reduce`` over the
                          pair's subgroup;
* ``positive_bit``        the parties blind their shares with a common positive scalar and send them to the dealer, which
                          deals shares of the sign bit (the helper-aided comparison of :mod:`.sharing`);
* ``get()``               both shares travel to the dealer, which reconstructs (model reveal, loss logging); the parties learn nothing.

Arithmetic is int64 with wrap-around (the ring Z_2^64) on every path, fixed-point with 3 fractional digits like the reference.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .sharing import BASE, _rand_ring, ring_matmul


class PartyContext:
    """Roles + transport of the distributed demo.  Duck-types :class:`~.sharing.CryptoProvider` (``share``, counters)."""

    def __init__(self, device: torch.device, *, dealer: int = 0, parties: Optional[Sequence[int]] = None, seed: int = 0,
                 group: Optional[dist.ProcessGroup] = None, p2p: Optional[bool] = None, xchg_elems: int = 1 << 16) -> None:
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world < 2:
            raise ValueError("the distributed encrypted demo needs at least 2 ranks (2 share holders; the dealer may share rank 0)")
        if parties is None:
            parties = (1, 2) if self.world >= 3 else (0, 1)
        self.dealer, (self.p0, self.p1) = int(dealer), (int(parties[0]), int(parties[1]))
        assert self.p0 != self.p1
        self.device = torch.device(device)
        self.is_dealer = self.rank == self.dealer
        self.party: Optional[int] = 0 if self.rank == self.p0 else (1 if self.rank == self.p1 else None)
        self.gen = torch.Generator().manual_seed(seed)                 # the dealer's randomness (triples, sharing masks)
        self.common = torch.Generator().manual_seed(seed * 7919 + 13)  # the two parties' common randomness (refresh, blinding)
        self.triples_dealt = 0
        self.comparisons = 0
        self.opens = 0
        self.bytes_between_parties = 0
        self.pair = dist.new_group([self.p0, self.p1]) if self.world > 2 else group     # every rank must call new_group
        # own transport for the opens on GPUs: symmetric exchange buffers + p2p_copy_kernel + release/acquire flags
        if p2p is None:
            p2p = self.device.type == "cuda" and os.environ.get("COLEARN_SMPC_P2P", "1") != "0"
        self.p2p = bool(p2p)
        self.xchg_elems = int(xchg_elems)
        self._opens_p2p = 0
        if self.p2p:
            from .. import ops
            from ..parallel.symm import SymmetricArena
            self.ext = ops._ext.require()
            # two slots of (recv int64 [xchg_elems]) per rank: the peer may be one open ahead, never two (see open())
            self.slot_stride = 2 * self.xchg_elems + 4          # fp32 words per slot (+ the float4 round-up of an odd element count)
            self.arena = SymmetricArena({"recv": (2 * self.slot_stride, torch.float32), "flags": (8, torch.int32)}, self.device, group)
            self.stage = torch.zeros(self.xchg_elems + 2, dtype=torch.int64, device=self.device)
            self.ext.set_spin_limit(float(os.environ.get("COLEARN_SPIN_TIMEOUT_S", "120")))

    # -- point-to-point helpers (int64) ---------------------------------------------------------------------------------
    def _send(self, t: torch.Tensor, dst: int) -> None:
        dist.send(t.contiguous().to(self.device), dst, group=self.group)

    def _recv(self, shape, src: int) -> torch.Tensor:
        buf = torch.empty(tuple(shape), dtype=torch.int64, device=self.device)
        dist.recv(buf, src, group=self.group)
        return buf

    def deal(self, secrets: Optional[List[torch.Tensor]], shapes: Sequence[Sequence[int]]) -> List[Optional[torch.Tensor]]:
        """Dealer: split every secret into two additive shares and hand them out.  Returns this rank's shares (None on a
        pure dealer).  ``shapes`` lets the receiving ranks size their buffers."""
        mine: List[Optional[torch.Tensor]] = [None] * len(shapes)
        if self.is_dealer:
            assert secrets is not None and len(secrets) == len(shapes)
            for i, s in enumerate(secrets):
                s = s.to(self.device)
                r = _rand_ring(s.shape, self.gen, self.device)
                for party, sh in ((self.p0, r), (self.p1, s - r)):
                    if party == self.rank:
                        mine[i] = sh
                    else:
                        self._send(sh, party)
        if self.party is not None and not self.is_dealer:
            for i, shp in enumerate(shapes):
                mine[i] = self._recv(shp, self.dealer)
        return mine

    def share(self, secret: Optional[torch.Tensor], shape: Optional[Sequence[int]] = None) -> "DistShared":
        """``.fix_precision().share(w1, w2, crypto_provider)``: the dealer passes the secret, the others its shape."""
        shape = tuple(secret.shape) if secret is not None else tuple(shape)
        return DistShared(self.deal([secret] if self.is_dealer else None, [shape])[0], self, shape)

    # -- the one party-to-party operation ------------------------------------------------------------------------------------
    def open(self, sh: Optional[torch.Tensor], shape) -> Optional[torch.Tensor]:
        """Both parties learn ``share_0 + share_1``; a pure dealer gets None."""
        if self.party is None:
            return None
        self.opens += 1
        self.bytes_between_parties += 8 * sh.numel()
        n = sh.numel()
        if self.p2p and 0 < n <= self.xchg_elems:
            # slot k & 1 of the peer's buffer <- my share (NVLink store from inside p2p_copy_kernel, flag raised with
            # st.release.sys by its last CTA); wait for the peer's flag of this open; add.  Two slots suffice: to start open
            # k + 2 the peer needs my data of open k + 1, which I send only after my add of open k was queued on this stream.
            self._opens_p2p += 1
            k = self._opens_p2p
            slot = k & 1
            peer = self.p1 if self.party == 0 else self.p0
            self.stage[:n].copy_(sh.reshape(-1))
            n_f = (2 * n + 3) // 4 * 4                                  # int64 elements as fp32 words, whole float4s
            self.ext.p2p_copy(self.arena.ptr("recv", peer, slot * self.slot_stride), self.stage.data_ptr(), n_f,
                              self.arena.ptr("flags", peer, slot), k, 16)
            self.ext.wait_flags(self.arena.ptr("flags", None, slot), 1, k)
            got = self.arena.tensor("recv")[slot * self.slot_stride: slot * self.slot_stride + 2 * n].view(torch.int64)
            return (sh.reshape(-1) + got).view(tuple(shape))
        out = sh.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.pair)
        return out


class DistShared:
    """One rank's view of an additively shared int64 tensor: ``share`` is this party's share (None on a pure dealer)."""

    def __init__(self, share: Optional[torch.Tensor], ctx: PartyContext, shape) -> None:
        self.share, self.ctx, self._shape = share, ctx, tuple(shape)

    @property
    def shape(self):
        return torch.Size(self._shape)

    @property
    def provider(self) -> PartyContext:
        return self.ctx

    def _new(self, share, shape=None) -> "DistShared":
        return DistShared(share, self.ctx, self._shape if shape is None else shape)

    def _local(self, f, shape=None) -> "DistShared":
        return self._new(f(self.share) if self.share is not None else None, shape)

    # -- reconstruction -----------------------------------------------------------------------------------------------------
    def get(self) -> Optional[torch.Tensor]:
        """``.get()``: both shares travel to the dealer (the coordinator), which reconstructs; the parties learn nothing and
        return None."""
        ctx = self.ctx
        if ctx.is_dealer:
            parts = []
            for party, rank in ((0, ctx.p0), (1, ctx.p1)):
                parts.append(self.share if ctx.party == party else ctx._recv(self._shape, rank))
            return parts[0] + parts[1]
        if ctx.party is not None:
            ctx._send(self.share, ctx.dealer)
        return None

    def refresh(self) -> "DistShared":
        r = _rand_ring(self._shape, self.ctx.common)          # drawn by everybody to keep the generators in step
        if self.share is None:
            return self
        r = r.to(self.share.device)
        return self._new(self.share + r if self.ctx.party == 0 else self.share - r)

    # -- linear ops are local -------------------------------------------------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, DistShared):
            return self._new(self.share + other.share if self.share is not None else None)
        return self._new(self.share + other if (self.share is not None and self.ctx.party == 0) else self.share)

    def __sub__(self, other):
        if isinstance(other, DistShared):
            return self._new(self.share - other.share if self.share is not None else None)
        return self._new(self.share - other if (self.share is not None and self.ctx.party == 0) else self.share)

    def __neg__(self):
        return self._local(lambda s: -s)

    def mul_public(self, c) -> "DistShared":
        return self._local(lambda s: s * c)

    def t(self) -> "DistShared":
        return self._local(lambda s: s.t(), tuple(reversed(self._shape)))

    def sum(self, dim=None) -> "DistShared":
        shape = () if dim is None else tuple(d for i, d in enumerate(self._shape) if i != (dim % len(self._shape)))
        return self._local((lambda s: s.sum()) if dim is None else (lambda s: s.sum(dim)), shape)

    def truncate(self, divisor: int = BASE) -> "DistShared":
        """SecureML local truncation (party 0 floors its share, party 1 floors the negation)."""
        if self.share is None:
            return self
        if self.ctx.party == 0:
            return self._new(torch.div(self.share, divisor, rounding_mode="floor"))
        return self._new(-torch.div(-self.share, divisor, rounding_mode="floor"))

    # -- Beaver multiplication ------------------------------------------------------------------------------------------------
    def _beaver(self, other: "DistShared", matmul: bool) -> "DistShared":
        ctx = self.ctx
        if matmul:
            out_shape = tuple(self._shape[:-1]) + tuple(other._shape[1:])
        else:
            out_shape = self._shape
        secrets = None
        if ctx.is_dealer:
            a, b = _rand_ring(self._shape, ctx.gen, ctx.device), _rand_ring(other._shape, ctx.gen, ctx.device)
            secrets = [a, b, ring_matmul(a, b) if matmul else a * b]
        ctx.triples_dealt += 1
        a_s, b_s, c_s = ctx.deal(secrets, [self._shape, other._shape, out_shape])
        if ctx.party is None:
            return self._new(None, out_shape)
        d = ctx.open(self.share - a_s, self._shape)
        e = ctx.open(other.share - b_s, other._shape)
        op = ring_matmul if matmul else (lambda x, y: x * y)
        z = c_s + op(d, b_s) + op(a_s, e)
        if ctx.party == 0:
            z = z + op(d, e)
        return self._new(z, out_shape)

    def matmul(self, other: "DistShared") -> "DistShared":
        return self._beaver(other, True)

    def mul(self, other: "DistShared") -> "DistShared":
        return self._beaver(other, False)

    def positive_bit(self) -> "DistShared":
        """Shares of ``[x > 0]``: the parties blind their shares with a common positive scalar, the dealer sees only that."""
        ctx = self.ctx
        t = int(torch.randint(1, 2 ** 16, (1,), generator=ctx.common))
        ctx.comparisons += 1
        blinded: List[Optional[torch.Tensor]] = [None, None]
        for party, rank in ((0, ctx.p0), (1, ctx.p1)):
            if ctx.party == party:
                if ctx.is_dealer:
                    blinded[party] = self.share * t
                else:
                    ctx._send(self.share * t, ctx.dealer)
            elif ctx.is_dealer:
                blinded[party] = ctx._recv(self._shape, rank)
        bit = ((blinded[0] + blinded[1]) > 0).to(torch.int64) if ctx.is_dealer else None
        return DistShared(ctx.deal([bit] if ctx.is_dealer else None, [self._shape])[0], ctx, self._shape)


def train_encrypted_dist(model, x: Optional[torch.Tensor], y: Optional[torch.Tensor], n_items: int, in_dim: int, out_dim: int, args,
                         ctx: PartyContext) -> Tuple[float, dict]:
    """``starting_training_enc`` (fc.py:394-472) on a process group: the dealer passes the (already shuffled) first
    ``n_items`` samples, the other ranks only their shapes.  Returns (last logged loss, counters); the dealer's ``model`` holds
    the decrypted parameters afterwards."""
    import time

    from .mlp import SharedMLP
    from .sharing import fix_precision, float_precision

    bsz = max(1, int(args.batch_size))
    n_batches = int(n_items / bsz)
    loader = []
    for i in range(n_batches):
        lo, hi = i * bsz, min((i + 1) * bsz, n_items)
        if hi <= lo:
            break
        if ctx.is_dealer:
            xb, yb = fix_precision(x[lo:hi], args.precision_fractional), fix_precision(y[lo:hi].view(hi - lo, -1), args.precision_fractional)
        else:
            xb = yb = None
        loader.append((ctx.share(xb, (hi - lo, in_dim)), ctx.share(yb, (hi - lo, out_dim))))
    shared = SharedMLP.from_module(model, ctx)
    t0 = time.time()
    last = 0.0
    for epoch in range(args.epochs):
        for batch_idx, (data, target) in enumerate(loader):
            loss = shared.step(data, target, args.lr)
            if batch_idx % args.log_interval == 0:
                v = loss.get()
                if ctx.is_dealer:
                    last = float(float_precision(v))
                    print("Train Epoch: {} [{}/{} ({:.0f}%)]\tLoss: {:.6f}\tTime: {:.3f}s".format(
                        epoch, batch_idx * bsz, len(loader) * bsz, 100.0 * batch_idx / max(1, len(loader)), last, time.time() - t0))
    shared.reveal_into(model)
    return last, {"batches": len(loader), "seconds": time.time() - t0, "triples": ctx.triples_dealt, "comparisons": ctx.comparisons,
                  "opens": ctx.opens, "bytes_between_parties": ctx.bytes_between_parties, "p2p_opens": ctx._opens_p2p,
                  "dealer": ctx.dealer, "parties": [ctx.p0, ctx.p1]}
