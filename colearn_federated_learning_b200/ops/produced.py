"""Producer-side reports of the fused  wgrad GEMM → FedAvg reduce  (csrc/produced.cuh, docs/KERNELS.md §3b).

During the last local SGD step of a round every writer of final parameters reports the arena elements it finished;
when a chunk is complete its epoch is published in the chunk owner's table, which the overlapped two-shot kernel
(``twoshot_overlap_kernel``) polls while the rest of the backward pass is still running.  :class:`ProducedSpec` carries
what the writers need:

* device form (CUDA extension, or the SIMT-on-CPU build of the same sources in tests): a ``ProducedSignal`` struct in
  device memory; the GEMM epilogue reports per warp block, :meth:`mark` launches ``produced_mark_kernel``;
* reference form (plain CPU tensors, no extension): the same counting in Python, so that the host-side bookkeeping —
  every element reported exactly once — is testable anywhere.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class ProducedSpec:
    def __init__(self, *, chunk_elems: int, n: int, world: int, rank: int, max_ctas: int = 0) -> None:
        assert chunk_elems >= 4 and chunk_elems & (chunk_elems - 1) == 0, "chunk_elems must be a power of two"
        self.chunk_elems, self.n, self.world, self.rank, self.max_ctas = int(chunk_elems), int(n), int(world), int(rank), int(max_ctas)
        self.n_chunks = (self.n + self.chunk_elems - 1) // self.chunk_elems
        self.mod = None
        self.sig: Optional[torch.Tensor] = None
        self.count: Optional[torch.Tensor] = None
        self.tables: Optional[torch.Tensor] = None
        self.epoch: Optional[torch.Tensor] = None
        self.epoch_add = 0

    # -- constructors ------------------------------------------------------------------------------------------
    @classmethod
    def device(cls, mod, count: torch.Tensor, flag_ptrs: Sequence[int], epoch: torch.Tensor, epoch_add: int, *, chunk_elems: int,
               n: int, rank: int, max_ctas: int = 0) -> "ProducedSpec":
        """``count``: int32 ``[n_chunks]`` zeroed scratch on the kernels' device; ``flag_ptrs[o]``: address of rank ``o``'s
        ``[world, n_chunks]`` uint32 table; ``epoch``: int32 tensor holding the value (minus ``epoch_add``) to publish."""
        sp = cls(chunk_elems=chunk_elems, n=n, world=len(flag_ptrs), rank=rank, max_ctas=max_ctas)
        assert count.dtype == torch.int32 and count.numel() >= sp.n_chunks and count.is_contiguous()
        assert epoch.dtype == torch.int32 and epoch.device == count.device
        sp.mod, sp.count, sp.epoch, sp.epoch_add = mod, count, epoch, int(epoch_add)
        packed = mod.produced_signal_pack(count.data_ptr(), [int(p) for p in flag_ptrs], epoch.data_ptr(), int(epoch_add),
                                          sp.chunk_elems, sp.rank, sp.n_chunks, sp.n)
        sp.sig = packed.to(count.device)
        return sp

    @classmethod
    def reference(cls, tables: torch.Tensor, epoch: torch.Tensor, epoch_add: int, *, chunk_elems: int, n: int, rank: int) -> "ProducedSpec":
        """``tables``: int32 ``[world (owner), world (producer), n_chunks]`` shared by the emulated ranks."""
        sp = cls(chunk_elems=chunk_elems, n=n, world=tables.shape[0], rank=rank)
        assert tables.shape == (sp.world, sp.world, sp.n_chunks)
        sp.tables, sp.epoch, sp.epoch_add = tables, epoch, int(epoch_add)
        sp.count = torch.zeros(sp.n_chunks, dtype=torch.int64)
        return sp

    # -- what the writers call -----------------------------------------------------------------------------------
    def gemm_arg(self, elem_offset: int) -> List[int]:
        """``produced`` argument of the extension's ``gemm_tcgen05`` for a master matrix that starts at ``elem_offset``."""
        assert self.sig is not None
        return [self.sig.data_ptr(), int(elem_offset), self.max_ctas]

    def mark(self, lo: int, hi: int) -> None:
        """Arena elements ``[lo, hi)`` are final (work queued so far on the current stream has written them)."""
        lo, hi = int(lo), min(int(hi), self.n)
        if hi <= lo:
            return
        if self.sig is not None:
            self.mod.produced_mark(self.sig.data_ptr(), self.chunk_elems, lo, hi)
            return
        ce = self.chunk_elems
        for c in range(lo // ce, (hi - 1) // ce + 1):
            seg = min(hi, (c + 1) * ce) - max(lo, c * ce)
            full = min(ce, self.n - c * ce)
            self.count[c] += seg
            assert self.count[c] <= full, f"chunk {c}: {int(self.count[c])} of {full} elements reported (something reported twice)"
            if self.count[c] == full:
                self.count[c] = 0
                self.tables[c % self.world, self.rank, c] = int(self.epoch) + self.epoch_add

    def idle(self) -> bool:
        """No chunk is partially reported (true between rounds: the completing report resets its counter)."""
        return bool((self.count[: self.n_chunks] == 0).all())
