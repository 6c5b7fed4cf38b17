"""Symmetric memory: one identically laid-out allocation per rank, mapped into every peer.

This is the substrate the fused broadcast / FedAvg kernels run on (SURVEY §5 "Distributed
communication backend", §7 phase 3).  The reference's data plane is PySyft websockets
(``fc.py:167``, ``cf.py:209-211``); here "sending a model" is a store to a peer pointer.

Two providers, tried in order (override with ``COLEARN_SYMM=torch|ipc``):

* ``torch`` — ``torch.distributed._symmetric_memory`` (cuMem + fabric/fd handles).  Also yields
  the NVLS **multicast** alias of the region when the box supports it, which
  ``star_round_kernel`` uses for a one-store broadcast (``multimem.st``).
* ``ipc``  — ``cudaMalloc`` + ``cudaIpcGetMemHandle`` exchanged through the process group's
  object collective and opened with ``cudaIpcOpenMemHandle`` (implemented in
  ``csrc/bindings.cpp``).  No multicast, P2P loads/stores only.

``world_size == 1`` needs no exchange at all.  NCCL/gloo are used only for this bootstrap.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import _ext

log = logging.getLogger(__name__)

_ALIGN = 1024
_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2, torch.uint8: 3}
_DTYPE_SIZE = {torch.float32: 4, torch.bfloat16: 2, torch.int32: 4, torch.uint8: 1}


@dataclass
class _Region:
    offset: int
    numel: int
    dtype: torch.dtype


class SymmetricArena:
    """Named sub-buffers inside one symmetric allocation."""

    def __init__(self, layout: Dict[str, Tuple[int, torch.dtype]], device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, provider: Optional[str] = None) -> None:
        self.device = torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.regions: Dict[str, _Region] = {}
        off = 0
        for name, (numel, dtype) in layout.items():
            self.regions[name] = _Region(off, int(numel), dtype)
            off += (int(numel) * _DTYPE_SIZE[dtype] + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total_bytes = max(off, _ALIGN)
        self.base_ptrs: List[int] = []
        self.mc_base: int = 0
        self.provider = "local"
        self._keep = []
        self._ipc_opened: List[int] = []
        self._owned_ptr = 0
        self._ext = _ext.require()
        provider = provider or os.environ.get("COLEARN_SYMM", "auto")
        if self.world == 1:
            self._alloc_local()
        else:
            errors = []
            done = False
            if provider in ("auto", "torch"):
                try:
                    self._alloc_torch()
                    done = True
                except Exception as e:  # noqa: BLE001 - fall through to IPC
                    errors.append(f"torch symmetric memory: {e!r}")
                    if provider == "torch":
                        raise
            if not done:
                try:
                    self._alloc_ipc()
                except Exception as e:  # noqa: BLE001
                    errors.append(f"cuda ipc: {e!r}")
                    raise RuntimeError("no symmetric-memory provider works on this box: " + "; ".join(errors))
            if errors:
                log.info("symmetric memory fell back to %s (%s)", self.provider, "; ".join(errors))
        self.zero_()

    # -- providers -----------------------------------------------------------------------------
    def _alloc_local(self) -> None:
        buf = torch.zeros(self.total_bytes, dtype=torch.uint8, device=self.device)
        self._keep.append(buf)
        self.base_ptrs = [buf.data_ptr()]
        self.provider = "local"

    def _alloc_torch(self) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        group = self.group or dist.group.WORLD
        buf = symm_mem.empty(self.total_bytes, dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(buf, group.group_name)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory handle layout")
        self._keep += [buf, hdl]
        self.base_ptrs = ptrs
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self.mc_base = mc if os.environ.get("COLEARN_MULTICAST", "1") != "0" else 0
        self.provider = "torch"

    def _alloc_ipc(self) -> None:
        ext = self._ext
        with torch.cuda.device(self.device):
            ptr = ext.ipc_alloc(self.total_bytes)
            handle = ext.ipc_get_handle(ptr)
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            gathered: List[Optional[tuple]] = [None] * self.world
            dist.all_gather_object(gathered, (bytes(handle), int(dev_index)), group=self.group)
            ptrs = []
            for r, (h, peer_dev) in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(ptr)
                else:
                    p = ext.ipc_open_handle(h)
                    self._ipc_opened.append(p)
                    ptrs.append(p)
        self._owned_ptr = ptr
        self.base_ptrs = ptrs
        self.provider = "ipc"

    # -- addressing ---------------------------------------------------------------------------------
    def ptr(self, name: str, rank: Optional[int] = None, elem_offset: int = 0) -> int:
        reg = self.regions[name]
        r = self.rank if rank is None else rank
        return self.base_ptrs[r] + reg.offset + elem_offset * _DTYPE_SIZE[reg.dtype]

    def peer_ptrs(self, name: str, elem_offset: int = 0) -> List[int]:
        return [self.ptr(name, r, elem_offset) for r in range(self.world)]

    def mc_ptr(self, name: str, elem_offset: int = 0) -> int:
        if not self.mc_base:
            return 0
        reg = self.regions[name]
        return self.mc_base + reg.offset + elem_offset * _DTYPE_SIZE[reg.dtype]

    def tensor(self, name: str) -> torch.Tensor:
        """Local view of a region as a torch tensor (no ownership)."""
        reg = self.regions[name]
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        return self._ext.tensor_from_ptr(self.ptr(name), reg.numel, _DTYPE_CODE[reg.dtype], dev_index)

    def zero_(self) -> None:
        for name in self.regions:
            self.tensor(name).zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    @property
    def has_multicast(self) -> bool:
        return bool(self.mc_base)

    def close(self) -> None:
        for p in self._ipc_opened:
            self._ext.ipc_close_handle(p)
        self._ipc_opened.clear()
        if self._owned_ptr:
            self._ext.ipc_free(self._owned_ptr)
            self._owned_ptr = 0
        self._keep.clear()
