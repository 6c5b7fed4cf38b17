"""Intra-op thread scope for the CPU control plane.

The reference's models have a few thousand parameters; a torch op on such a tensor that decides to run as an OpenMP
parallel region pays a fork/join that dwarfs the work — on a busy or over-committed host (a VM with fewer cores than
vCPUs, a Raspberry Pi doing something else) tens of milliseconds per op: measured here, one advanced-indexing op on a
2 x 4 866 tensor took 30 – 68 ms with 8 OpenMP threads and BASELINE config 1 ran at 22 windows/s instead of 157.  The
coordinator and the worker fit handler therefore run their small-model sections with one intra-op thread and restore the
previous setting afterwards.  Large models (conv nets, wide MLPs on a CPU box) keep the pool.
"""
from __future__ import annotations

import contextlib

import torch

SMALL_MODEL_PARAMS = 1 << 20


@contextlib.contextmanager
def small_model_threads(n_params: int, device=None):
    """``with small_model_threads(P, device):`` — one CPU intra-op thread while a small model trains."""
    # `device` is informational: with the model on a GPU the host-side ops of the window (state-dict copies, checkpoint
    # (de)serialisation, index bookkeeping) are just as small, and the setting only concerns CPU intra-op threads
    del device
    if n_params > SMALL_MODEL_PARAMS or torch.get_num_threads() == 1:
        yield
        return
    before = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        yield
    finally:
        torch.set_num_threads(before)
