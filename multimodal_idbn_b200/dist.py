"""Data parallelism over one 8 x B200 box: one process per GPU (torchrun), ``torch.distributed``.

Training (``RBM.train_epoch`` / ``train_epoch_clamped``): every rank holds the full parameters and
momenta, receives its own shard of the minibatch, computes the local CD statistics
``[dS | dh | dv | sum pos_h | squared error]`` with the same kernels, and ONE sum all-reduce (NCCL
over NVLink / NVSwitch) per update makes them global; every rank then applies the identical update,
so replicas stay bit-identical.  Random numbers are addressed by the GLOBAL row index, so the result
does not depend on the number of ranks (up to the summation order of the all-reduce).

Inference (``conditional_gibbs``, ``noisy_meanfield_annealed``, ``_cross_reconstruct``): chains are
independent -- shard the rows, no collective.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as td


class _State:
    def __init__(self, group, rank: int, world: int):
        self.group, self.rank, self.world = group, rank, world

    def all_reduce(self, t: torch.Tensor) -> None:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)


_state: Optional[_State] = None


def enable(group=None) -> None:
    """Turn on statistic all-reduce for every RBM update in this process.  Every rank must feed an
    equally sized shard of each minibatch, rank r holding global rows [r*B, (r+1)*B)."""
    global _state
    if not td.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state = _State(group, td.get_rank(group), td.get_world_size(group))


def disable() -> None:
    global _state
    _state = None


def state() -> Optional[_State]:
    return _state


def init_from_env(backend: Optional[str] = None) -> int:
    """torchrun helper: pick the device from LOCAL_RANK, join the process group, return the rank."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not td.is_initialized():
        td.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                              rank=rank, world_size=world)
    return rank


def shard_rows(n: int, rank: int, world: int):
    """Contiguous, equally sized row shards (the remainder is dropped so every rank agrees)."""
    per = n // world
    return rank * per, (rank + 1) * per
