"""Data parallelism over one 8 x B200 box: one process per GPU (torchrun), ``torch.distributed``.

Training (``RBM.train_epoch`` / ``train_epoch_clamped``): every rank holds the full parameters and
momenta, receives its own shard of the minibatch, computes the local CD statistics
``[dS | dh | dv | sum pos_h | squared error]`` with the same kernels, and ONE sum all-reduce (NCCL
over NVLink / NVSwitch) per update makes them global; every rank then applies the identical update,
so replicas stay bit-identical.  Identical STARTING parameters are established by the first update after
``enable()``: it broadcasts rank 0's W, biases and momenta (``RBM._dp_sync_params``).  Random numbers are addressed by the GLOBAL row index, so the result
does not depend on the number of ranks (up to the summation order of the all-reduce).

With NCCL ranks on one NVLink / NVSwitch box the exchange does not go through NCCL at all (``p2p``
mode, the default when symmetric memory is available): the statistics buffers and the weight matrices live
in peer-mapped symmetric memory and ``imdbn_dp_update`` does reduce-scatter + slab update + all-gather in
one kernel over NVLink loads / stores (csrc/dp_update.cuh).  The momentum matrix ``W_m`` is then only
maintained on the rank that owns the slab; ``sync_momenta`` (called by ``save_model``) makes it whole.

Inference (``conditional_gibbs``, ``noisy_meanfield_annealed``, ``_cross_reconstruct``): chains are
independent -- shard the rows, no collective.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as td


class _State:
    def __init__(self, group, rank: int, world: int, p2p: bool = False):
        self.group, self.rank, self.world, self.p2p = group, rank, world, p2p
        # In-switch reduction / broadcast through NVLS multicast addresses.  Per GPU and direction the links
        # carry 2(N-1)/N * S bytes with unicast loads / stores and (N+1)/N * S with multimem (a rank's own
        # slab also travels to the switch), so multimem pays from 4 ranks up; measured at N = 2: 175 us against
        # 109 us for the 60 MB layer.  IMDBN_DP_MULTICAST=0/1 overrides.
        mc = os.environ.get("IMDBN_DP_MULTICAST")
        self.multicast = p2p and (world >= 4 if mc is None else mc != "0")

    def all_reduce(self, t: torch.Tensor) -> None:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)

    # ---- peer-mapped (symmetric) memory: torch allocates and exchanges the CUDA IPC handles --------------
    def symm_empty(self, numel: int):
        """fp32 buffer of ``numel`` elements mapped into every rank; returns ``(tensor, handle)`` with
        ``handle.buffer_ptrs[r]`` = rank r's buffer in this process' address space.  Collective."""
        import torch.distributed._symmetric_memory as sm
        t = sm.empty(int(numel), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        h = sm.rendezvous(t, self.group if self.group is not None else td.group.WORLD)
        return t, h

    def slab(self, n_quads: int):
        """The float4 range of a flattened weight matrix that this rank owns (imdbn_dp_update)."""
        return n_quads * self.rank // self.world, n_quads * (self.rank + 1) // self.world


_state: Optional[_State] = None


def _p2p_available(group, world: int) -> bool:
    if os.environ.get("IMDBN_DP_P2P", "1") == "0" or world < 2 or world > 8 or not torch.cuda.is_available():
        return False
    if td.get_backend(group) != "nccl":
        return False
    try:                                     # one tiny collective allocation proves the whole path works
        import torch.distributed._symmetric_memory as sm
        t = sm.empty(64, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        h = sm.rendezvous(t, group if group is not None else td.group.WORLD)
        ok = len(h.buffer_ptrs) == world
        h.barrier(channel=0)
    except Exception:                        # noqa: BLE001 - any failure means "use the NCCL all-reduce"
        ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    td.all_reduce(flag, op=td.ReduceOp.MIN, group=group)          # all ranks must agree
    return bool(flag.item())


def enable(group=None, p2p: Optional[bool] = None) -> None:
    """Turn on the cross-rank statistics exchange for every RBM update in this process.  Every rank must
    feed an equally sized shard of each minibatch, rank r holding global rows [r*B, (r+1)*B).
    ``p2p`` = exchange over peer memory instead of an NCCL all-reduce (default: when available)."""
    global _state
    if not td.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world = td.get_world_size(group)
    use_p2p = _p2p_available(group, world) if p2p is None or p2p else False
    if p2p and not use_p2p:
        raise RuntimeError("peer-memory data parallelism is not available in this process group")
    _state = _State(group, td.get_rank(group), world, use_p2p)


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
