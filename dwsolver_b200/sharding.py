"""Block sharding and the per-round exchange for N > 1 ranks (one process per GPU).

The reference's "communication" is one cond-var broadcast of the duals per round and one semaphore
post per block (src/dw_phases.c:526-548, src/dw_subprob.c:417-427).  With blocks sharded over ranks
the same round becomes: duals down (broadcast of L doubles from the rank that runs the master),
proposals up (all-gather of the fixed-size per-block records).  Works with any torch.distributed
backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

RECORD_FIELDS = ("obj", "cost", "status", "len", "ind", "val")


def shard_range(K: int, world: int, rank: int) -> range:
    """Static contiguous block range of a rank; K must divide evenly."""
    if K % world:
        raise ValueError(f"{K} blocks do not divide over {world} ranks")
    per = K // world
    return range(rank * per, (rank + 1) * per)


def duals_down(dist, pi_tensor, src: int = 0):
    """Broadcast the L master duals from `src` (in place)."""
    dist.broadcast(pi_tensor, src=src)
    return pi_tensor


def proposals_up(dist, local: dict, gathered: dict):
    """All-gather every record field: gathered[f] has the rank-major concatenation of local[f]."""
    for f in RECORD_FIELDS:
        dist.all_gather_into_tensor(gathered[f].view(-1), local[f].reshape(-1))
    return gathered


def proposals_up_record(dist, local_record, gathered_records):
    """ONE all-gather per round: every rank's whole result record (dwgpu_device_record, a uint8 view) into
    the rank-major concatenation that dwgpu_pack_gathered_records() reads."""
    dist.all_gather_into_tensor(gathered_records, local_record)
    return gathered_records


def alloc_gathered(torch, K_total: int, L: int, device):
    return {"obj": torch.empty(K_total, dtype=torch.float64, device=device),
            "cost": torch.empty(K_total, dtype=torch.float64, device=device),
            "status": torch.empty(K_total, dtype=torch.int32, device=device),
            "len": torch.empty(K_total, dtype=torch.int32, device=device),
            "ind": torch.empty((K_total, L), dtype=torch.int32, device=device),
            "val": torch.empty((K_total, L), dtype=torch.float64, device=device)}
