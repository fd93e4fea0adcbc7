"""N>1 host logic on CPU: 2 gloo ranks shard the blocks (dwsolver_b200.problem.shard_range), run rounds with duals
down / proposals up and must reproduce the single-process result exactly.  The per-rank subproblem side is the CPU
oracle here (no GPU in this suite) and the exchange is gloo; on GPUs the same round is dwgpu_round_comm (include/dwgpu.h):
ncclBroadcast of the duals, proposals stored into rank 0's up-link buffer at the rank's block offset."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dwsolver_b200 import problem as P


class S:
    """the round of SURVEY.md §8e restated over torch.distributed (test infrastructure)"""
    RECORD_FIELDS = ("obj", "cost", "status", "len", "ind", "val")
    shard_range = staticmethod(P.shard_range)

    @staticmethod
    def duals_down(dist_, pi_tensor, src=0):
        dist_.broadcast(pi_tensor, src=src)
        return pi_tensor

    @staticmethod
    def proposals_up(dist_, local, gathered):
        """every rank's records land at its block offset of the K_total-wide arrays (rank-major = block order)"""
        for f in S.RECORD_FIELDS:
            dist_.all_gather_into_tensor(gathered[f].view(-1), local[f].reshape(-1))
        return gathered

    @staticmethod
    def alloc_gathered(torch_, K_total, L_, device):
        return {"obj": torch_.empty(K_total, dtype=torch_.float64, device=device),
                "cost": torch_.empty(K_total, dtype=torch_.float64, device=device),
                "status": torch_.empty(K_total, dtype=torch_.int32, device=device),
                "len": torch_.empty(K_total, dtype=torch_.int32, device=device),
                "ind": torch_.empty((K_total, L_), dtype=torch_.int32, device=device),
                "val": torch_.empty((K_total, L_), dtype=torch_.float64, device=device)}


K, NODES, COLS, L, SEED = 8, 8, 20, 4, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(props, L):
    k = len(props)
    rec = {"obj": torch.tensor([p["obj"] for p in props], dtype=torch.float64),
           "cost": torch.tensor([p["cost"] for p in props], dtype=torch.float64),
           "status": torch.tensor([p["status"] for p in props], dtype=torch.int32),
           "len": torch.tensor([p["len"] for p in props], dtype=torch.int32),
           "ind": torch.zeros((k, L), dtype=torch.int32), "val": torch.zeros((k, L), dtype=torch.float64)}
    for i, p in enumerate(props):
        rec["ind"][i, :p["len"]] = torch.from_numpy(p["ind"].astype(np.int32))
        rec["val"][i, :p["len"]] = torch.from_numpy(p["val"])
    return rec


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng_blocks = S.shard_range(K, world, rank)
    prob = P.gen_mcf(K, NODES, COLS, L, SEED, block_range=rng_blocks)
    blocks = O.OracleBlocks(prob)
    gathered = S.alloc_gathered(torch, K, L, "cpu")
    S.proposals_up(dist, _records(blocks.initial(), L), gathered)
    pis = np.random.default_rng(1).random((3, L)) * -4.0
    for r in range(3):
        pi = torch.from_numpy(pis[r].copy()) if rank == 0 else torch.zeros(L, dtype=torch.float64)
        S.duals_down(dist, pi, src=0)
        S.proposals_up(dist, _records(blocks.iterate(pi.numpy(), r == 0), L), gathered)
    if rank == 0:
        torch.save({k_: v.clone() for k_, v in gathered.items()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    assert list(S.shard_range(8, 2, 1)) == [4, 5, 6, 7]
    with pytest.raises(ValueError):
        S.shard_range(7, 2, 0)


def test_two_rank_rounds_match_single_process(tmp_path):
    from oracle import oracle as O
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    prob = P.gen_mcf(K, NODES, COLS, L, SEED)
    blocks = O.OracleBlocks(prob)
    blocks.initial()
    pis = np.random.default_rng(1).random((3, L)) * -4.0
    for r in range(3):
        props = blocks.iterate(pis[r], r == 0)
    ref = _records(props, L)
    for f in S.RECORD_FIELDS:
        assert torch.equal(got[f], ref[f]), f
