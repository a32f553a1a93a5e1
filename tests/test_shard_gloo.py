"""CPU, world_size 2 over gloo: the multi-GPU path is "shard the pair list, align shards
independently, gather on the host" -- no data-path collective.  Here each rank's compute is
stood in for by the oracle (tests may use it); what is under test is the sharding, the
per-rank result layout and the host-side gather into input order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    from sequencealigning_b200 import synth
    from sequencealigning_b200.engine import AlignResult
    from sequencealigning_b200.shard import gather_results, shard_indices

    batch = synth.random_pairs(600, 60, 0.1, True, seed=5)
    # make it ragged so LPT has something to balance
    batch.q_len[::3] = 30
    idx = shard_indices(batch, world)
    mine = batch.select(idx[rank])
    stride = 200
    r = ob.affine_batch(mine.residues, mine.q_off, mine.q_len, mine.d_off, mine.d_len, cigar_stride=stride)
    off = np.zeros(mine.n_pairs, np.uint64)
    off[1:] = np.cumsum(r.cigar_len[:-1], dtype=np.uint64)
    mask = np.arange(stride)[None, :] < r.cigar_len[:, None]
    local = AlignResult(r.score, r.status, off, r.cigar_len, r.cigar_pool[mask])
    gathered = [None] * world
    dist.all_gather_object(gathered, (local.score, local.status, local.cigar_off, local.cigar_len, local.cigar))
    cells = torch.tensor([float(mine.cells)], dtype=torch.float64)
    dist.all_reduce(cells)
    if rank == 0:
        full = gather_results(batch.n_pairs, idx, [AlignResult(*g) for g in gathered])
        ref = ob.affine_batch(batch.residues, batch.q_off, batch.q_len, batch.d_off, batch.d_len, cigar_stride=stride)
        ok = (np.array_equal(full.score, ref.score) and np.array_equal(full.status, ref.status)
              and np.array_equal(full.cigar_len, ref.cigar_len)
              and all(full.cigar_of(p) == ref.cigar(p) for p in range(0, batch.n_pairs, 7))
              and float(cells.item()) == float(batch.cells)
              and sorted(np.concatenate(idx).tolist()) == list(range(batch.n_pairs)))
        loads = [float((batch.q_len[i].astype(np.float64) * batch.d_len[i]).sum()) for i in idx]
        q.put((ok, max(loads) / (sum(loads) / world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    from oracle import binding as ob
    ob.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, imbalance = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
    assert imbalance < 1.01
