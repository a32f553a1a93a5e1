"""Multi-GPU sharding: independent length-balanced partitions, no data-path collective.

The reference's pair loop (main.rs:61-62) has no cross-iteration state, so the pair list is
split across ranks by greedy LPT on n1*n2 (sa_partition_lpt) and every rank aligns its own
shard with its own engine; results are gathered on the host in input order.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import _capi
from .engine import AlignResult, PairBatch


def partition_lpt(q_len: np.ndarray, d_len: np.ndarray, n_parts: int) -> np.ndarray:
    q_len = np.ascontiguousarray(q_len, np.uint32)
    d_len = np.ascontiguousarray(d_len, np.uint32)
    part = np.zeros(len(q_len), np.int32)
    rc = _capi.lib().sa_partition_lpt(q_len.ctypes.data, d_len.ctypes.data, len(q_len), n_parts, part.ctypes.data)
    if rc != 0:
        raise ValueError(f"sa_partition_lpt failed: {rc}")
    return part


def plan_shards(q_len: np.ndarray, d_len: np.ndarray, n_parts: int, want_part: bool = True):
    """sa_plan_shards: (begin[n_parts+1], part[n_pairs] or None, contiguous) -- what a multi-device
    sa_align_batch does with this pair list."""
    q_len = np.ascontiguousarray(q_len, np.uint32)
    d_len = np.ascontiguousarray(d_len, np.uint32)
    begin = np.zeros(n_parts + 1, np.uint64)
    part = np.zeros(len(q_len), np.int32) if want_part else None
    contiguous = C.c_int()
    rc = _capi.lib().sa_plan_shards(q_len.ctypes.data, d_len.ctypes.data, len(q_len), n_parts, begin.ctypes.data,
                                    part.ctypes.data if want_part else None, C.byref(contiguous))
    if rc != 0:
        raise ValueError(f"sa_plan_shards failed: {rc}")
    return begin, part, bool(contiguous.value)


def shard_indices(batch: PairBatch, world_size: int) -> List[np.ndarray]:
    part = partition_lpt(batch.q_len, batch.d_len, world_size)
    return [np.nonzero(part == r)[0] for r in range(world_size)]


def gather_results(n_pairs: int, indices: List[np.ndarray], results: List[AlignResult]) -> AlignResult:
    """Reassemble per-rank results into input order (host side; CIGAR pool re-concatenated)."""
    score = np.zeros(n_pairs, np.int32)
    status = np.zeros(n_pairs, np.uint8)
    ln = np.zeros(n_pairs, np.uint32)
    for idx, r in zip(indices, results):
        score[idx] = r.score
        status[idx] = r.status
        ln[idx] = r.cigar_len
    off = np.zeros(n_pairs, np.uint64)
    if n_pairs:
        off[1:] = np.cumsum(ln[:-1], dtype=np.uint64)
    pool = np.zeros(int(ln.sum()), np.uint32)
    for idx, r in zip(indices, results):
        for k, p in enumerate(idx):
            l = int(r.cigar_len[k])
            if l:
                o = int(r.cigar_off[k])
                pool[int(off[p]): int(off[p]) + l] = r.cigar[o:o + l]
    return AlignResult(score, status, off, ln, pool)
