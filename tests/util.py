"""Shared helpers for the parity tests."""
from __future__ import annotations

import random
from typing import List, Tuple

import numpy as np


def mutate(rng: random.Random, s: bytes, rate: float, indel: bool = True, alphabet: bytes = b"ACGT") -> bytes:
    out = bytearray()
    for c in s:
        if rng.random() < rate:
            k = rng.random()
            if not indel or k < 0.5:
                out.append(rng.choice([b for b in alphabet if b != c] or list(alphabet)))
            elif k < 0.75:
                out.append(c)
                out.append(rng.choice(alphabet))
        else:
            out.append(c)
    return bytes(out)


def random_seq(rng: random.Random, n: int, alphabet: bytes = b"ACGT") -> bytes:
    return bytes(rng.choice(alphabet) for _ in range(n))


def random_pair_list(seed: int, n_pairs: int, min_len: int, max_len: int, rates=(0.05, 0.15, 0.4),
                     alphabet: bytes = b"ACGTN", unrelated: float = 0.2) -> List[Tuple[bytes, bytes]]:
    rng = random.Random(seed)
    pairs = []
    for _ in range(n_pairs):
        n = rng.randint(min_len, max_len)
        q = random_seq(rng, n, alphabet)
        if rng.random() < unrelated:
            d = random_seq(rng, rng.randint(min_len, max_len), alphabet)
        else:
            d = mutate(rng, q, rng.choice(rates), True, alphabet)
        pairs.append((q, d))
    return pairs


def check_against_oracle(oracle, batch, res, n_threads: int = 8, what: str = ""):
    """Bit-exact comparison of (score, status, CIGAR) with the literal oracle."""
    stride = int((batch.q_len.astype(np.int64) + batch.d_len.astype(np.int64)).max()) + 1 if batch.n_pairs else 1
    ref = oracle.affine_batch(batch.residues, batch.q_off, batch.q_len, batch.d_off, batch.d_len,
                              cigar_stride=stride, n_threads=n_threads)
    bad = np.nonzero(ref.score != res.score)[0]
    assert bad.size == 0, f"{what}: {bad.size} score mismatches, first pair {bad[0]}: gpu {res.score[bad[0]]} oracle {ref.score[bad[0]]} q={batch.query(int(bad[0]))!r} d={batch.db(int(bad[0]))!r}"
    bad = np.nonzero(ref.status != res.status)[0]
    assert bad.size == 0, f"{what}: {bad.size} status mismatches, first pair {bad[0]}: gpu {res.status[bad[0]]} oracle {ref.status[bad[0]]} q={batch.query(int(bad[0]))!r} d={batch.db(int(bad[0]))!r}"
    bad = np.nonzero(ref.cigar_len != res.cigar_len)[0]
    assert bad.size == 0, f"{what}: {bad.size} cigar length mismatches, first pair {bad[0]}"
    # offsets are the exclusive scan of the lengths (monotone, in pair order)
    if batch.n_pairs:
        exp_off = np.zeros(batch.n_pairs, np.uint64)
        exp_off[1:] = np.cumsum(res.cigar_len[:-1], dtype=np.uint64)
        assert np.array_equal(exp_off, res.cigar_off), f"{what}: cigar_off is not the scan of cigar_len"
        assert int(res.cigar_len.sum()) == res.cigar.size
    # compare all run words at once
    mask = np.arange(stride)[None, :] < ref.cigar_len[:, None]
    assert np.array_equal(ref.cigar_pool[mask], res.cigar), f"{what}: CIGAR words differ"
    return ref


def rescore_cigars(batch, cigar_off, cigar_len, cigar, scheme=(5, -4, -8, -6)) -> np.ndarray:
    """Score of every pair's CIGAR under the reference's affine scheme (nw_affine.rs:15-20): a diagonal column
    scores match / mismatch, a gap of length L costs open + L * ext, and a gap that STARTS the alignment sits on
    the boundary chain and pays one more extension (:195, :207).  Vectorised: usable on a million pairs.
    Pairs without a CIGAR get 0."""
    match, mismatch, gap_open, gap_ext = scheme
    n = batch.n_pairs
    cigar_len = np.asarray(cigar_len)
    has = cigar_len > 0
    op, ln = cigar & 3, (cigar >> 2).astype(np.int64)
    starts = np.asarray(cigar_off)[has].astype(np.int64)
    runs = cigar_len[has].astype(np.int64)
    pair_of_run = np.repeat(np.nonzero(has)[0], runs)
    first = np.zeros(cigar.size, bool)
    first[starts] = True
    c1 = np.where(op != 2, ln, 0)  # M and I consume the query
    c2 = np.where(op != 1, ln, 0)  # M and D consume the db sequence
    before1 = np.cumsum(c1) - c1
    before2 = np.cumsum(c2) - c2
    pos1 = before1 - np.repeat(before1[starts], runs)  # query residues of the pair before the run
    pos2 = before2 - np.repeat(before2[starts], runs)
    m = op == 0
    col_run = np.repeat(np.nonzero(m)[0], ln[m])
    within = np.arange(col_run.size) - np.repeat(np.cumsum(ln[m]) - ln[m], ln[m])
    i1 = batch.q_off[pair_of_run[col_run]].astype(np.int64) + pos1[col_run] + within
    i2 = batch.d_off[pair_of_run[col_run]].astype(np.int64) + pos2[col_run] + within
    col_score = np.where(batch.residues[i1] == batch.residues[i2], match, mismatch).astype(np.int64)
    score = np.bincount(pair_of_run[col_run], weights=col_score, minlength=n).astype(np.int64)
    gap = ~m
    gap_cost = gap_open + gap_ext * ln + gap_ext * (first & gap)
    score += np.bincount(pair_of_run[gap], weights=gap_cost[gap], minlength=n).astype(np.int64)
    return score
