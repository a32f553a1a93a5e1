"""Deterministic synthetic read pairs for the BASELINE.json configs.

Uniform A/C/G/T queries; db = query mutated per base at `divergence`, either substitutions
only ("2S": the reference prints exactly one alignment and never panics) or with
sub:ins:del = 2:1:1 ("2G").  numpy PCG64 seeded per config; fully vectorised so that a
million pairs take about a second.
"""
from __future__ import annotations

import numpy as np

from .engine import PairBatch

ALPHABET = np.frombuffer(b"ACGT", np.uint8)
SEEDS = {"config1": 0x5A01, "config2": 0x5A02, "config3": 0x5A03, "config4": 0x5A04, "config5": 0x5A05}


def random_pairs(n_pairs: int, length: int, divergence: float = 0.05, indels: bool = True,
                 seed: int = 0x5A02) -> PairBatch:
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 4, size=(n_pairs, length), dtype=np.uint8)
    u = rng.random((n_pairs, length), dtype=np.float32)
    shift = rng.integers(1, 4, size=(n_pairs, length), dtype=np.uint8)
    mutated = u < divergence
    if indels:
        kind = rng.random((n_pairs, length), dtype=np.float32)
        sub = mutated & (kind < 0.5)
        ins = mutated & (kind >= 0.5) & (kind < 0.75)
        dele = mutated & (kind >= 0.75)
    else:
        sub, ins, dele = mutated, np.zeros_like(mutated), np.zeros_like(mutated)
    first = np.where(sub, (q + shift) & 3, q)           # a substitution always changes the base
    count = np.ones((n_pairs, length), np.int64)
    count[ins] = 2
    count[dele] = 0
    d_len = count.sum(axis=1).astype(np.uint32)
    flat_count = count.ravel()
    total = int(flat_count.sum())
    starts = np.cumsum(flat_count) - flat_count           # output position of each query base
    d = np.empty(total, np.uint8)
    keep = flat_count >= 1
    d[starts[keep]] = first.ravel()[keep]
    two = flat_count == 2
    d[starts[two] + 1] = rng.integers(0, 4, size=int(two.sum()), dtype=np.uint8)
    # layout: all queries, then all db sequences
    residues = np.concatenate([ALPHABET[q.ravel()], ALPHABET[d]])
    q_off = (np.arange(n_pairs, dtype=np.uint64) * np.uint64(length))
    q_len = np.full(n_pairs, length, np.uint32)
    d_off = np.uint64(n_pairs * length) + (np.cumsum(d_len, dtype=np.uint64) - d_len)
    return PairBatch(residues, q_off, q_len, d_off.astype(np.uint64), d_len)


def config2(n_pairs: int = 1_000_000, indels: bool = True, seed: int = SEEDS["config2"]) -> PairBatch:
    """Affine NW, 150 bp read pairs at 5 % divergence (BASELINE.json configs[1])."""
    return random_pairs(n_pairs, 150, 0.05, indels, seed)


def config3(n_pairs: int = 10_000_000, seed: int = SEEDS["config3"]) -> PairBatch:
    """Affine NW, 250 bp pairs, 5 % with indels (BASELINE.json configs[2])."""
    return random_pairs(n_pairs, 250, 0.05, True, seed)


def config1(n_db: int = 1000, seed: int = SEEDS["config1"]) -> PairBatch:
    """1 query x n_db db sequences, 150 bp, db = query mutated 5 % (BASELINE.json configs[0])."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 4, size=150, dtype=np.uint8)
    many = random_pairs(n_db, 150, 0.05, True, seed + 1)
    # replace every query by the single shared one and re-derive db from it
    rng2 = np.random.default_rng(seed + 2)
    dbs = []
    for _ in range(n_db):
        out = []
        for b in q:
            x = rng2.random()
            if x < 0.05:
                k = rng2.random()
                if k < 0.5:
                    out.append((b + rng2.integers(1, 4)) & 3)
                elif k < 0.75:
                    out.append(b); out.append(rng2.integers(0, 4))
            else:
                out.append(b)
        dbs.append(np.array(out, np.uint8))
    del many
    d_len = np.array([len(x) for x in dbs], np.uint32)
    residues = np.concatenate([ALPHABET[q]] + [ALPHABET[x] for x in dbs])
    d_off = np.uint64(150) + (np.cumsum(d_len, dtype=np.uint64) - d_len)
    return PairBatch(residues, np.zeros(n_db, np.uint64), np.full(n_db, 150, np.uint32), d_off.astype(np.uint64), d_len)


def config4(n_pairs: int = 100_000, seed: int = SEEDS["config4"], min_len: int = 1000, max_len: int = 10000) -> PairBatch:
    """WFA workload (BASELINE.json configs[3]): query length log-uniform in [1, 10] kbp, per-pair
    error rate uniform in [1 %, 15 %], sub:ins:del = 2:1:1."""
    rng = np.random.default_rng(seed)
    lens = np.round(10 ** rng.uniform(np.log10(min_len), np.log10(max_len), n_pairs)).astype(np.int64)
    err = rng.uniform(0.01, 0.15, n_pairs)
    total = int(lens.sum())
    q = rng.integers(0, 4, size=total, dtype=np.uint8)
    pair_of = np.repeat(np.arange(n_pairs), lens)
    u = rng.random(total, dtype=np.float32)
    mutated = u < err[pair_of].astype(np.float32)
    kind = rng.random(total, dtype=np.float32)
    sub = mutated & (kind < 0.5)
    ins = mutated & (kind >= 0.5) & (kind < 0.75)
    dele = mutated & (kind >= 0.75)
    shift = rng.integers(1, 4, size=total, dtype=np.uint8)
    first = np.where(sub, (q + shift) & 3, q)
    count = np.ones(total, np.int64)
    count[ins] = 2
    count[dele] = 0
    starts = np.cumsum(count) - count
    d = np.empty(int(count.sum()), np.uint8)
    keep = count >= 1
    d[starts[keep]] = first[keep]
    two = count == 2
    d[starts[two] + 1] = rng.integers(0, 4, size=int(two.sum()), dtype=np.uint8)
    q_off = np.cumsum(lens) - lens
    d_len = np.add.reduceat(count, q_off).astype(np.uint32)
    d_off = total + (np.cumsum(d_len, dtype=np.uint64) - d_len)
    residues = np.concatenate([ALPHABET[q], ALPHABET[d]])
    return PairBatch(residues, q_off.astype(np.uint64), lens.astype(np.uint32), d_off.astype(np.uint64), d_len)
