"""CPU: oracle of the single-matrix ("linear / pseudo-affine") NW, oracle/nw_linear.c, against
the hand-derived matrix of SURVEY.md 8a(C), the frozen vectors, and the literal Python model.

Reference: /root/reference/src/needleman_wunsch.rs:36-117,180-272 (dead code at the reference
commit, no tests of its own -> parity unpinned by reference vectors).
"""
import json
import os
import random

import numpy as np

from oracle import literal_model as L
from tests.util import mutate, random_seq

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "linear_golden.json")


def test_hand_derived_matrix(oracle):
    s, mv, g = oracle.linear_matrices(b"ACGT", b"AGT")
    exp = [[-16, -14, -20, -26], [-14, -11, -18, -24], [-20, -18, -15, -22], [-26, -24, -13, -19], [-32, -30, -21, -8]]
    assert s.tolist() == exp  # scores[0][0] is initialised twice (:45-64) -> -16
    r = oracle.linear_align(b"ACGT", b"AGT")
    assert (r.score, r.n_hits) == (-8, 1)
    # single hit ACGT / -AGT: the -16 origin makes a leading gap beat A-GT
    assert r.cigar == [(1 << 2) | oracle.OP_I, (3 << 2) | oracle.OP_M]


def test_golden_vectors(oracle):
    vec = json.load(open(GOLDEN))["vectors"]
    assert len(vec) >= 150
    for v in vec:
        q, d = v["seq1"].encode(), v["seq2"].encode()
        r = oracle.linear_align(q, d)
        assert (r.score, r.n_hits) == (v["score"], v["n_hits"]), v
        s, _, _ = oracle.linear_matrices(q, d)
        assert s[-1].tolist() == v["last_row"]
        if v["first_hit"]:
            row1, row2, st1, st2 = v["first_hit"]
            assert r.cigar == L.columns_to_cigar(row1.encode(), row2.encode())
            assert (r.start1, r.start2) == (st1, st2)


def test_cross_check_with_literal_model_global_and_local(oracle):
    rng = random.Random(21)
    for _ in range(300):
        q = random_seq(rng, rng.randint(0, 30), b"ACGTN")
        d = mutate(rng, q, rng.choice([0.1, 0.3])) if rng.random() < 0.7 else random_seq(rng, rng.randint(0, 30))
        for local in (False, True):
            o = L.linear_align(q, d, local=local, max_hits=20000)
            if o.truncated:
                continue
            r = oracle.linear_align(q, d, local=local)
            s, _, _ = oracle.linear_matrices(q, d, local=local)
            assert s.tolist() == o.scores
            assert r.n_hits == len(o.hits), (q, d, local)
            if not local:
                assert r.score == o.score
            if o.hits:
                row1, row2, st1, st2 = o.hits[0]
                assert r.cigar == L.columns_to_cigar(row1.encode(), row2.encode()), (q, d, local)
                assert (r.start1, r.start2) == (st1, st2)


def _hit_text(h):
    """println!("\\nHit: {}\\n", hit) with Display for Hit (needleman_wunsch.rs:155-178, :207)."""
    q, d, s1, s2 = h
    bars = "".join("|" if a == b else " " for a, b in zip(q, d))
    return f"\nHit: \nseq1: {q}\n      {bars}\nseq2: {d}\nstart in seq1: {s1}\nstart in seq2: {s2}\n\n\n\n"


def test_hit_text_and_end_cells_against_the_literal_model(oracle):
    """The oracle's printer (every hit, in the reference's order) and the start cell of the first
    hit against the object-literal Python model, both modes."""
    rng = random.Random(33)
    for it in range(400):
        alpha = b"ACGT" if it % 3 else b"AC"
        q = random_seq(rng, rng.randint(0, 14), alpha)
        d = mutate(rng, q, 0.25, True, alpha) if rng.random() < 0.5 else random_seq(rng, rng.randint(0, 14), alpha)
        for local in (False, True):
            o = L.linear_align(q, d, local=local, max_hits=200)
            text, n = oracle.linear_print_hits(q, d, local, 200)
            assert n == len(o.hits) and text == "".join(_hit_text(h) for h in o.hits), (q, d, local)
            r = oracle.linear_align(q, d, local=local)
            if not local:
                assert (r.end1, r.end2) == (len(q), len(d))
            else:  # the first maximum in row-major order (:256-272)
                best = max(max(row) for row in o.scores)
                cells = [(i, j) for i, row in enumerate(o.scores) for j, v in enumerate(row) if v == best]
                assert (r.end1, r.end2) == cells[0] and r.score == best


def test_local_batch_matches_single_calls(oracle):
    from sequencealigning_b200 import PairBatch
    from tests.util import random_pair_list
    pairs = random_pair_list(9, 60, 0, 40)
    b = PairBatch.from_pairs(pairs)
    ref = oracle.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=90, n_threads=3, local=True)
    for p, (q, d) in enumerate(pairs):
        r = oracle.linear_align(q, d, local=True)
        assert (ref.score[p], ref.end1[p], ref.end2[p], ref.cigar(p)) == (r.score, r.end1, r.end2, r.cigar)
