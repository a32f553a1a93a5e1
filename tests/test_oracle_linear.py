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
