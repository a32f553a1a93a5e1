"""CPU: the C oracle of the affine aligner (oracle/nw_affine.c) against (a) hand-derived
known answers, (b) the frozen vectors of tests/golden/affine_golden.json, (c) the independent
object-graph-literal Python transliteration on fresh random pairs.

Reference: /root/reference/src/needleman_wunsch_affine.rs (its own tests :458-470 are empty, so
parity for this aligner is "unpinned" by reference vectors; these are the pins we have).
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import literal_model as L
from tests.util import mutate, random_seq

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "affine_golden.json")


def _status_of(o, oracle):
    if o.panicked:
        return oracle.REF_PANIC if o.alignments else oracle.REF_PANIC_EARLY
    return oracle.OK if o.alignments else oracle.REF_NO_OUTPUT


def test_boundary_rows_are_the_references(oracle):
    # nw_affine.rs:172-216: M[0][0]=0, D[0][y]=(y+1)*-6-8, I[x][0]=-8+(x+1)*-6, everything else -32768
    m, i, d, par = oracle.affine_matrices(b"ACGTA", b"ACG")
    assert m[0, 0] == 0 and i[0, 0] == -32768 and d[0, 0] == -32768
    assert list(d[0, 1:]) == [-8 - 6 * (y + 1) for y in range(1, 6)]
    assert list(i[1:, 0]) == [-8 - 6 * (x + 1) for x in range(1, 4)]
    assert (m[0, 1:] == -32768).all() and (i[0, 1:] == -32768).all()
    assert (m[1:, 0] == -32768).all() and (d[1:, 0] == -32768).all()
    # interior: D[1][y] extends D[0][y] (M[0][y]+open is the sentinel): one extra extension
    assert d[1, 2] == d[0, 2] - 6 and i[2, 1] == i[2, 0] - 6
    assert m[1, 1] == 5 and m[1, 2] == d[0, 1] - 4


def test_hand_derived_known_answers(oracle):
    kats = [
        (b"ACGT", b"ACGT", 20, [("ACGT", "ACGT")]),
        (b"ACGT", b"AGT", 1, [("ACGT", "A-GT")]),
        (b"AGT", b"ACGT", 1, [("A-GT", "ACGT")]),
        (b"ACGTT", b"ACGT", 6, [("ACGTT", "ACG-T"), ("ACGTT", "ACGT-")]),
        (b"AAAA", b"AAA", 1, [("AAAA", "AA-A"), ("AAAA", "A-AA"), ("AAAA", "AAA-")]),
        (b"ACGTACGT", b"ACGGT", -1, [("ACGTACGT", "ACG---GT")]),
    ]
    for q, d, score, aligns in kats:
        r = oracle.affine_align(q, d)
        assert (r.status, r.score, r.n_cooptimal) == (oracle.OK, score, len(aligns))
        assert r.cigar == L.columns_to_cigar(aligns[0][0].encode(), aligns[0][1].encode())
        text, n, pan = oracle.affine_print_all(q, d)
        assert n == len(aligns) and not pan
        exp = "".join("alignment found\n\nseq1: %s\n      %s\nseq2: %s\n" % (a, "".join("|" if x == y else " " for x, y in zip(a, b)), b) for a, b in aligns)
        assert text == exp
    for q, d in [(b"GACGT", b"ACGT"), (b"ACGT", b"GACGT")]:  # first path starts with a gap: :299 / :303
        r = oracle.affine_align(q, d)
        assert r.status == oracle.REF_PANIC_EARLY and r.cigar == [] and r.any_panic
    r = oracle.affine_align(b"", b"")
    assert (r.status, r.score, r.cigar, r.n_cooptimal) == (oracle.OK, 0, [], 1)


def test_golden_vectors(oracle):
    vec = json.load(open(GOLDEN))["vectors"]
    assert len(vec) >= 150
    for v in vec:
        q, d = v["seq1"].encode(), v["seq2"].encode()
        r = oracle.affine_align(q, d)
        exp_status = (oracle.REF_PANIC if v["n_printed"] else oracle.REF_PANIC_EARLY) if v["panicked"] else (oracle.OK if v["n_printed"] else oracle.REF_NO_OUTPUT)
        assert r.score == v["score"], v
        assert r.status == exp_status, v
        assert r.cigar == v["first_cigar"], v
        assert oracle.affine_score(q, d) == v["score"]
        text, n, pan = oracle.affine_print_all(q, d)
        assert n == v["n_printed"] and pan == v["panicked"]
        if v["n_printed"]:
            assert text.split("alignment found\n")[1] == v["stdout_first"]
            assert r.n_cooptimal == v["n_printed"] or v["panicked"]


@pytest.mark.parametrize("seed", [11, 12])
def test_cross_check_with_literal_model(oracle, seed):
    rng = random.Random(seed)
    seen = set()
    for _ in range(400):
        n = rng.randint(0, 34)
        q = random_seq(rng, n, b"ACGTN")
        d = mutate(rng, q, rng.choice([0.05, 0.15, 0.4])) if rng.random() < 0.8 else random_seq(rng, rng.randint(0, 34))
        o = L.affine_align(q, d, max_pops=150000)
        if o.truncated:
            continue
        r = oracle.affine_align(q, d)
        st = _status_of(o, oracle)
        seen.add(st)
        assert (r.status, r.score) == (st, o.score), (q, d)
        if o.alignments:
            assert r.cigar == L.columns_to_cigar(*o.alignments[0]), (q, d)
        if not o.panicked:
            assert r.n_cooptimal == len(o.alignments)
        text, n_printed, pan = oracle.affine_print_all(q, d)
        assert (text, n_printed, pan) == (o.stdout, len(o.alignments), o.panicked)
    assert {oracle.OK, oracle.REF_PANIC_EARLY} <= seen


def test_custom_scheme_and_batch(oracle):
    rng = random.Random(3)
    pairs = [(random_seq(rng, rng.randint(1, 30)), random_seq(rng, rng.randint(1, 30))) for _ in range(50)]
    scheme = (2, -3, -5, -2)
    for q, d in pairs[:20]:
        o = L.affine_align(q, d, L.AffineScheme(gap_opening=-5, gap_extension=-2, mismatch=-3, match_=2), max_pops=100000)
        if not o.truncated:
            assert oracle.affine_align(q, d, scheme).score == o.score
    from sequencealigning_b200.engine import PairBatch
    b = PairBatch.from_pairs(pairs)
    for threads in (1, 3):
        r = oracle.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=64, n_threads=threads)
        for k, (q, d) in enumerate(pairs):
            one = oracle.affine_align(q, d)
            assert (r.score[k], r.status[k], r.cigar(k)) == (one.score, one.status, one.cigar)


def test_sentinel_leak_at_long_lengths(oracle):
    # n1 + n2 > ~5.4k: the finite -32768 "minus infinity" beats real scores:
    # M[1][y] for large y takes the sentinel M[0][y-1] instead of D[0][y-1].
    q = b"A" * 5600
    d = b"C"
    m, i, dd, par = oracle.affine_matrices(q, d)
    assert m[1, 5600] == -32768 - 4          # sentinel + mismatch, not D[0][5599] - 4
    assert dd[0, 5599] < -32768
    assert oracle.affine_score(q, d) == oracle.affine_align(q, d).score


def test_rescoring_helper_agrees_with_the_oracle(oracle):
    """tests/util.rescore_cigars (used on the GPU at full BASELINE sizes) against the oracle's scores: the first
    printed alignment of every pair re-scores to the reported score."""
    from sequencealigning_b200 import PairBatch
    from tests.util import random_pair_list, rescore_cigars
    b = PairBatch.from_pairs(random_pair_list(5, 1500, 0, 120, alphabet=b"ACGT"))
    stride = int((b.q_len.astype(np.int64) + b.d_len).max()) + 1
    ref = oracle.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=stride, n_threads=4)
    off = np.zeros(b.n_pairs, np.uint64)
    off[1:] = np.cumsum(ref.cigar_len[:-1], dtype=np.uint64)
    mask = np.arange(stride)[None, :] < ref.cigar_len[:, None]
    score = rescore_cigars(b, off, ref.cigar_len, ref.cigar_pool[mask])
    has = ref.cigar_len > 0
    assert has.sum() > 1000 and np.array_equal(score[has], ref.score[has].astype(np.int64))
