"""GPU parity: the single-matrix ("linear") NW kernel through the C ABI against the literal
oracle (oracle/nw_linear.c).  Reference: /root/reference/src/needleman_wunsch.rs:36-117,180-254."""
import json
import os

import numpy as np
import pytest

from tests.util import random_pair_list

pytestmark = pytest.mark.gpu


def _check(oracle, b, r, what):
    stride = int((b.q_len.astype(np.int64) + b.d_len).max()) + 1 if b.n_pairs else 1
    ref = oracle.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=stride, n_threads=8)
    bad = np.nonzero(ref.score != r.score)[0]
    assert bad.size == 0, f"{what}: score mismatch at pair {bad[0]}: gpu {r.score[bad[0]]} oracle {ref.score[bad[0]]} {b.query(int(bad[0]))!r} {b.db(int(bad[0]))!r}"
    assert (r.status == 0).all()
    bad = np.nonzero(ref.cigar_len != r.cigar_len)[0]
    assert bad.size == 0, f"{what}: cigar length mismatch at pair {bad[0]}: {b.query(int(bad[0]))!r} {b.db(int(bad[0]))!r} gpu {r.cigar_string(int(bad[0]))}"
    mask = np.arange(stride)[None, :] < ref.cigar_len[:, None]
    assert np.array_equal(ref.cigar_pool[mask], r.cigar), f"{what}: CIGAR words differ"


def test_known_answer_and_golden(engine, oracle):
    from sequencealigning_b200 import ALGO_NW_LINEAR, PairBatch
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "linear_golden.json")))["vectors"]
    pairs = [(b"ACGT", b"AGT")] + [(v["seq1"].encode(), v["seq2"].encode()) for v in vec]
    b = PairBatch.from_pairs(pairs)
    r = engine.align(b, algo=ALGO_NW_LINEAR)
    assert r.score[0] == -8 and r.cigar_string(0) == "1I3M"  # ACGT / -AGT
    for k, v in enumerate(vec, start=1):
        assert r.score[k] == v["score"], v
    _check(oracle, b, r, "golden")


@pytest.mark.parametrize("seed,n,lo,hi", [(1, 3000, 0, 40), (2, 1500, 1, 200), (3, 600, 150, 320)])
def test_random_ragged(engine, oracle, seed, n, lo, hi):
    from sequencealigning_b200 import ALGO_NW_LINEAR, PairBatch
    b = PairBatch.from_pairs(random_pair_list(100 + seed, n, lo, hi))
    _check(oracle, b, engine.align(b, algo=ALGO_NW_LINEAR), f"ragged {seed}")


def test_config1_one_query_many_db(engine, oracle):
    """BASELINE.json configs[0]: 1 query x 1,000 db sequences of ~150 bp."""
    from sequencealigning_b200 import ALGO_NW_LINEAR, synth
    b = synth.config1(1000)
    _check(oracle, b, engine.align(b, algo=ALGO_NW_LINEAR), "config1")


@pytest.mark.parametrize("g", [1, 2, 8, 32])
def test_lane_group_widths(oracle, g, monkeypatch):
    from sequencealigning_b200 import ALGO_NW_LINEAR, Engine, PairBatch
    monkeypatch.setenv("SA_FORCE_G", str(g))
    with Engine(0) as eng:
        b = PairBatch.from_pairs(random_pair_list(300 + g, 500, 1, 150))
        _check(oracle, b, eng.align(b, algo=ALGO_NW_LINEAR), f"G={g}")


def test_long_pairs_take_the_literal_kernel(engine, oracle):
    """Pairs beyond the packed 16-bit range, mixed with short ones in one call: the long ones run
    through the block-per-pair 32-bit kernel (256 and 1024 threads), results in pair order."""
    import random
    from sequencealigning_b200 import ALGO_NW_LINEAR, PairBatch
    from tests.util import mutate, random_seq
    rng = random.Random(77)
    pairs = random_pair_list(400, 200, 1, 150)
    for n in (2100, 2600, 3000):
        q = random_seq(rng, n, b"ACGT")
        pairs.insert(rng.randrange(len(pairs)), (q, mutate(rng, q, 0.08, True, b"ACGT")))
    pairs.append((random_seq(rng, 2500, b"ACGT"), random_seq(rng, 1900, b"ACGT")))
    pairs.append((b"A" * 4000, b"C" * 2))
    b = PairBatch.from_pairs(pairs)
    _check(oracle, b, engine.align(b, algo=ALGO_NW_LINEAR), "long linear")
    wide = PairBatch.from_pairs([(random_seq(rng, 300, b"ACGT"), random_seq(rng, 9000, b"ACGT")), (b"ACGT", b"AGT")])
    _check(oracle, wide, engine.align(wide, algo=ALGO_NW_LINEAR), "wide linear")


# ---------------------------------------------------------------------------------------------
# LOCAL mode (needleman_wunsch.rs:43, :88-89, :107-111, :256-272): nw_local.cuh
# ---------------------------------------------------------------------------------------------
def _check_local(oracle, b, r, what):
    stride = int((b.q_len.astype(np.int64) + b.d_len).max()) + 1 if b.n_pairs else 1
    ref = oracle.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=stride, n_threads=8, local=True)
    for name, got, exp in (("score", r.score, ref.score), ("end1", r.end1, ref.end1), ("end2", r.end2, ref.end2),
                           ("cigar_len", r.cigar_len, ref.cigar_len)):
        bad = np.nonzero(exp != got)[0]
        assert bad.size == 0, (f"{what}: {name} mismatch at pair {bad[0]} of {bad.size}: gpu {got[bad[0]]} oracle {exp[bad[0]]} "
                               f"{b.query(int(bad[0]))!r} {b.db(int(bad[0]))!r}")
    assert (r.status == 0).all()
    exp_off = np.zeros(b.n_pairs, np.uint64)
    exp_off[1:] = np.cumsum(r.cigar_len[:-1], dtype=np.uint64)
    assert np.array_equal(exp_off, r.cigar_off), f"{what}: cigar_off is not the scan of cigar_len"
    mask = np.arange(stride)[None, :] < ref.cigar_len[:, None]
    assert np.array_equal(ref.cigar_pool[mask], r.cigar), f"{what}: CIGAR words differ"
    return ref


def test_local_known_answers(engine, oracle):
    from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL, PairBatch, render_linear_hit
    pairs = [
        (b"TTTTACGTACGTTTTT", b"GGGGACGTACGGGG"),  # one common core
        (b"AAAA", b"CCCC"),       # no positive cell: the maximum 0 is first held by (0, 0) -> empty hit
        (b"", b"ACGT"), (b"ACGT", b""), (b"", b""),
        (b"A", b"A"),
        (b"ACGTACGT", b"ACGTACGT"),
        (b"AACCGGTT", b"TTGGCCAA"),
        (b"ACGTTTTTTTTACGT", b"ACGTACGT"),          # ties between two maxima: row-major first
    ]
    b = PairBatch.from_pairs(pairs)
    r = engine.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    _check_local(oracle, b, r, "local KAT")
    assert r.score[0] == 35 and r.cigar_string(0) == "7M" and (r.end1[0], r.end2[0]) == (11, 11)
    assert r.score[1] == 0 and r.cigar_len[1] == 0 and (r.end1[1], r.end2[1]) == (0, 0)
    assert r.score[5] == 5 and r.cigar_string(5) == "1M"
    # the text of the first hit, byte for byte (Display for Hit, :155-178)
    for p, (q, d) in enumerate(pairs):
        exp, n = oracle.linear_print_hits(q, d, True, 1)
        assert render_linear_hit(q, d, r.cigar_of(p), r.end1[p], r.end2[p]) == exp, (q, d)


@pytest.mark.parametrize("seed,n,lo,hi,alphabet", [(1, 3000, 0, 40, b"ACGTN"), (2, 1500, 1, 200, b"ACGT"), (3, 400, 150, 330, b"ACGT"),
                                                   (4, 2000, 0, 30, b"AC")])
def test_local_random_ragged(engine, oracle, seed, n, lo, hi, alphabet):
    from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL, PairBatch
    b = PairBatch.from_pairs(random_pair_list(700 + seed, n, lo, hi, alphabet=alphabet, unrelated=0.4))
    r = engine.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    _check_local(oracle, b, r, f"local ragged {seed}")
    # score + end cell only (no CIGAR pool): the same numbers
    r2 = engine.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL, cigar=False)
    assert np.array_equal(r.score, r2.score) and np.array_equal(r.end1, r2.end1) and np.array_equal(r.end2, r2.end2)


def test_local_multi_pass_and_global_scratch(engine, oracle):
    """Pairs wider than one pass of 32 x 16 columns (boundary column between passes), matrices too
    large for shared memory (the warp's global scratch), 2-bit packed input, and the resident path."""
    import random
    from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL, PairBatch
    from tests.util import mutate, random_seq
    rng = random.Random(91)
    pairs = random_pair_list(801, 40, 1, 120, alphabet=b"ACGT")
    for n1, n2 in ((700, 1300), (1500, 600), (513, 512), (40, 2100), (2100, 40)):
        q = random_seq(rng, n1, b"ACGT")
        d = mutate(rng, q, 0.1, True, b"ACGT")[:n2] if n2 <= n1 else random_seq(rng, n2 - n1, b"ACGT") + mutate(rng, q, 0.1, True, b"ACGT")
        pairs.insert(rng.randrange(len(pairs)), (q, d))
    b = PairBatch.from_pairs(pairs)
    r = engine.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    _check_local(oracle, b, r, "local multi-pass")
    rp = engine.align(b.packed(), algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    assert np.array_equal(r.score, rp.score) and np.array_equal(r.cigar, rp.cigar) and np.array_equal(r.end2, rp.end2)
    rb = engine.upload(b)
    rb.align(algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    rr = rb.download()
    rb.free()
    assert np.array_equal(r.score, rr.score) and np.array_equal(r.cigar, rr.cigar) and np.array_equal(r.end1, rr.end1)


def test_global_end_cells_and_hit_text(engine, oracle):
    """Global mode reports (n1, n2) as the end cell; the hit text of the first hit equals the oracle's."""
    from sequencealigning_b200 import ALGO_NW_LINEAR, PairBatch, render_linear_hit
    pairs = random_pair_list(55, 60, 0, 30, alphabet=b"ACGT") + [(b"ACGT", b"AGT"), (b"", b"AC"), (b"AC", b""), (b"", b"")]
    b = PairBatch.from_pairs(pairs)
    r = engine.align(b, algo=ALGO_NW_LINEAR)
    assert np.array_equal(r.end1, b.q_len) and np.array_equal(r.end2, b.d_len)
    for p, (q, d) in enumerate(pairs):
        exp, n = oracle.linear_print_hits(q, d, False, 1)
        assert render_linear_hit(q, d, r.cigar_of(p), r.end1[p], r.end2[p]) == exp, (q, d)


def test_all_hits_text_matches_reference_order(engine, oracle):
    """sa_linear_all_hits: EVERY hit of every start cell, in the reference's order and text
    (needleman_wunsch.rs:106-116, :205-254), global and local, incl. the truncation by max_hits and the
    snprintf-style sizing."""
    pairs = random_pair_list(77, 40, 0, 14, alphabet=b"AC") + random_pair_list(78, 20, 0, 24, alphabet=b"ACGT") + [
        (b"ACGT", b"AGT"), (b"", b"AC"), (b"AC", b""), (b"", b""), (b"AAAA", b"CCCC"), (b"AAAAAA", b"AAA"),
        (b"ACGTTTTTTTTACGT", b"ACGTACGT")]
    seen_many = 0
    for q, d in pairs:
        for local in (False, True):
            exp, n = oracle.linear_print_hits(q, d, local, 500)
            got, m = engine.linear_all_hits(q, d, local=local, max_hits=500)
            assert (got, m) == (exp, n), (q, d, local)
            seen_many += n > 1
            if n > 2:
                exp2, n2 = oracle.linear_print_hits(q, d, local, 2)
                assert engine.linear_all_hits(q, d, local=local, max_hits=2) == (exp2, n2)
    assert seen_many > 10
