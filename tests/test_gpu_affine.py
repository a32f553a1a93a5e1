"""GPU parity tests proper: the CUDA affine-NW path, called through the C ABI, against the
literal oracle (oracle/nw_affine.c) on the same inputs.  Bit-exact: score, status, CIGAR.

Reference path under test: /root/reference/src/needleman_wunsch_affine.rs:169-334.
"""
import numpy as np
import pytest

from tests.util import check_against_oracle, random_pair_list, rescore_cigars

pytestmark = pytest.mark.gpu


def _batch(pairs):
    from sequencealigning_b200 import PairBatch
    return PairBatch.from_pairs(pairs)


def test_known_answers(engine, oracle):
    # SURVEY.md 8c seed vectors (hand-checkable) -- (seq1, seq2, score, status, cigar string)
    from sequencealigning_b200 import OK, REF_PANIC_EARLY
    kats = [
        (b"ACGT", b"ACGT", 20, OK, "4M"),
        (b"ACGT", b"AGT", 1, OK, "1M1I2M"),
        (b"AGT", b"ACGT", 1, OK, "1M1D2M"),
        (b"ACGTT", b"ACGT", 6, OK, "3M1I1M"),
        (b"AAAA", b"AAA", 1, OK, "2M1I1M"),
        (b"ACGTACGT", b"ACGGT", -1, OK, "3M3I2M"),
        (b"GACGT", b"ACGT", 0, REF_PANIC_EARLY, ""),
        (b"ACGT", b"GACGT", 0, REF_PANIC_EARLY, ""),
        (b"", b"", 0, OK, ""),
        (b"ACG", b"", -8 - 6 * 4, REF_PANIC_EARLY, ""),
        (b"", b"AC", -8 - 6 * 3, REF_PANIC_EARLY, ""),
        (b"NNNN", b"NNNN", 20, OK, "4M"),
    ]
    b = _batch([(k[0], k[1]) for k in kats])
    r = engine.align(b)
    for i, (_, _, score, status, cig) in enumerate(kats):
        assert r.score[i] == score, (i, kats[i], r.score[i])
        assert r.status[i] == status, (i, kats[i], r.status[i])
        assert r.cigar_string(i) == cig, (i, kats[i], r.cigar_string(i))
    check_against_oracle(oracle, b, r, what="kats")


@pytest.mark.parametrize("seed,n,lo,hi", [(1, 3000, 0, 40), (2, 2000, 1, 130), (3, 1500, 100, 300)])
def test_random_ragged(engine, oracle, seed, n, lo, hi):
    b = _batch(random_pair_list(seed, n, lo, hi))
    r = engine.align(b)
    ref = check_against_oracle(oracle, b, r, what=f"ragged seed {seed}")
    # the set must exercise every status the short-pair path can produce
    if seed == 1:
        assert {0, 1, 4} <= set(np.unique(ref.status).tolist())


@pytest.mark.parametrize("length,indels", [(150, False), (150, True), (250, True)])
def test_config_shapes(engine, oracle, length, indels):
    from sequencealigning_b200 import synth
    b = synth.random_pairs(20000, length, 0.05, indels, seed=0x5A02 + length + indels)
    r = engine.align(b)
    ref = check_against_oracle(oracle, b, r, what=f"{length}bp indels={indels}")
    if not indels:  # substitutions only: the reference prints exactly one alignment, never panics
        assert (ref.status == 0).all()
        assert (r.cigar_len == 1).all() and (r.cigar == (150 << 2)).all()


@pytest.mark.parametrize("g", [1, 2, 4, 8, 16, 32])
def test_every_lane_group_width(oracle, g, monkeypatch):
    """Every instantiation of the fill kernel (lanes per pair-of-pairs) gives the same bits."""
    from sequencealigning_b200 import Engine
    monkeypatch.setenv("SA_FORCE_G", str(g))
    with Engine(0) as eng:
        b = _batch(random_pair_list(40 + g, 700, 1, 220))
        r = eng.align(b)
        check_against_oracle(oracle, b, r, what=f"G={g}")


def test_chunked_and_refill_slices(oracle, monkeypatch):
    """A tiny traceback budget forces many chunks and several refill slices per chunk."""
    from sequencealigning_b200 import Engine
    monkeypatch.setenv("SA_TB_BUDGET_MB", "2")
    with Engine(0) as eng:
        b = _batch(random_pair_list(77, 4000, 0, 60, rates=(0.3, 0.5)))
        r = eng.align(b)
        ref = check_against_oracle(oracle, b, r, what="chunked")
        assert (ref.status != 0).sum() > 100


def test_many_small_segments_stream_results(oracle, monkeypatch):
    """Segments of 700 pairs: inputs, per-pair results and the CIGAR pool all stream in pieces."""
    from sequencealigning_b200 import Engine
    monkeypatch.setenv("SA_SEG_PAIRS", "700")
    with Engine(0) as eng:
        b = _batch(random_pair_list(78, 5000, 0, 90))
        r = eng.align(b)
        check_against_oracle(oracle, b, r, what="small segments")
        rb = eng.upload(b)
        rb.align()
        r2 = rb.download()
        rb.free()
        assert np.array_equal(r.cigar, r2.cigar) and np.array_equal(r.score, r2.score)


@pytest.mark.parametrize("sort_mode", ["1", "2"])
def test_shape_bucketing_does_not_change_results(oracle, sort_mode, monkeypatch):
    """SA_SORT=1 forces the per-segment (rows, columns) bucketing, 2 disables it."""
    from sequencealigning_b200 import ALGO_NW_LINEAR, Engine
    monkeypatch.setenv("SA_SORT", sort_mode)
    monkeypatch.setenv("SA_SEG_PAIRS", "1500")
    with Engine(0) as eng:
        b = _batch(random_pair_list(91, 4000, 0, 200))
        r = eng.align(b)
        check_against_oracle(oracle, b, r, what=f"SA_SORT={sort_mode}")
        rl = eng.align(b, algo=ALGO_NW_LINEAR)
        ref = oracle.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=401, n_threads=8)
        assert np.array_equal(ref.score, rl.score) and np.array_equal(ref.cigar_len, rl.cigar_len)


def test_two_bit_packed_input_gives_identical_results(engine, oracle):
    """sa_batch_t.packing = 1 (2-bit codes, offsets in residues) vs the byte format."""
    from sequencealigning_b200 import ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD, synth
    b = synth.random_pairs(6000, 150, 0.08, True, seed=21)
    b.q_len[::7] = 97          # ragged, and unaligned starts inside bytes
    pb = b.packed()
    assert pb.residues.size * 4 >= b.residues.size > pb.residues.size * 3
    for algo in (0, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD):
        r0, r1 = engine.align(b, algo=algo), engine.align(pb, algo=algo)
        for a, c in ((r0.score, r1.score), (r0.status, r1.status), (r0.cigar_len, r1.cigar_len), (r0.cigar, r1.cigar)):
            assert np.array_equal(a, c), algo
    check_against_oracle(oracle, b, engine.align(pb), what="packed")
    rb = engine.upload(pb)
    rb.align()
    r2 = rb.download()
    rb.free()
    assert np.array_equal(r2.cigar, engine.align(b).cigar)


def test_maximum_packed_size_and_range_extremes(engine, oracle):
    """Pairs at the edge of the 16-bit packed range (n1pad + n2 ~ 3.6 k), including the most
    negative scores the range bound has to cover (nothing matches) and the most positive."""
    import random
    from sequencealigning_b200 import ALGO_NW_LINEAR
    rng = random.Random(9)
    q = bytes(rng.choice(b"ACGT") for _ in range(1740))
    d = bytearray(q)
    for _ in range(80):
        d[rng.randrange(len(d))] = rng.choice(b"ACGT")
    del d[500:520]
    pairs = [(q, bytes(d)), (b"A" * 1700, b"C" * 1760), (b"A" * 1760, b"A" * 1760), (b"ACGT" * 430, b"TGCA" * 440),
             (b"A" * 1750, b"C" * 3), (b"G" * 5, b"T" * 1700)]
    b = _batch(pairs)
    r = engine.align(b)
    check_against_oracle(oracle, b, r, what="max size")
    rl = engine.align(b, algo=ALGO_NW_LINEAR)
    ref = oracle.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=3600, n_threads=6)
    assert np.array_equal(ref.score, rl.score) and np.array_equal(ref.cigar_len, rl.cigar_len)
    # beyond the packed range the affine path switches to the general 32-bit kernel ...
    big = _batch([(b"A" * 2000, b"C" * 2000)])
    check_against_oracle(oracle, big, engine.align(big), what="beyond packed range")
    # ... and so does the linear aligner (literal 32-bit kernel, nw_general.cuh)
    rb = engine.align(big, algo=ALGO_NW_LINEAR)
    refb = oracle.linear_batch(big.residues, big.q_off, big.q_len, big.d_off, big.d_len, cigar_stride=4001, n_threads=1)
    assert np.array_equal(refb.score, rb.score) and np.array_equal(refb.cigar_len, rb.cigar_len)
    assert list(refb.cigar_pool[0, :refb.cigar_len[0]]) == rb.cigar_of(0)


def test_score_only_and_capacity(engine, oracle):
    from sequencealigning_b200 import EngineError
    b = _batch(random_pair_list(5, 500, 20, 80))
    r = engine.align(b, cigar=False)
    ref = oracle.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len)
    assert np.array_equal(r.score, ref.score) and np.array_equal(r.status, ref.status)
    with pytest.raises(EngineError) as ei:
        engine.align(b, cigar_capacity=3)
    assert ei.value.code == -4


def test_non_global_modes_not_implemented(engine):
    from sequencealigning_b200 import MODE_LOCAL, MODE_SEMIGLOBAL, NOT_IMPLEMENTED
    b = _batch([(b"ACGT", b"ACGT"), (b"AC", b"A")])
    for mode in (MODE_LOCAL, MODE_SEMIGLOBAL):  # nw_affine:433-434
        r = engine.align(b, mode=mode)
        assert (r.status == NOT_IMPLEMENTED).all() and (r.cigar_len == 0).all()


def test_one_query_many_db_aliasing(engine, oracle):
    """main.rs:61-62 cross product: every pair shares the same query bytes."""
    from sequencealigning_b200 import PairBatch, Record, synth
    b0 = synth.random_pairs(300, 150, 0.05, True, seed=9)
    query = [Record(b0.query(0), b">q")]
    db = [Record(b0.db(i), b">d%d" % i) for i in range(300)]
    b = PairBatch.from_records(query, db)
    assert (b.q_off == b.q_off[0]).all()
    r = engine.align(b)
    check_against_oracle(oracle, b, r, what="1xN")


def test_resident_path_matches_host_path(engine):
    from sequencealigning_b200 import synth
    b = synth.random_pairs(5000, 150, 0.05, True, seed=11)
    r1 = engine.align(b)
    rb = engine.upload(b)
    rb.align()
    rb.align()  # idempotent
    r2 = rb.download()
    rb.free()
    for a, c in ((r1.score, r2.score), (r1.status, r2.status), (r1.cigar_len, r2.cigar_len), (r1.cigar, r2.cigar)):
        assert np.array_equal(a, c)


def test_render_matches_reference_text(engine, oracle):
    from sequencealigning_b200 import render_affine
    for q, d in [(b"ACGT", b"AGT"), (b"ACGTACGT", b"ACGGT"), (b"AAAA", b"AAA")]:
        b = _batch([(q, d)])
        r = engine.align(b)
        text, n, pan = oracle.affine_print_all(q, d, max_alignments=1)
        assert render_affine(q, d, r.cigar_of(0)) == text


def test_all_cooptimal_alignments_match_reference_order(engine, oracle):
    """every co-optimal alignment, in the reference's LIFO-DFS print order, and the
    point where the reference panics (device-computed parent sets, host traversal)."""
    import json, os, random
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "affine_golden.json")))["vectors"]
    late = 0
    for v in vec:
        q, d = v["seq1"].encode(), v["seq2"].encode()
        text, n, pan = engine.all_alignments(q, d)
        exp_text, exp_n, exp_pan = oracle.affine_print_all(q, d)
        assert (text, n, pan) == (exp_text, exp_n, exp_pan), v
        assert n == v["n_printed"] and pan == v["panicked"]
        late += pan and n > 0
    assert late >= 4
    rng = random.Random(12)
    for _ in range(60):
        q = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(20, 90)))
        d = bytes((c if rng.random() > 0.1 else rng.choice(b"ACGT")) for c in q)
        cut = rng.randrange(len(d))
        d = d[:cut] + d[cut + rng.randint(0, 2):]
        assert engine.all_alignments(q, d, max_alignments=500) == oracle.affine_print_all(q, d, max_alignments=500)
    # custom scheme
    assert engine.all_alignments(b"ACGTT", b"ACGT", scheme=(2, -3, -5, -2)) == oracle.affine_print_all(b"ACGTT", b"ACGT", (2, -3, -5, -2))
    # the first printed alignment is the batched path's CIGAR
    from sequencealigning_b200 import PairBatch, render_affine
    q, d = b"AAAA", b"AAA"
    r = engine.align(PairBatch.from_pairs([(q, d)]))
    assert engine.all_alignments(q, d)[0].startswith(render_affine(q, d, r.cigar_of(0)))


def test_long_pairs_take_the_general_kernel(engine, oracle):
    """Pairs outside the packed 16-bit range (n1 + n2 above ~3.6 k) go through the literal
    32-bit kernel (nw_general.cuh) inside the same batched call, mixed with short pairs."""
    import random
    rng = random.Random(31)

    def related(n, err):
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        d = bytearray()
        for c in q:
            u = rng.random()
            if u < err * 0.5:
                d.append(rng.choice(b"ACGT"))
            elif u < err * 0.75:
                d.append(c); d.append(rng.choice(b"ACGT"))
            elif u < err:
                pass
            else:
                d.append(c)
        return q, bytes(d)

    pairs = random_pair_list(32, 300, 1, 200)
    pairs[7] = related(2100, 0.05)
    pairs[150] = related(2500, 0.02)
    pairs[151] = (b"A" * 2000, b"C" * 2200)           # everything mismatches: deeply negative scores
    pairs[299] = related(1900, 0.1)
    b = _batch(pairs)
    r = engine.align(b)
    check_against_oracle(oracle, b, r, what="mixed long/short")
    assert r.cigar_len[7] > 0 and r.cigar_len[150] > 0


def test_long_pairs_in_waves_and_omitted_alignments(oracle, monkeypatch):
    """A small scratch budget: long pairs are launched in several waves that reuse the traceback
    words; a pair that does not fit its share keeps exact score and status, flagged
    SA_ALIGNMENT_OMITTED, with an empty CIGAR."""
    import random
    from sequencealigning_b200 import Engine
    from tests.util import mutate, random_seq
    monkeypatch.setenv("SA_LONG_LITERAL", "1")     # the literal kernels (the tiled path's fallback)
    monkeypatch.setenv("SA_TB_BUDGET_MB", "400")   # refill/long share 80 MB: ~9 pairs of 2.1 kbp per wave
    rng = random.Random(41)
    longs = []
    for _ in range(22):
        q = random_seq(rng, rng.randint(2000, 2150), b"ACGT")
        longs.append((q, mutate(rng, q, 0.06, True, b"ACGT")))
    big = random_seq(rng, 4200, b"ACGT")
    longs.append((big, mutate(rng, big, 0.03, True, b"ACGT")))   # 35 MB of words: more than 1/8 of the share
    with Engine(0) as eng:
        mixed = random_pair_list(55, 150, 1, 160) + longs
        rng.shuffle(mixed)
        for pairs in (mixed, longs):   # mixed with short pairs / a segment of long pairs only
            b = _batch(pairs)
            r = eng.align(b)
            stride = int((b.q_len.astype(np.int64) + b.d_len.astype(np.int64)).max()) + 1
            ref = oracle.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=stride, n_threads=8)
            omitted = (r.status & 0x80) != 0
            assert np.array_equal(ref.score, r.score)
            assert np.array_equal(ref.status, r.status & 0x7F)
            assert (r.cigar_len[omitted] == 0).all()
            keep = ~omitted
            assert np.array_equal(ref.cigar_len[keep], r.cigar_len[keep])
            for i in np.nonzero(keep)[0]:
                assert list(ref.cigar_pool[i, :ref.cigar_len[i]]) == r.cigar_of(int(i)), i
            n_long = sum(1 for q, d in pairs if len(q) >= 2000)
            assert omitted.sum() <= 1 and keep.sum() >= len(pairs) - 1
            if pairs is mixed:
                assert omitted.sum() == 1   # the 4.2 kbp pair (its words exceed 1/8 of the 80 MB share)


def test_wide_long_pair_kernel_and_pass_boundaries(engine, oracle):
    """Queries of 8192+ columns take the 512-thread form of the long-pair kernel; query lengths
    around the pass width (threads x 8 columns) and the thread width exercise the edge column and
    the end-cell owner; short db sides keep the oracle cheap."""
    import random
    from tests.util import mutate, random_seq
    rng = random.Random(91)
    pairs = []
    for n1, n2 in [(8192, 37), (8193, 120), (12289, 64), (4096, 300), (4097, 300), (2049, 900), (2048, 2000), (6150, 5),
                   (40, 8200), (300, 9000)]:
        q = random_seq(rng, n1, b"ACGT")
        d = mutate(rng, q[: max(n2, 1) * 2], 0.1, True, b"ACGT")[:n2] if rng.random() < 0.7 else random_seq(rng, n2, b"ACGT")
        pairs.append((q, d))
    b = _batch(pairs)
    r = engine.align(b)
    check_against_oracle(oracle, b, r, n_threads=8, what="wide long pairs")


def test_checkpointed_traceback_of_long_pairs(oracle, monkeypatch):
    """SA_LONG_CKPT=1 sends every long pair through the checkpointed traceback (forward kernel keeps
    the right edge of every column pass; the backward kernel recomputes pass by pass into a block
    and walks through it) -- the path that 100 kbp pairs take.  Same bits as the oracle: one,
    two and several passes, the 512-thread form, the sentinel regime, panics and dead ends."""
    import random
    from sequencealigning_b200 import Engine
    from tests.util import mutate, random_seq
    monkeypatch.setenv("SA_LONG_CKPT", "1")
    monkeypatch.setenv("SA_LONG_LITERAL", "1")
    rng = random.Random(97)
    pairs = random_pair_list(56, 60, 1, 150)
    for n1, n2, err in [(2100, 2050, 0.05), (2500, 2600, 0.02), (4100, 700, 0.1), (6200, 300, 0.08), (8200, 120, 0.1),
                        (12300, 60, 0.2), (1900, 1950, 0.3), (300, 4000, 0.1)]:
        q = random_seq(rng, n1, b"ACGT")
        d = mutate(rng, (q * 2)[: n2 * 2], err, True, b"ACGT")[:n2]
        pairs.insert(rng.randrange(len(pairs)), (q, d))
    pairs.append((b"A" * 2000, b"C" * 2200))               # deeply negative scores
    pairs.append((b"A" * 5600, b"C" * 3))                   # sentinel regime
    pairs.append((b"ACGT" * 700, b"TTGCA" * 600))
    pairs.append((b"G" + random_seq(rng, 2300, b"ACGT"), random_seq(rng, 2300, b"ACGT")))
    b = _batch(pairs)
    with Engine(0) as eng:
        r = eng.align(b)
        assert not (r.status & 0x80).any()
        check_against_oracle(oracle, b, r, n_threads=8, what="checkpointed traceback")
        only_long = _batch([pq for pq in pairs if len(pq[0]) + len(pq[1]) > 3700])
        check_against_oracle(oracle, only_long, eng.align(only_long), n_threads=8, what="checkpointed, long pairs only")


def test_sentinel_regime_long_pairs(engine, oracle):
    """n1 + n2 > ~5.4 k: the reference's finite -32768 'minus infinity' leaks into the matrix
    (SURVEY 7); the general kernel reproduces score, status and first alignment there too."""
    import random
    rng = random.Random(33)
    q = bytes(rng.choice(b"ACGT") for _ in range(3000))
    d = bytearray(q)
    for _ in range(120):
        d[rng.randrange(len(d))] = rng.choice(b"ACGT")
    del d[1000:1010]
    pairs = [(q, bytes(d)), (b"A" * 5600, b"C" * 3), (b"ACGT" * 700, b"TTGCA" * 600)]
    b = _batch(pairs)
    r = engine.align(b)
    check_against_oracle(oracle, b, r, n_threads=3, what="sentinel regime")


def test_argument_errors_are_per_call_codes(engine):
    """Bad inputs come back as SA_E_ARG / SA_E_UNSUPPORTED with a message, never as a crash."""
    import ctypes as C
    from sequencealigning_b200 import EngineError, PairBatch, _capi
    b = PairBatch.from_pairs([(b"ACGT", b"ACGA"), (b"AC", b"A")])
    bad = PairBatch(b.residues, b.q_off, b.q_len, b.d_off + np.uint64(1000), b.d_len)
    with pytest.raises(EngineError) as ei:
        engine.align(bad)
    assert ei.value.code == -2 and "residues_len" in str(ei.value)
    with pytest.raises(EngineError) as ei:
        engine.align(b, scheme=(5, 6, -8, -6))   # mismatch above match
    assert ei.value.code == -5
    with pytest.raises(EngineError) as ei:
        engine.align(b, algo=9)
    assert ei.value.code == -2
    lib = _capi.lib()
    assert lib.sa_align_batch(None, 0, 0, None, None, None) == -2
    empty = PairBatch.from_pairs([])
    r = engine.align(empty)
    assert r.score.size == 0 and r.cigar.size == 0
    # two engines on the same device work side by side
    from sequencealigning_b200 import Engine
    with Engine(0) as e2:
        r1, r2 = engine.align(b), e2.align(b)
        assert np.array_equal(r1.score, r2.score) and np.array_equal(r1.cigar, r2.cigar)


@pytest.mark.parametrize("k,g", [(8, 4), (13, 8), (13, 16), (16, 8), (16, 16), (19, 8), (19, 16)])
def test_every_kernel_form_at_its_column_boundaries(oracle, k, g, monkeypatch):
    """Each compiled (K columns per lane, G lanes) form of the fill kernel, forced, on query
    lengths around its strip and pass boundaries (single-pass forms fall back to the 8-column
    kernel when K*G does not cover the query: that hand-over is part of what is tested)."""
    from sequencealigning_b200 import Engine
    import random
    from tests.util import mutate, random_seq
    monkeypatch.setenv("SA_FORCE_K", str(k))
    monkeypatch.setenv("SA_FORCE_G", str(g))
    rng = random.Random(1000 * k + g)
    lens = sorted({1, 2, k - 1, k, k + 1, 8 * 3 + 1, k * g - k, k * g - 1, k * g, min(k * g + 1, 330), 150, 151})
    with Engine(0) as eng:
        for n1 in lens:
            pairs = []
            for _ in range(48):
                q = random_seq(rng, n1, b"ACGTN")
                d = mutate(rng, q, rng.choice((0.03, 0.1, 0.3)), True, b"ACGTN") if rng.random() < 0.8 else random_seq(rng, rng.randint(1, 200), b"ACGTN")
                pairs.append((q, d))
            b = _batch(pairs)
            r = eng.align(b)
            check_against_oracle(oracle, b, r, what=f"K={k} G={g} n1={n1}")


@pytest.mark.parametrize("scheme", [(1, -1, -1, -1), (2, -3, -5, -2), (5, -4, 0, -6), (1, -30, -2, -1), (10, 0, -20, 0)])
def test_other_scoring_schemes(engine, oracle, scheme):
    """The transform constants (diagonal constant, bias, penalty) are functions of the scheme:
    mismatch costlier than two extensions, zero open, zero extension."""
    b = _batch(random_pair_list(sum(scheme) & 0xFF, 1500, 1, 180))
    r = engine.align(b, scheme=scheme)
    stride = int((b.q_len.astype(np.int64) + b.d_len.astype(np.int64)).max()) + 1
    ref = oracle.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=stride, n_threads=8, scheme=scheme)
    assert np.array_equal(ref.score, r.score)
    assert np.array_equal(ref.status, r.status)
    assert np.array_equal(ref.cigar_len, r.cigar_len)
    mask = np.arange(stride)[None, :] < ref.cigar_len[:, None]
    assert np.array_equal(ref.cigar_pool[mask], r.cigar)


def test_cooptimal_counts_match_the_oracle(engine, oracle):
    """sa_affine_count_cooptimal: per pair, the number of alignments the reference prints when
    nothing panics (oracle: n_cooptimal, the same parent-list path count), incl. empty sides,
    repeats (many co-optimal paths), saturation, the 2-bit input format and another scheme."""
    pairs = random_pair_list(77, 600, 0, 90)
    pairs += [(b"", b""), (b"ACGT", b""), (b"", b"ACGT"), (b"A" * 60, b"A" * 40), (b"AC" * 40, b"CA" * 38),
              (b"ACGT" * 30, b"ACGT" * 30), (b"A" * 200, b"A" * 120)]
    b = _batch(pairs)
    got = engine.count_cooptimal(b)
    exp = np.array([oracle.affine_align(q, d).n_cooptimal for q, d in pairs], np.int64)
    assert np.array_equal(got, exp), np.nonzero(got != exp)[0][:8]
    assert (exp > 1).any() and (exp == 0).any()
    # consistent with the enumeration (sa_affine_all_alignments) on pairs the reference completes
    for i in [k for k in range(len(pairs)) if 0 < exp[k] <= 50][:10]:
        _, n_printed, panicked = engine.all_alignments(*pairs[i])
        if not panicked:
            assert n_printed == exp[i], (i, n_printed, exp[i])
    sch = (2, -3, -5, -2)
    got2 = engine.count_cooptimal(b, scheme=sch)
    exp2 = np.array([oracle.affine_align(q, d, sch).n_cooptimal for q, d in pairs], np.int64)
    assert np.array_equal(got2, exp2)
    # free gap opening: splitting a gap costs nothing, the count explodes and saturates
    sch0 = (5, -4, 0, -6)
    got3 = engine.count_cooptimal(b, scheme=sch0)
    exp3 = np.array([oracle.affine_align(q, d, sch0).n_cooptimal for q, d in pairs], np.int64)
    assert np.array_equal(got3, exp3)
    assert exp3.max() == (2**63 - 1) // 4
    acgt = _batch(random_pair_list(78, 300, 1, 120, alphabet=b"ACGT"))
    assert np.array_equal(engine.count_cooptimal(acgt), engine.count_cooptimal(acgt.packed()))


# ---- the tiled long-pair path (nw_long.cuh) -----------------------------------------------------

def _cigar_score(q: bytes, d: bytes, words, match=5, mismatch=-4, gap_open=-8, gap_ext=-6):
    """Score of an alignment given as (len << 2 | op) runs; checks that it consumes both sequences."""
    qa, da = np.frombuffer(q, np.uint8), np.frombuffer(d, np.uint8)
    y = x = score = 0
    for w in words:
        op, ln = int(w) & 3, int(w) >> 2
        if op == 0:
            eq = int((qa[y:y + ln] == da[x:x + ln]).sum())
            score += eq * match + (ln - eq) * mismatch
            y += ln; x += ln
        elif op == 1:
            score += gap_open + ln * gap_ext
            y += ln
        else:
            score += gap_open + ln * gap_ext
            x += ln
    assert y == len(q) and x == len(d), (y, len(q), x, len(d))
    return score


@pytest.mark.parametrize("s_r", [(0, 0), (1, 256), (2, 512), (4, 1024), (1, 1024)])
def test_tiled_long_pairs_every_tile_shape(oracle, s_r, monkeypatch):
    """Pairs outside the packed range through the tiled 32-bit path, for every tile shape: query and
    db lengths around the strip width (512), the tile heights and the checkpoint spacing (2048)."""
    import random
    from sequencealigning_b200 import Engine
    from tests.util import mutate, random_seq
    if s_r[0]:
        monkeypatch.setenv("SA_LONG_S", str(s_r[0]))
        monkeypatch.setenv("SA_LONG_R", str(s_r[1]))
    rng = random.Random(101 + s_r[0] + s_r[1])
    pairs = random_pair_list(61, 40, 1, 150)
    shapes = [(2049, 2047), (2048, 2048), (4097, 300), (300, 4097), (513, 3600), (3600, 255), (1025, 2900), (4100, 2100),
              (2047, 2049), (1800, 2050), (5000, 600), (2560, 2561)]
    for n1, n2 in shapes:
        q = random_seq(rng, n1, b"ACGT")
        err = rng.choice([0.02, 0.08, 0.25])
        d = mutate(rng, (q * 3)[: n2 * 2], err, True, b"ACGT")[:n2] if rng.random() < 0.8 else random_seq(rng, n2, b"ACGT")
        pairs.insert(rng.randrange(len(pairs)), (q, d))
    pairs.append((b"G" + random_seq(rng, 2300, b"ACGT"), random_seq(rng, 2300, b"ACGT")))
    pairs.append((b"ACGT" * 600, b"ACGT" * 600))            # one run of 2400 M; a leading-gap alternative does not tie
    pairs.append((b"T" + b"ACGT" * 600, b"ACGT" * 600))     # the best alignment starts with a gap: the reference panics
    b = _batch(pairs)
    with Engine(0) as eng:
        r = eng.align(b)
        assert not (r.status & 0x80).any()
        check_against_oracle(oracle, b, r, n_threads=8, what=f"tiled long pairs S,R={s_r}")
        only_long = _batch([pq for pq in pairs if len(pq[0]) + len(pq[1]) > 3700])
        check_against_oracle(oracle, only_long, eng.align(only_long), n_threads=8, what="tiled, long pairs only")
        ro = eng.align(b, cigar=False)
        assert np.array_equal(ro.score, r.score) and np.array_equal(ro.status, r.status)


def test_tiled_long_pairs_hand_dead_ends_to_the_literal_kernel(engine, oracle):
    """A gap of more than ~5.4 k residues costs less through the reference's finite -32768 sentinel
    than through the gap recurrences: the optimal path then STARTS at a parentless sentinel cell,
    the reference prints nothing (or not the greedy path).  The tiled path detects this from the
    provenance bonus of the end cell and hands the pair to the literal kernel."""
    import random
    from sequencealigning_b200 import REF_NO_OUTPUT
    from tests.util import mutate, random_seq
    rng = random.Random(7)
    x = random_seq(rng, 600, b"ACGT")
    pairs = [
        (random_seq(rng, 7000, b"ACGT") + x, mutate(rng, x, 0.03, True, b"ACGT")),   # 7 k leading query residues
        (mutate(rng, x, 0.03, True, b"ACGT"), random_seq(rng, 6500, b"ACGT") + x),   # 6.5 k leading db residues
        (b"A" * 5600, b"C" * 3),
        (random_seq(rng, 3000, b"ACGT"), random_seq(rng, 3100, b"ACGT")),             # unrelated: ordinary long pair
    ]
    q = random_seq(rng, 2600, b"ACGT")
    pairs.append((q, mutate(rng, q, 0.05, True, b"ACGT")))
    b = _batch(pairs)
    r = engine.align(b)
    ref = check_against_oracle(oracle, b, r, n_threads=5, what="dead ends")
    assert (ref.status[:2] == REF_NO_OUTPUT).all()


def test_long_pairs_score_pinned_at_30_and_100_kbp(engine, oracle):
    """BASELINE.json configs[4] sizes, beyond what the full oracle can hold: the score is pinned by
    the oracle's score-only DP (same recurrences and sentinel, O(n1) memory), the alignment by
    re-scoring its CIGAR (it must consume both sequences and add up to the reported score)."""
    from sequencealigning_b200 import synth
    for length, n in ((30_000, 3), (100_000, 2)):
        b = synth.random_pairs(n, length, 0.05, True, seed=0x5A05 + length)
        r = engine.align(b)
        assert not (r.status & 0x80).any()
        for p in range(n):
            if p == 0:   # one oracle pass per size (9e8 / 1e10 cells on one host core)
                assert int(r.score[p]) == oracle.affine_score(b.query(p), b.db(p)), (length, p)
            assert r.status[p] in (0, 1) and r.cigar_len[p] > 0
            assert _cigar_score(b.query(p), b.db(p), r.cigar_of(p)) == int(r.score[p]), (length, p)


def test_config2_full_size_properties(engine, oracle):
    """BASELINE.json configs[1] at its full size (1 M x 150 bp): size-independent properties of every pair --
    each CIGAR consumes exactly n1 query and n2 db residues, offsets are the scan of the lengths, every score is
    the score its own CIGAR re-scores to under the reference's scheme (5 / -4 / -8 / -6 with the boundary gap's
    extra extension, nw_affine.rs:15-20, :195, :207), 2-bit and byte input agree -- plus the oracle on a sample."""
    from sequencealigning_b200 import REF_PANIC, REF_PANIC_EARLY, synth
    b = synth.random_pairs(1_000_000, 150, 0.05, True, seed=synth.SEEDS["config2"])
    r = engine.align(b)
    n = b.n_pairs
    assert set(np.unique(r.status)) <= {0, REF_PANIC, REF_PANIC_EARLY}
    exp_off = np.zeros(n, np.uint64)
    exp_off[1:] = np.cumsum(r.cigar_len[:-1], dtype=np.uint64)
    assert np.array_equal(exp_off, r.cigar_off) and int(r.cigar_len.sum()) == r.cigar.size
    has = r.cigar_len > 0
    assert (has | (r.status == REF_PANIC_EARLY)).all()          # only an early panic prints nothing
    op, ln = r.cigar & 3, (r.cigar >> 2).astype(np.int64)
    starts = r.cigar_off[has].astype(np.int64)
    use1 = np.add.reduceat(np.where(op != 2, ln, 0), starts)      # M and I consume the query
    use2 = np.add.reduceat(np.where(op != 1, ln, 0), starts)      # M and D consume the db sequence
    assert np.array_equal(use1, b.q_len[has].astype(np.int64)) and np.array_equal(use2, b.d_len[has].astype(np.int64))
    score = rescore_cigars(b, r.cigar_off, r.cigar_len, r.cigar)
    bad = np.nonzero(has & (score != r.score))[0]
    assert bad.size == 0, (bad[:5], score[bad[:5]], r.score[bad[:5]])
    rp = engine.align(b.packed())
    assert np.array_equal(r.score, rp.score) and np.array_equal(r.status, rp.status) and np.array_equal(r.cigar, rp.cigar)
    sub = b.select(np.arange(0, n, 997))
    check_against_oracle(oracle, sub, engine.align(sub), what="every 997th pair of config 2")


def test_many_runs_take_the_second_walk(engine, oracle):
    """Pairs with more CIGAR runs than the count walk parks per pair (kTmpRuns = 48, nw_walk.cuh) are walked a second
    time after the scan; mixed with ordinary pairs, around the limit, through the refill as well."""
    import random
    from tests.util import random_seq
    rng = random.Random(48)
    pairs = random_pair_list(480, 400, 50, 200)
    for period in (3, 4, 5, 6, 7, 9):           # every period-th base deleted: ~2 * n / period runs
        for n in (150, 300, 600):
            q = random_seq(rng, n, b"ACGT")
            d = bytes(c for k, c in enumerate(q) if k % period != period - 1)
            pairs.insert(rng.randrange(len(pairs)), (q, d))
            pairs.insert(rng.randrange(len(pairs)), (d, q))
    rng = random.Random(49)
    for _ in range(400):                        # unrelated heads of unequal length: some co-optimal path starts with a gap
        q = random_seq(rng, 200, b"ACGT")
        d = bytes(c for k, c in enumerate(q) if k % 4 != 3)
        pairs.append((random_seq(rng, rng.randrange(1, 12), b"ACGT") + q, random_seq(rng, rng.randrange(1, 12), b"ACGT") + d))
    b = _batch(pairs)
    r = engine.align(b)
    check_against_oracle(oracle, b, r, what="many runs")
    assert (r.cigar_len > 48).sum() >= 10
    assert ((r.cigar_len > 48) & (r.status == 1)).sum() >= 1   # REF_PANIC pairs with long CIGARs (refill + second walk)
