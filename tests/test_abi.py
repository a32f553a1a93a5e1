"""CPU: the C-ABI shared library loads and exports every symbol include/sa_engine.h declares;
the host-only entry points work; the compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sequencealigning_b200 import _capi
    from sequencealigning_b200.build import build_all
    build_all()
    return _capi.lib()


def test_every_declared_symbol_is_exported(lib):
    from sequencealigning_b200 import _capi
    header = open(os.path.join(ROOT, "include", "sa_engine.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sa_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libsa_engine.so does not export {name}"
    assert lib.sa_abi_version() == 3


def test_no_oracle_in_the_product():
    """The product must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "sequencealigning_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "build.py", f"{f} mentions the oracle"


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sequencealigning_b200 import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine(0)
    assert ei.value.code == -1 and "no CPU fallback" in str(ei.value)


def test_partition_lpt(lib):
    from sequencealigning_b200.shard import partition_lpt
    rng = np.random.default_rng(0)
    q = rng.integers(50, 400, 5000).astype(np.uint32)
    d = rng.integers(50, 400, 5000).astype(np.uint32)
    for parts in (1, 2, 4, 8):
        part = partition_lpt(q, d, parts)
        assert part.min() == 0 and part.max() == parts - 1
        load = np.bincount(part, weights=q.astype(np.float64) * d, minlength=parts)
        assert load.max() / load.mean() < 1.01  # length-balanced on n1*n2
        assert np.array_equal(part, partition_lpt(q, d, parts))  # deterministic
    # uniform batches are dealt round-robin
    u = np.full(64, 150, np.uint32)
    assert np.array_equal(partition_lpt(u, u, 4), np.arange(64) % 4)


def test_plan_shards_contiguous_for_many_reads(lib):
    """sa_plan_shards: what a multi-device sa_align_batch does with a pair list.  Many reads ->
    contiguous, cell-balanced ranges of the caller's arrays (no gather, input order kept)."""
    from sequencealigning_b200.shard import plan_shards
    rng = np.random.default_rng(1)
    q = rng.integers(100, 300, 200_000).astype(np.uint32)
    d = (q.astype(np.int64) + rng.integers(-8, 9, q.size)).astype(np.uint32)
    w = q.astype(np.float64) * d + 1
    for parts in (1, 2, 3, 4, 8):
        begin, part, contiguous = plan_shards(q, d, parts)
        assert contiguous
        assert begin[0] == 0 and begin[-1] == q.size and (np.diff(begin.astype(np.int64)) > 0).all()
        assert np.array_equal(part, np.repeat(np.arange(parts), np.diff(begin.astype(np.int64))))
        load = np.array([w[int(begin[k]):int(begin[k + 1])].sum() for k in range(parts)])
        assert load.max() / load.mean() < 1.002
    # sorted by length (length-bucketed batches): the ranges differ in pair count, not in cells
    order = np.argsort(q.astype(np.int64) * d, kind="stable")
    begin, _, contiguous = plan_shards(q[order], d[order], 8)
    assert contiguous
    load = np.array([w[order][int(begin[k]):int(begin[k + 1])].sum() for k in range(8)])
    assert load.max() / load.mean() < 1.002 and np.diff(begin.astype(np.int64)).max() > 1.5 * np.diff(begin.astype(np.int64)).min()


def test_plan_shards_sampled_for_very_large_lists(lib):
    """Shards of >= 512 Ki pairs are planned from sampled block sums (1/16 of the pairs): the ranges stay
    cell-balanced for random lengths, for a length-sorted list and for a periodic one (a db x query cross product)."""
    from sequencealigning_b200.shard import plan_shards
    rng = np.random.default_rng(7)
    n = 2_300_000
    q = rng.integers(100, 300, n).astype(np.uint32)
    d = (q.astype(np.int64) + rng.integers(-8, 9, n)).astype(np.uint32)
    periodic = (100 + (np.arange(n) % 4801) // 24).astype(np.uint32)
    for name, (a, b) in {"random": (q, d), "sorted": (np.sort(q), np.sort(d)), "periodic": (periodic, periodic)}.items():
        w = a.astype(np.float64) * b + 1
        for parts in (2, 4):
            begin, _, contiguous = plan_shards(a, b, parts, want_part=False)
            assert contiguous, name
            assert begin[0] == 0 and begin[-1] == n and (np.diff(begin.astype(np.int64)) > 0).all()
            load = np.array([w[int(begin[k]):int(begin[k + 1])].sum() for k in range(parts)])
            assert load.max() / load.mean() < 1.005, (name, parts, load)


def test_plan_shards_lpt_for_few_uneven_pairs(lib):
    from sequencealigning_b200.shard import partition_lpt, plan_shards
    q = np.array([100_000, 10, 100_000, 20, 50_000, 30, 50_000, 40, 7, 9], np.uint32)
    d = q.copy()
    begin, part, contiguous = plan_shards(q, d, 2)
    assert not contiguous                      # no contiguous cut balances these
    assert np.array_equal(part, partition_lpt(q, d, 2))
    load = np.bincount(part, weights=q.astype(np.float64) * d, minlength=2)
    assert load.max() / load.mean() < 1.0001   # LPT finds {100k, 50k, ..} twice: 1.25e10 cells each
    # degenerate inputs
    begin, part, contiguous = plan_shards(np.zeros(0, np.uint32), np.zeros(0, np.uint32), 4)
    assert contiguous and list(begin) == [0, 0, 0, 0, 0]
    begin, part, contiguous = plan_shards(np.array([5], np.uint32), np.array([5], np.uint32), 4)
    assert sorted(part.tolist()) == [part[0]] and 0 <= part[0] < 4
    import ctypes as C
    assert lib.sa_plan_shards(None, None, 3, 2, None, None, C.byref(C.c_int())) == -2


def test_render_affine_text(lib):
    from sequencealigning_b200 import render_affine
    text = render_affine(b"ACGTACGT", b"ACGGT", [(3 << 2) | 0, (3 << 2) | 1, (2 << 2) | 0])
    assert text == "alignment found\n\nseq1: ACGTACGT\n      |||   ||\nseq2: ACG---GT\n"
    assert render_affine(b"NN", b"NA", [(2 << 2)]) == "alignment found\n\nseq1: NN\n      | \nseq2: NA\n"
    with pytest.raises(ValueError):
        render_affine(b"AC", b"AC", [(5 << 2)])


def test_pair_batch_cross_product_is_db_major():
    # main.rs:61-62: for d in db { for q in query { .. } }
    from sequencealigning_b200 import PairBatch, Record
    b = PairBatch.from_records([Record(b"AA"), Record(b"CCC")], [Record(b"G"), Record(b"TTTT"), Record(b"AC")])
    assert [(b.query(p), b.db(p)) for p in range(b.n_pairs)] == [
        (b"AA", b"G"), (b"CCC", b"G"), (b"AA", b"TTTT"), (b"CCC", b"TTTT"), (b"AA", b"AC"), (b"CCC", b"AC")]


def test_pack_2bit(lib):
    from sequencealigning_b200 import PairBatch
    b = PairBatch.from_pairs([(b"ACGTTGCA", b"GATTACA"), (b"T", b"")])
    pb = b.packed()
    assert pb.packing == 1 and [pb.query(0), pb.db(0), pb.query(1), pb.db(1)] == [b"ACGTTGCA", b"GATTACA", b"T", b""]
    assert pb.residues[0] == 0b11100100 and pb.residues[1] == 0b00011011   # A=0 C=1 G=2 T=3, little-endian pairs of bits
    with pytest.raises(ValueError):
        PairBatch.from_pairs([(b"ACGN", b"A")]).packed()


def test_render_linear_hit_is_the_reference_text(lib, oracle):
    """sa_render_linear_hit (pure host code): from (CIGAR, end cell) of the first hit to the text of
    needleman_wunsch.rs:155-178/:207, against the oracle's literal printer, both modes."""
    import random
    from sequencealigning_b200 import render_linear_hit
    from tests.util import mutate, random_seq
    rng = random.Random(17)
    for it in range(300):
        q = random_seq(rng, rng.randint(0, 25), b"ACGT")
        d = mutate(rng, q, 0.2, True, b"ACGT") if it % 2 else random_seq(rng, rng.randint(0, 25), b"ACGT")
        for local in (False, True):
            r = oracle.linear_align(q, d, local=local)
            exp, n = oracle.linear_print_hits(q, d, local, 1)
            assert n == 1 and render_linear_hit(q, d, r.cigar, r.end1, r.end2) == exp, (q, d, local)
    with pytest.raises(ValueError):
        render_linear_hit(b"AC", b"AC", [(3 << 2) | 0], 2, 2)  # consumes more than the end cell allows


@pytest.mark.gpu
def test_host_register_pins_caller_memory_in_place(lib):
    """sa_host_register / sa_host_unregister: page-lock a buffer the caller already owns (the CLI's parser output)."""
    buf = np.zeros(1 << 20, np.uint8)
    assert lib.sa_host_register(buf.ctypes.data, buf.size) == 0
    assert lib.sa_host_unregister(buf.ctypes.data) == 0
    assert lib.sa_host_register(None, 16) == -2 and lib.sa_host_unregister(None) == -2
