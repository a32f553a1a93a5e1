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
    assert lib.sa_abi_version() == 1


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
