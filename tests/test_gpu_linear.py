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
