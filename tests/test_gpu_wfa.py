"""GPU parity for the WFA path (through the C ABI).

Literal mode (SA_ALGO_WFA): status and printed score equal to the literal oracle of
/root/reference/src/wfa.rs on the golden corpus and on random pairs -- including the pairs on
which the reference panics (wfa.rs:577/:603) or never converges (:189).
Standard mode (SA_ALGO_WFA_STANDARD): the optimal gap-affine cost (x=4, o=2, e=6), equal to a
Gotoh cost DP; this is what config-sized inputs can exercise, because the reference itself
produces no result there.
"""
import json
import os
import random

import numpy as np
import pytest

from tests.util import mutate

pytestmark = pytest.mark.gpu
ST = {"OK": 0, "PANIC": 1, "NO_CONVERGENCE": 2}


def _batch(pairs):
    from sequencealigning_b200 import PairBatch
    return PairBatch.from_pairs(pairs)


def test_literal_golden_corpus(engine):
    from sequencealigning_b200 import ALGO_WFA
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wfa_golden.json")))["vectors"]
    b = _batch([(v["seq1"].encode(), v["seq2"].encode()) for v in vec])
    r = engine.align(b, algo=ALGO_WFA)
    for k, v in enumerate(vec):
        assert r.status[k] == ST[v["status"]], (k, v, r.status[k])
        if v["status"] == "OK":
            assert r.score[k] == v["printed_score"], (k, v, r.score[k])
    assert (r.cigar_len == 0).all()  # the reference's WFA traceback never finds a parent (wfa.rs:645-681)


@pytest.mark.parametrize("seed", [1, 2])
def test_literal_random_vs_oracle(engine, oracle, seed):
    from sequencealigning_b200 import ALGO_WFA
    rng = random.Random(seed)
    pairs = []
    for _ in range(1500):
        n = rng.choice([rng.randint(0, 12), rng.randint(8, 60), rng.randint(100, 260)])
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        d = bytes((c if rng.random() > 0.1 else rng.choice(b"ACGT")) for c in q)
        if rng.random() < 0.5 and len(d) > 2:
            cut = rng.randrange(len(d))
            d = d[:cut] + d[cut + 1:]
        pairs.append((q, d))
    b = _batch(pairs)
    r = engine.align(b, algo=ALGO_WFA)
    score, status = oracle.wfa_literal_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len)
    bad = np.nonzero(status != r.status)[0]
    assert bad.size == 0, (pairs[int(bad[0])], status[bad[0]], r.status[bad[0]])
    assert np.array_equal(score, r.score)
    assert set(np.unique(status).tolist()) == {0, 1, 2}


def test_literal_config_sized_inputs_panic_like_the_reference(engine, oracle):
    from sequencealigning_b200 import ALGO_WFA, synth
    b = synth.random_pairs(2000, 150, 0.05, True, seed=8)
    r = engine.align(b, algo=ALGO_WFA)
    score, status = oracle.wfa_literal_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len)
    assert np.array_equal(status, r.status) and np.array_equal(score, r.score)
    assert (r.status == 1).mean() > 0.9


def _related(rng, n, err):
    q = bytes(rng.choice(b"ACGT") for _ in range(n))
    out = bytearray()
    for c in q:
        u = rng.random()
        if u < err:
            k = rng.random()
            if k < 0.5:
                out.append(rng.choice([x for x in b"ACGT" if x != c]))
            elif k < 0.75:
                out.append(c)
                out.append(rng.choice(b"ACGT"))
        else:
            out.append(c)
    return q, bytes(out)


def test_standard_equals_gotoh_cost_small(engine, oracle):
    from sequencealigning_b200 import ALGO_WFA_STANDARD
    rng = random.Random(3)
    pairs = [(b"", b""), (b"A", b""), (b"", b"ACG"), (b"ACGT", b"ACGT"), (b"ACGT", b"AGT"), (b"ACGT", b"ACTT")]
    for _ in range(1500):
        if rng.random() < 0.3:
            pairs.append((bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 70))), bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 70)))))
        else:
            pairs.append(_related(rng, rng.randint(1, 300), rng.choice([0.01, 0.05, 0.15, 0.3])))
    b = _batch(pairs)
    r = engine.align(b, algo=ALGO_WFA_STANDARD)
    exp = np.array([oracle.wfa_gotoh_cost(q, d) for q, d in pairs], np.int32)
    bad = np.nonzero(exp != r.score)[0]
    assert bad.size == 0, (pairs[int(bad[0])], exp[bad[0]], r.score[bad[0]])
    assert (r.status == 0).all()


def test_standard_config4_shaped_pairs(engine, oracle):
    """BASELINE.json configs[3]: 1-10 kbp pairs at 1-15 % error (a sample; cost DP is O(n^2))."""
    from sequencealigning_b200 import ALGO_WFA_STANDARD
    rng = random.Random(4)
    pairs = []
    for _ in range(48):
        n = int(round(10 ** rng.uniform(3, 4)))
        pairs.append(_related(rng, n, rng.uniform(0.01, 0.15)))
    b = _batch(pairs)
    r = engine.align(b, algo=ALGO_WFA_STANDARD)
    exp = np.array([oracle.wfa_gotoh_cost(q, d) for q, d in pairs], np.int32)
    assert np.array_equal(exp, r.score)
    # custom penalties
    r2 = engine.align(b, algo=ALGO_WFA_STANDARD, scheme=(0, 3, 1, 2))
    exp2 = np.array([oracle.wfa_gotoh_cost(q, d, 3, 1, 2) for q, d in pairs[:12]], np.int32)
    assert np.array_equal(exp2, r2.score[:12])


def test_standard_long_pairs(engine, oracle):
    """Pairs beyond what fits the on-chip staging (sequences read through L2): 30 kbp at 5 %."""
    from sequencealigning_b200 import ALGO_WFA_STANDARD
    rng = random.Random(6)
    pairs = [_related(rng, 30000, 0.05), _related(rng, 24000, 0.02)]
    b = _batch(pairs)
    r = engine.align(b, algo=ALGO_WFA_STANDARD)
    exp = [oracle.wfa_gotoh_cost(q, d) for q, d in pairs]
    assert r.score.tolist() == exp


def test_wfa_non_global_not_implemented(engine):
    from sequencealigning_b200 import ALGO_WFA, MODE_LOCAL, NOT_IMPLEMENTED
    r = engine.align(_batch([(b"ACGT", b"ACGA")]), algo=ALGO_WFA, mode=MODE_LOCAL)  # wfa.rs:26
    assert r.status[0] == NOT_IMPLEMENTED


def test_reference_stdout_from_the_traced_literal_kernel(engine, oracle):
    """sa_wfa_reference_stdout: the device run yields the lo/hi lines and the converged element, the text
    is the reference's, byte for byte against the oracle's printer -- also where the
    reference panics (text up to the panic) or never converges (the first lines of endless output)."""
    import random
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wfa_golden.json")))["vectors"]
    pairs = [(v["seq1"].encode(), v["seq2"].encode()) for v in vec[:150]] + [(b"ACGT", b"ACGA"), (b"A", b"C"), (b"A", b"A")]
    rng = random.Random(12)
    for _ in range(150):
        a = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 40)))
        pairs.append((a, mutate(rng, a, rng.choice([0.05, 0.2]), True, b"ACGT") or b"A"))
    seen = set()
    for a, b in pairs:
        text, st = engine.wfa_reference_stdout(a, b)
        exp, est = oracle.wfa_print(a, b)
        assert (text, st) == (exp, est), (a, b)
        seen.add(st)
    assert seen == {0, 1, 2}
