"""CPU: oracle of the reference's gap-affine WFA (oracle/wfa.c).

Pins: the reference's OWN unit tests (wfa.rs:994-1186, 1268-1294) restated on the literal
Python model; the frozen corpus tests/golden/wfa_golden.json (status, printed score,
`lo/hi` lines, converged element); a fresh random cross-check C vs Python; and, for the
standard-mode algorithm the GPU uses on config-sized inputs, WFA == Gotoh cost DP.
"""
import json
import os
import random

import pytest

from oracle import literal_model as L

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "wfa_golden.json")
ST = {"OK": 0, "PANIC": 1, "NO_CONVERGENCE": 2}


def test_reference_unit_tests_on_the_literal_model():
    # test_wavefront_tensor_new_all_none (wfa.rs:994-1000)
    assert L.wf_tensor_new(None, None, None) is None
    # test_initial (wfa.rs:1104-1186): exact expected tensors
    initial = L.WfTensor(None, None, L.WaveFront(0, 0, [L.WfElement(0, [], "M")]))
    assert L.wf_tensor_new(initial, None, None).key() == (
        (1, 1, ((1, ("M",), "I"),)), (-1, -1, ((0, ("M",), "D"),)), (1, -1, ((0, ("D",), "M"), None, (1, ("I",), "M"))))
    assert L.wf_tensor_new(None, None, initial).key() == (None, None, (0, 0, ((1, ("M",), "M"),)))
    # recurrance_eq (wfa.rs:1002-1102): which components `new` reads
    def wf(hi, lo, n, st):
        return L.WaveFront(hi, lo, [L.WfElement(1, [], st) for _ in range(n)])
    full = L.WfTensor(wf(-1, 2, 4, "I"), wf(-2, 3, 1, "D"), wf(3, -2, 6, "I"))
    simple = L.WfTensor(None, None, wf(3, -2, 6, "I"))
    gap = L.WfTensor(wf(-1, 2, 4, "I"), wf(-2, 3, 1, "D"), None)
    k = lambda t: None if t is None else t.key()
    assert k(L.wf_tensor_new(simple, None, None)) == k(L.wf_tensor_new(full, None, None))
    assert k(L.wf_tensor_new(None, None, simple)) == k(L.wf_tensor_new(None, None, full))
    assert k(L.wf_tensor_new(None, gap, None)) == k(L.wf_tensor_new(None, full, None))
    # test_iteration (wfa.rs:1268-1286) and test_converge (:1288-1294)
    wfs = L.wfa_global_initial()
    for _ in range(6):
        L.wfa_expand(wfs, b"AAAATTTTCCCC", b"AAAATCTCC")
    assert L.wfa_global_initial()[-1].is_converged(b"AACATCAY", b"ATAGTAG") is None


def test_survey_known_answers(oracle):
    kats = [(b"ACGT", b"ACGA", 5), (b"ACGTA", b"ACGTC", 5), (b"GATTACA", b"GATCACT", 9), (b"GATTACAG", b"GATACAT", 13),
            (b"GATACAG", b"GATTACAT", 13), (b"AAAATTTTCCCC", b"AAAATCTCC", 25)]
    for q, d, score in kats:
        r = oracle.wfa_literal(q, d)
        assert (r.status, r.printed_score) == (oracle.OK, score)
    r, lines, _ = oracle.wfa_literal_ex(b"AAAATTTTCCCC", b"AAAATCTCC")
    assert lines == [(-1, 1), (-1, 1), (-2, 2), (-2, 2), (-2, 2), (-3, 3), (-3, 3), (-3, 3), (-4, 4)]
    for q in (b"ACGT", b"GATTACA"):  # identical sequences overshoot (n2-1, n1-1): never converge
        assert oracle.wfa_literal(q, q).status == oracle.REF_NO_CONVERGENCE


def test_golden_vectors(oracle):
    vec = json.load(open(GOLDEN))["vectors"]
    assert len(vec) >= 250 and {v["status"] for v in vec} == {"OK", "PANIC", "NO_CONVERGENCE"}
    for v in vec:
        q, d = v["seq1"].encode(), v["seq2"].encode()
        r, lines, ce = oracle.wfa_literal_ex(q, d)
        assert r.status == ST[v["status"]], v
        assert [list(x) for x in lines[:64]] == v["lo_hi"], v
        if v["status"] == "OK":
            assert r.printed_score == v["printed_score"]
            assert [ce[0], list(ce[2]), ce[1]] == v["converged"]
        if v["status"] == "PANIC":
            assert str(r.panic_line) in v["panic_site"]


def test_cross_check_with_literal_model(oracle):
    rng = random.Random(77)
    seen = set()
    for _ in range(250):
        n = rng.choice([rng.randint(0, 12), rng.randint(8, 60), rng.randint(100, 200)])
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        d = bytes((c if rng.random() > 0.12 else rng.choice(b"ACGT")) for c in q)
        if rng.random() < 0.5 and len(d) > 2:
            cut = rng.randrange(len(d))
            d = d[:cut] + d[cut + 1:]
        cap = oracle.wfa_literal_cap(len(q), len(d))
        o = L.wfa_align(q, d, max_score=cap)
        r = oracle.wfa_literal(q, d)
        seen.add(o.status)
        assert r.status == ST[o.status], (q, d)
        if o.status == "OK":
            assert r.printed_score == o.printed_score
    assert seen == {"OK", "PANIC", "NO_CONVERGENCE"}


def test_config_sized_pairs_panic(oracle):
    """on 150 bp+ inputs the reference dies in trim's rotate_left (wfa.rs:577/:603)."""
    from sequencealigning_b200 import synth
    b = synth.random_pairs(60, 150, 0.05, True, seed=3)
    score, status = oracle.wfa_literal_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len)
    assert (status == oracle.REF_PANIC).mean() > 0.9


def test_standard_wfa_equals_gotoh_cost(oracle):
    rng = random.Random(5)
    for _ in range(300):
        n = rng.randint(0, 80)
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        if rng.random() < 0.3:
            d = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 80)))
        else:
            d = bytes((c if rng.random() > 0.1 else rng.choice(b"ACGT")) for c in q)[: rng.randint(max(0, n - 6), n + 1)]
        assert oracle.wfa_standard(q, d) == oracle.wfa_gotoh_cost(q, d), (q, d)
    assert oracle.wfa_gotoh_cost(b"ACGT", b"ACGT") == 0
    assert oracle.wfa_gotoh_cost(b"ACGT", b"AGT") == 8      # one gap of length 1: o + e
    assert oracle.wfa_gotoh_cost(b"ACGT", b"ACTT") == 4     # one mismatch


SURVEY_A2 = ("lo: -1, hi: 1\nconverged with score 5: \nhuhu, diag: 0\nElement {\n\tstate: M\n\toffset: 3\n\tparents: [\n    M,\n]\n}\n"
             "\nscore: 5\nyeah, score: 1\nwell shit\nwell shit\nhuh\n\n\n\nAlignment {\n    seq1: [],\n    seq2: [],\n}\n")


def test_reference_stdout_text(oracle):
    """The `-a wfa` stdout (wfa.rs:251, :36, :650, :104-116, :667-678, :851, :38-39):
    the worked example, then the C printer against the object-literal Python model on the golden
    corpus and random tiny pairs, all three outcomes."""
    import json
    import os
    import random
    from oracle import literal_model as L
    text, st = oracle.wfa_print(b"ACGT", b"ACGA")
    assert st == oracle.OK and text == SURVEY_A2
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wfa_golden.json")))["vectors"]
    pairs = [(v["seq1"].encode(), v["seq2"].encode()) for v in vec]
    rng = random.Random(3)
    for _ in range(250):
        a = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 12)))
        b = bytearray(a)
        for _ in range(rng.randint(0, 3)):
            k = rng.randrange(len(b))
            r = rng.random()
            if r < 0.5:
                b[k] = rng.choice(b"ACGT")
            elif r < 0.75 and len(b) > 1:
                del b[k]
            else:
                b.insert(k, rng.choice(b"ACGT"))
        pairs.append((a, bytes(b)))
    seen = set()
    codes = {"OK": oracle.OK, "PANIC": oracle.REF_PANIC, "NO_CONVERGENCE": oracle.REF_NO_CONVERGENCE}
    for a, b in pairs:
        text, st = oracle.wfa_print(a, b)
        exp, o = L.wfa_stdout(a, b, oracle.wfa_literal_cap(len(a), len(b)))
        assert (text, st) == (exp, codes[o.status]), (a, b)
        seen.add(st)
        if st == oracle.OK:
            seen.update(w for w in ("ret", "extend", "open") if f"\n{w}\n" in text)
    assert {oracle.OK, oracle.REF_PANIC, oracle.REF_NO_CONVERGENCE, "open"} <= seen, seen
