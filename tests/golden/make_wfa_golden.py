"""Generates tests/golden/wfa_golden.json from oracle/literal_model.py (the literal Python
transliteration of /root/reference/src/wfa.rs).  It freezes (a) the expectations of the
reference's OWN unit tests (wfa.rs:994-1186,1268-1294) as data, checked when this script runs,
and (b) the outcome of wfa_align on a corpus of small pairs: status, printed score, the
`lo: .., hi: ..` lines and the converged element.

    python tests/golden/make_wfa_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import literal_model as L  # noqa: E402


def check_reference_unit_tests():
    # test_wavefront_tensor_new_all_none :994-1000
    assert L.wf_tensor_new(None, None, None) is None
    # test_initial :1104-1186
    initial = L.WfTensor(None, None, L.WaveFront(0, 0, [L.WfElement(0, [], "M")]))
    true_res_o = ((1, 1, ((1, ("M",), "I"),)), (-1, -1, ((0, ("M",), "D"),)),
                  (1, -1, ((0, ("D",), "M"), None, (1, ("I",), "M"))))
    true_res_m = (None, None, (0, 0, ((1, ("M",), "M"),)))
    assert L.wf_tensor_new(initial, None, None).key() == true_res_o
    assert L.wf_tensor_new(None, None, initial).key() == true_res_m
    # recurrance_eq :1002-1102 (note the deliberately inconsistent hi/lo of the fixtures)
    def wf(hi, lo, n, st):
        return L.WaveFront(hi, lo, [L.WfElement(1, [], st) for _ in range(n)])
    full = L.WfTensor(wf(-1, 2, 4, "I"), wf(-2, 3, 1, "D"), wf(3, -2, 6, "I"))
    simple = L.WfTensor(None, None, wf(3, -2, 6, "I"))
    simple_gap = L.WfTensor(wf(-1, 2, 4, "I"), wf(-2, 3, 1, "D"), None)
    k = lambda t: None if t is None else t.key()
    assert k(L.wf_tensor_new(simple, None, None)) == k(L.wf_tensor_new(full, None, None))
    assert k(L.wf_tensor_new(None, None, simple)) == k(L.wf_tensor_new(None, None, full))
    assert k(L.wf_tensor_new(None, simple_gap, None)) == k(L.wf_tensor_new(None, full, None))
    # test_iteration :1268-1286: six expands do not error
    wfs = L.wfa_global_initial()
    for _ in range(6):
        L.wfa_expand(wfs, b"AAAATTTTCCCC", b"AAAATCTCC")
    # test_converge :1288-1294
    assert L.wfa_global_initial()[-1].is_converged(b"AACATCAY", b"ATAGTAG") is None
    return {"test_initial_open": repr(true_res_o), "test_initial_mismatch": repr(true_res_m)}


def mutate(rng, s, rate):
    out = bytearray()
    for c in s:
        if rng.random() < rate:
            k = rng.random()
            if k < 0.5:
                out.append(rng.choice([b for b in b"ACGT" if b != c]))
            elif k < 0.75:
                out.append(c)
                out.append(rng.choice(b"ACGT"))
        else:
            out.append(c)
    return bytes(out)


def main():
    pinned = check_reference_unit_tests()
    rng = random.Random(20261019)
    pairs = [(b"ACGT", b"ACGA"), (b"ACGTA", b"ACGTC"), (b"GATTACA", b"GATCACT"), (b"GATTACAG", b"GATACAT"),
             (b"GATACAG", b"GATTACAT"), (b"AAAATTTTCCCC", b"AAAATCTCC"), (b"ACGT", b"ACGT"), (b"GATTACA", b"GATTACA"),
             (b"A", b"A"), (b"A", b"C"), (b"AC", b"A"), (b"", b""), (b"A", b""), (b"", b"A")]
    for _ in range(260):
        n = rng.choice([rng.randint(1, 14), rng.randint(10, 40), rng.randint(30, 160)])
        q = bytes(rng.choice(b"ACGT") for _ in range(n))
        d = mutate(rng, q, rng.choice([0.05, 0.15, 0.3]))
        pairs.append((q, d))
    vec = []
    for q, d in pairs:
        cap = min(8 * (len(q) + len(d)) + 64, 2048)
        o = L.wfa_align(q, d, max_score=cap)
        vec.append({"seq1": q.decode(), "seq2": d.decode(), "status": o.status, "printed_score": o.printed_score,
                    "panic_site": o.panic_site, "lo_hi": o.lo_hi[:64],
                    "converged": [o.converged[0], list(o.converged[1]), o.converged[2]] if o.converged else None})
    json.dump({"generator": "tests/golden/make_wfa_golden.py (oracle/literal_model.py)", "reference_unit_tests": pinned,
               "vectors": vec}, open(os.path.join(HERE, "wfa_golden.json"), "w"), indent=0)
    from collections import Counter
    print(len(vec), "vectors", Counter(v["status"] for v in vec))


if __name__ == "__main__":
    main()
