"""Generates tests/golden/affine_golden.json and linear_golden.json.

The vectors come from oracle/literal_model.py -- the object-graph-literal Python
transliteration of the reference's Rust (needleman_wunsch_affine.rs, needleman_wunsch.rs) --
because the Rust itself cannot be executed in this image (no cargo/rustc).  They freeze, per
pair: score, every alignment the reference would print (in order), whether and where it would
panic, and the printed text.  Both the C oracle and the CUDA engine are tested against them.

    python tests/golden/make_affine_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import literal_model as L  # noqa: E402


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
    rng = random.Random(20261018)
    pairs = [
        (b"ACGT", b"ACGT"), (b"ACGT", b"AGT"), (b"AGT", b"ACGT"), (b"ACGTT", b"ACGT"), (b"AAAA", b"AAA"),
        (b"ACGTACGT", b"ACGGT"), (b"GACGT", b"ACGT"), (b"ACGT", b"GACGT"), (b"", b""), (b"A", b""), (b"", b"AC"),
        (b"A", b"A"), (b"A", b"C"), (b"NNAN", b"NNNN"), (b"ACGTACGTAC", b"TTTTTTTT"), (b"AAAAAAAA", b"AAAA"),
        (b"GATTACA", b"GATCACT"), (b"AAAATTTTCCCC", b"AAAATCTCC"), (b"TACGT", b"ACGT"), (b"CCCCC", b"CCCC"),
        # the reference prints its first alignment(s) and THEN panics (late panic)
        (b"AGTGTGG", b"CTATGAAAGACT"), (b"ATATTGAACCGCGG", b"TGGTGTGTATCCT"), (b"CCGCCAA", b"TTCAATTCC"),
        (b"AACTTCA", b"CGGGTATCCGAG"),
    ]
    for _ in range(140):
        n = rng.randint(1, 36)
        q = bytes(rng.choice(b"ACGTN") for _ in range(n))
        d = mutate(rng, q, rng.choice([0.05, 0.2, 0.4])) if rng.random() < 0.8 else bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 36)))
        pairs.append((q, d))
    aff, lin = [], []
    for q, d in pairs:
        o = L.affine_align(q, d, max_pops=400000)
        if o.truncated:
            continue
        aff.append({
            "seq1": q.decode(), "seq2": d.decode(), "score": o.score, "panicked": o.panicked, "panic_site": o.panic_site,
            "n_printed": len(o.alignments),
            "first_cigar": L.columns_to_cigar(*o.alignments[0]) if o.alignments else [],
            "alignments": [[a.decode(), b.decode()] for a, b in o.alignments[:8]],
            "stdout_first": o.stdout.split("alignment found\n")[1] if o.alignments else "",
        })
        lo = L.linear_align(q, d, local=False, max_hits=5000)
        if not lo.truncated:
            lin.append({"seq1": q.decode(), "seq2": d.decode(), "score": lo.score, "n_hits": len(lo.hits),
                        "first_hit": list(lo.hits[0]) if lo.hits else None, "last_row": lo.scores[-1]})
    json.dump({"generator": "tests/golden/make_affine_golden.py (oracle/literal_model.py)", "vectors": aff},
              open(os.path.join(HERE, "affine_golden.json"), "w"), indent=0)
    json.dump({"generator": "tests/golden/make_affine_golden.py (oracle/literal_model.py)", "vectors": lin},
              open(os.path.join(HERE, "linear_golden.json"), "w"), indent=0)
    print(len(aff), "affine vectors,", len(lin), "linear vectors")


if __name__ == "__main__":
    main()
