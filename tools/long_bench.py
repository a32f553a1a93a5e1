"""Developer check: affine NW on pairs outside the packed 16-bit range (the literal 32-bit kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
b = synth.random_pairs(n, L, 0.05, True, seed=0x5A05)
with Engine(0) as eng:
    for rep in range(2):
        t0 = time.perf_counter()
        r = eng.align(b)
        dt = time.perf_counter() - t0
    print("pairs", n, "len", L, "seconds", round(dt, 3), "GCUPS", round(b.cells / dt / 1e9, 2), "status", np.unique(r.status, return_counts=True),
          "cigar runs", int(r.cigar_len.sum()))


def cigar_score(q: bytes, d: bytes, words, match=5, mismatch=-4, gap_open=-8, gap_ext=-6):
    """Score of an alignment given as (len << 2 | op) runs; also checks that it consumes both sequences."""
    qa, da = np.frombuffer(q, np.uint8), np.frombuffer(d, np.uint8)
    y = x = 0
    score = 0
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


bad = 0
for i in range(b.n_pairs):
    if r.status[i] in (0, 1) and r.cigar_len[i]:
        s = cigar_score(b.query(i), b.db(i), r.cigar_of(i))
        bad += s != int(r.score[i])
print("alignments whose score recomputed from the CIGAR differs from the reported score:", bad)
assert bad == 0
