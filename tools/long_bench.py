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
