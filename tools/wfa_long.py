"""Developer check: WFA (standard) on a few 100 kbp pairs -- BASELINE.json configs[4] shape --
against the CPU restatement on the first pairs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, synth, ALGO_WFA_STANDARD

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
check = int(sys.argv[3]) if len(sys.argv) > 3 else 1
b = synth.random_pairs(n, L, 0.05, True, seed=0x5A05)
with Engine(0) as eng:
    for rep in range(2):
        t0 = time.perf_counter()
        r = eng.align(b, algo=ALGO_WFA_STANDARD)
        dt = time.perf_counter() - t0
    print("pairs", n, "len", L, "seconds", round(dt, 3), "aln/s", round(n / dt, 2), "GCUPS-eq", round(b.cells / dt / 1e9, 1),
          "status", np.unique(r.status), "scores", r.score[:4])
if check:
    from oracle import binding
    for i in range(check):
        t0 = time.perf_counter()
        s = binding.wfa_standard(b.query(i), b.db(i))
        print("oracle", i, s, "gpu", int(r.score[i]), "cpu seconds", round(time.perf_counter() - t0, 2))
        assert s == int(r.score[i])
