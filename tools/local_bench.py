#!/usr/bin/env python
"""Linear-NW LOCAL mode throughput (nw_local.cuh) on synthetic read pairs; checks a sample against the oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL, MODE_GLOBAL, Engine, synth
from oracle import binding as ob

def main():
    ob.build()
    with Engine(0) as eng:
        for n, L in ((200_000, 150), (100_000, 250), (2_000, 2000), (64, 20000)):
            b = synth.random_pairs(n, L, 0.05, True, seed=0x10CA1 + L)
            for cigar in (True, False):
                eng.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL, cigar=cigar)
                t0 = time.perf_counter()
                r = eng.align(b, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL, cigar=cigar)
                dt = time.perf_counter() - t0
                t = eng.timing()
                print(f"local {n} x {L} bp cigar={cigar}: {dt*1e3:.1f} ms wall, kernels {t['kernels_ms']:.1f} ms, "
                      f"{b.cells/ (t['kernels_ms']*1e-3)/1e9:.1f} GCUPS (kernels), {b.cells/dt/1e9:.1f} GCUPS (e2e)", flush=True)
            k = min(n, 300)
            sub = b.select(np.arange(k))
            ref = ob.linear_batch(sub.residues, sub.q_off, sub.q_len, sub.d_off, sub.d_len, cigar_stride=2 * L + 64, n_threads=8, local=True) if L <= 2000 else None
            if ref is not None:
                rr = eng.align(sub, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
                ok = np.array_equal(ref.score, rr.score) and np.array_equal(ref.end1, rr.end1) and np.array_equal(ref.cigar_len, rr.cigar_len)
                print("  parity vs oracle on", k, "pairs:", ok, flush=True)
        # global linear for comparison
        b = synth.random_pairs(200_000, 150, 0.05, True, seed=5)
        eng.align(b, algo=ALGO_NW_LINEAR)
        t0 = time.perf_counter(); eng.align(b, algo=ALGO_NW_LINEAR); dt = time.perf_counter() - t0
        print(f"global linear 200000 x 150: {b.cells/dt/1e9:.1f} GCUPS e2e")

if __name__ == "__main__":
    main()
