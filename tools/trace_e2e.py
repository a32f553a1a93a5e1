"""Developer tool: SA_TRACE timeline of one streamed sa_align_batch call (1 M x 150 bp, 2-bit input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, synth
from sequencealigning_b200.engine import PinnedResult, pin_batch
batch = synth.random_pairs(1000000, 150, 0.05, True, seed=0x5A02)
packed = pin_batch(batch.packed())
with Engine(0) as eng:
    pres = PinnedResult(batch.n_pairs, 40_000_000)
    for _ in range(3):
        eng.align(packed, out=pres)
    os.environ["SA_TRACE"] = "1"
    t0 = time.perf_counter()
    eng.align(packed, out=pres)
    print("total ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
