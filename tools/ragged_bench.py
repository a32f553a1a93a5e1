"""developer tool: throughput on a ragged batch (lengths uniform in [50, 300]) with and without
the per-segment shape bucketing"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencealigning_b200 import Engine, PairBatch

rng = np.random.default_rng(1)
n = 300000
ql = rng.integers(50, 301, n)
dl = np.clip(ql + rng.integers(-5, 6, n), 1, None)
res = rng.integers(0, 4, int(ql.sum() + dl.sum()), dtype=np.uint8)
res = np.frombuffer(b"ACGT", np.uint8)[res]
qo = np.cumsum(ql) - ql
do = int(ql.sum()) + np.cumsum(dl) - dl
b = PairBatch(res, qo, ql, do, dl)
for mode in ("2", "0"):
    os.environ["SA_SORT"] = mode
    with Engine(0) as eng:
        rb = eng.upload(b)
        rb.align(); rb.align()
        t0 = time.perf_counter()
        for _ in range(5):
            rb.align()
        eng.synchronize()
        dt = (time.perf_counter() - t0) / 5
        rb.free()
    print("SA_SORT", mode, "ms", round(dt * 1e3, 2), "GCUPS", round(b.cells / dt / 1e9, 1))
