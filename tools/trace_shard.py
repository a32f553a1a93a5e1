"""Developer tool: where the wall time of ONE sharded sa_align_batch call goes (config 3 shape, 250 bp, 2-bit input):
SA_TRACE timeline of the per-device call + the shard report, for a shard of `pairs` pairs per listed device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, synth
from sequencealigning_b200.engine import PinnedResult, pin_batch
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
devices = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
batch = synth.random_pairs(pairs * len(devices), 250, 0.05, True, seed=0x5A03)
packed = pin_batch(batch.packed())
with Engine(devices=devices) as eng:
    pres = PinnedResult(batch.n_pairs, 40 * batch.n_pairs)
    for _ in range(3):
        eng.align(packed, out=pres)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        eng.align(packed, out=pres)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("wall ms per call", [round(t, 2) for t in ts], file=sys.stderr)
    print("timing", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in eng.timing().items()}, file=sys.stderr)
    for s in eng.shards():
        print("shard", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()}, file=sys.stderr)
    os.environ["SA_TRACE"] = "1"
    t0 = time.perf_counter()
    eng.align(packed, out=pres)
    print("traced call total ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
