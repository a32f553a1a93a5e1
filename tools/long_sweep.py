"""Developer sweep of the tiled long-pair path (nw_long.cuh): tile shape and register allocation."""
import itertools, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
b = synth.random_pairs(n, L, 0.05, True, seed=0x5A05)
base = None
for minb, (S, R) in itertools.product((4, 5), ((4, 1024), (2, 1024), (4, 512), (8, 1024), (2, 2048), (1, 2048))):
    os.environ.update(SA_LONG_MINB=str(minb), SA_LONG_S=str(S), SA_LONG_R=str(R))
    with Engine(0) as eng:
        rb = eng.upload(b)
        ts = []
        for rep in range(3):
            t0 = time.perf_counter(); rb.align(); eng.synchronize(); ts.append(time.perf_counter() - t0)
        tim = eng.timing()
        r = rb.download(); rb.free()
    if base is None:
        base = r
    same = bool(np.array_equal(base.score, r.score) and np.array_equal(base.cigar_len, r.cigar_len))
    print(json.dumps({"minb": minb, "S": S, "R": R, "s": round(min(ts), 4), "gcups": round(b.cells / min(ts) / 1e9, 1),
                      "fwd_ms": round(tim["long_fwd_ms"], 1), "back_ms": round(tim["long_back_ms"], 1), "same": same}), flush=True)
