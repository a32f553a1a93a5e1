"""developer tool: a small mixed workload for compute-sanitizer (memcheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencealigning_b200 import Engine, PairBatch, synth, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD
from tests.util import random_pair_list

pairs = random_pair_list(5, 300, 0, 120) + [(b"", b""), (b"A", b""), (b"", b"ACGT")]
b = PairBatch.from_pairs(pairs)
u = synth.random_pairs(600, 150, 0.05, True, seed=3)
for g in ("1", "4", "32"):
    os.environ["SA_FORCE_G"] = g
    os.environ["SA_SEG_PAIRS"] = "200"
    with Engine(0) as eng:
        for batch in (b, u, u.packed()):
            for algo in (0, ALGO_NW_LINEAR, ALGO_WFA, ALGO_WFA_STANDARD):
                r = eng.align(batch, algo=algo)
        rb = eng.upload(u); rb.align(); rb.download(); rb.free()
print("sanitize workload done")
