"""Kernel-variant sweep (developer tool): times the device-resident affine path for each
(SA_FORCE_G, SA_ORMASK) on one workload and checks results against the first variant."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sequencealigning_b200 import Engine, synth  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    length = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    gs = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["4"])]
    masks = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0x0F"]
    batch = synth.random_pairs(pairs, length, 0.05, True, seed=0x5A02)
    ref = None
    for g in gs:
        for m in masks:
            os.environ["SA_FORCE_G"] = str(g)
            os.environ["SA_ORMASK"] = m
            with Engine(0) as eng:
                rb = eng.upload(batch)
                for _ in range(2):
                    rb.align()
                eng.synchronize()
                st = torch.cuda.ExternalStream(eng.stream)
                n = 8
                times, fills = [], []
                for _ in range(n):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    rb.align()
                    e1.record(st)
                    e1.synchronize()
                    times.append(e0.elapsed_time(e1))
                    fills.append(eng.timing()["fill_ms"])
                ms = min(times)
                r = rb.download()
                rb.free()
            if ref is None:
                ref = r
            same = np.array_equal(ref.score, r.score) and np.array_equal(ref.cigar, r.cigar) and np.array_equal(ref.status, r.status)
            print(json.dumps({"G": g, "ormask": m, "ms": round(ms, 3), "ms_med": round(float(np.median(times)), 3), "fill_ms": round(min(fills), 3), "fill_gcups": round(batch.cells / min(fills) / 1e6, 1), "gcups": round(batch.cells / ms / 1e6, 1), "same": bool(same)}), flush=True)


if __name__ == "__main__":
    main()
