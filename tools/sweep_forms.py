"""Kernel-form sweep (developer tool): times the device-resident affine path for each K:G form
(SA_FORCE_K / SA_FORCE_G) on one workload; checks results against the first."""
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
    forms = sys.argv[3].split(",") if len(sys.argv) > 3 else ["8:4", "19:8"]
    batch = synth.random_pairs(pairs, length, 0.05, True, seed=0x5A02)
    ref = None
    for f in forms:
        parts = f.split(":")
        for k in ("SA_FORCE_K", "SA_FORCE_G"):
            os.environ.pop(k, None)
        if parts[0] != "auto":
            os.environ["SA_FORCE_K"], os.environ["SA_FORCE_G"] = parts[0], parts[1]
        os.environ["SA_SEG_PAIRS"] = str(pairs)  # one segment: the fill launch is timed alone
        with Engine(0) as eng:
            rb = eng.upload(batch)
            for _ in range(2):
                rb.align()
            eng.synchronize()
            st = torch.cuda.ExternalStream(eng.stream)
            times, fills = [], []
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                rb.align()
                e1.record(st)
                e1.synchronize()
                times.append(e0.elapsed_time(e1))
                fills.append(eng.timing()["fill_ms"])
            r = rb.download()
            rb.free()
        if ref is None:
            ref = r
        same = np.array_equal(ref.score, r.score) and np.array_equal(ref.cigar, r.cigar) and np.array_equal(ref.status, r.status)
        print(json.dumps({"form": f, "step_ms": round(min(times), 3), "fill_ms": round(min(fills), 3),
                          "fill_gcups": round(batch.cells / min(fills) / 1e6, 1),
                          "step_gcups": round(batch.cells / min(times) / 1e6, 1), "same": bool(same)}), flush=True)


if __name__ == "__main__":
    main()
