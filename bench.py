#!/usr/bin/env python
"""bench.py -- headline benchmark: batched affine-gap NW (score + traceback), GCUPS.

Workload (BASELINE.json configs[1]): synthetic 150 bp read pairs at 5 % divergence
(sub:ins:del = 2:1:1), 1,000,000 pairs PER GPU; a "step" is one pass of the hot path over
the whole batch.  GCUPS = sum(n1*n2) / seconds / 1e9.

    python bench.py --gpus 1 --steps 5 --warmup 3            # this engine
    python bench.py --impl reference ...                      # CPU oracle of the reference
    torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU, no collective
                                                              # on the data path (weak scaling)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS, batched affine-gap NW score+traceback (alignments/s in config)"
OPS_PER_CELL_ISSUED = 8.0     # SASS lane-instructions per DP cell in the fill (16 per packed pair of cells)
OPS_PER_CELL_S32_EQUIV = 16.0  # SURVEY.md 8d / BASELINE.md figure for score + 4-bit traceback
PROBE_PAIRS = 262144          # one-segment launch used to time the fill kernel alone


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5", "config5nw"],
                    help="config2 (default, the headline): 150 bp affine NW; config3: 250 bp affine NW; "
                         "config4: WFA (standard mode) on 1-10 kbp pairs at 1-15 %% error; "
                         "config5: WFA (standard mode) on 1 k pairs of 100 kbp at 5 %%; "
                         "config5nw: affine NW on 1 k pairs of 100 kbp (score, status and CIGAR; the tiled long-pair path)")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (0 = the workload's default)")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--divergence", type=float, default=0.05)
    ap.add_argument("--no-indels", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--configs", default="all", choices=["all", "sharded", "none"],
                    help="with the default workload: also measure the other BASELINE.json configs at their stated sizes "
                         "(`configs` block, N = 1 only) and config 3 (10 M x 250 bp) from ONE host list through ONE "
                         "multi-device call (`sharded` block, every N)")
    ap.add_argument("--sharded-pairs", type=int, default=10_000_000)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


WORKLOADS = {  # name -> (default pairs per GPU, length, BASELINE.json index)
    "config2": (1_000_000, 150, 1),
    "config3": (1_000_000, 250, 2),
    "config4": (100_000, 0, 3),
    "config5": (1_000, 100_000, 4),
    "config5nw": (1_000, 100_000, 4),
}
WFA_WORKLOADS = ("config4", "config5")
DTYPES = {"config2": "u16x2 (two pairs per 32-bit register; exact integer)", "config3": "u16x2 (two pairs per 32-bit register; exact integer)",
          "config4": "s32 (wavefront offsets)", "config5": "s32 (wavefront offsets)", "config5nw": "s32 (4*V - 4*ext*(x+y) + provenance bits)"}


def resolve_workload(args):
    pairs, length, _ = WORKLOADS[args.workload]
    args.pairs = args.pairs or pairs
    args.length = args.length or length


def make_batch(args, rank: int):
    from sequencealigning_b200 import synth
    if args.workload == "config4":
        return synth.config4(args.pairs, seed=synth.SEEDS["config4"] + 7919 * rank)
    if args.workload in ("config5", "config5nw"):
        return synth.random_pairs(args.pairs, args.length, 0.05, True, seed=synth.SEEDS["config5"] + 7919 * rank)
    return synth.random_pairs(args.pairs, args.length, args.divergence, not args.no_indels,
                              seed=synth.SEEDS[args.workload] + 7919 * rank)


def workload_config(args, n_gpus: int) -> dict:
    if args.workload == "config5nw":
        return {"workload": f"affine NW, {args.pairs} synthetic pairs per GPU of {args.length} bp at 5 % divergence (sub:ins:del 2:1:1) "
                            f"(BASELINE.json configs[4], NW half): exact score, status and first alignment through the tiled 32-bit "
                            f"long-pair path (nw_long.cuh: one warp per tile, tile anti-diagonals over all SMs, checkpointed edges, "
                            f"region-wise traceback; the reference's -32768 sentinel is live at this length)",
                "pairs_per_gpu": args.pairs, "length": args.length, "n_gpus": n_gpus,
                "scheme": "match 5 / mismatch -4 / open -8 / ext -6 (nw_affine.rs:15-20)"}
    if args.workload == "config5":
        return {"workload": f"gap-affine WFA (standard mode, x=4 o=2 e=6), {args.pairs} synthetic pairs per GPU of {args.length} bp "
                            f"at 5 % divergence (sub:ins:del 2:1:1) (BASELINE.json configs[4], WFA half); GCUPS is EQUIVALENT cells n1*n2/s",
                "pairs_per_gpu": args.pairs, "length": args.length, "n_gpus": n_gpus,
                "note": "score only (an extension: the reference's own WFA produces no result on inputs of this size)"}
    if args.workload == "config4":
        return {"workload": f"gap-affine WFA (standard mode, x=4 o=2 e=6), {args.pairs} synthetic pairs per GPU of 1-10 kbp "
                            f"(log-uniform) at 1-15 % error (BASELINE.json configs[3]); GCUPS is EQUIVALENT cells n1*n2/s",
                "pairs_per_gpu": args.pairs, "n_gpus": n_gpus,
                "note": "the reference's own WFA panics (wfa.rs:577/603) or never converges (wfa.rs:189) on inputs of this size"}
    return {
        "workload": f"affine NW score+traceback, {args.pairs} synthetic {args.length} bp read pairs per GPU at "
                    f"{args.divergence:.0%} divergence ({'sub:ins:del 2:1:1' if not args.no_indels else 'substitutions only'}) "
                    f"(BASELINE.json configs[{WORKLOADS[args.workload][2]}])",
        "pairs_per_gpu": args.pairs, "length": args.length, "n_gpus": n_gpus,
        "scheme": "match 5 / mismatch -4 / open -8 / ext -6 (nw_affine.rs:15-20)",
        "sharding": "independent shards per rank, no collective on the data path",
        "l2": "inputs (~%.0f MB) + traceback scratch (GBs) exceed the 126 MB L2; no flush needed" % (args.pairs * args.length * 2 / 1e6),
    }


def cpu_baseline_wfa(batch, n_sample: int) -> dict:
    """config4: textbook gap-affine WFA on the CPU (oracle/wfa.c), 1 core, a bounded sample."""
    from oracle import binding as ob
    ob.build()
    n_sample = min(n_sample, batch.n_pairs)
    t0 = time.perf_counter()
    cells = 0
    for p in range(n_sample):
        ob.wfa_standard(batch.query(p), batch.db(p))
        cells += int(batch.q_len[p]) * int(batch.d_len[p])
    dt = time.perf_counter() - t0
    return {"value": cells / dt / 1e9, "unit": "GCUPS (equivalent cells)", "cores": 1, "kind": "port",
            "alignments_per_s": n_sample / dt, "seconds": dt,
            "sample": f"first {n_sample} pairs, oracle/wfa.c sao_wfa_standard (the reference's own wfa.rs produces no result on these inputs)"}


def cpu_baseline_long(batch, prefix: int = 20000) -> dict:
    """config5nw: the oracle's score-only affine DP (same recurrences and sentinel, O(n1) memory) on a
    prefix of the first pair, 1 core."""
    from oracle import binding as ob
    ob.build()
    q, d = batch.query(0)[:prefix], batch.db(0)[:prefix]
    t0 = time.perf_counter()
    ob.affine_score(q, d)
    dt = time.perf_counter() - t0
    return {"value": len(q) * len(d) / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"first {len(q)} x {len(d)} residues of pair 0, oracle/nw_affine.c sao_affine_score"}


def cpu_baseline(batch, n_sample: int, n_threads: int, min_seconds: float = 0.0) -> dict:
    """Times the CPU oracle (literal restatement of the reference) on a bounded sample."""
    from oracle import binding as ob
    ob.build()
    n_sample = min(n_sample, batch.n_pairs)
    idx = np.arange(n_sample)
    sub = batch.select(idx)
    stride = int((sub.q_len.astype(np.int64) + sub.d_len).max()) + 1
    t0 = time.perf_counter()
    reps = 0
    while True:
        ob.affine_batch(sub.residues, sub.q_off, sub.q_len, sub.d_off, sub.d_len, cigar_stride=stride, n_threads=n_threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    cells = sub.cells * reps
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": n_threads, "kind": "port",
            "alignments_per_s": n_sample * reps / dt, "seconds": dt,
            "sample": f"first {n_sample} pairs of the workload x {reps} pass(es), oracle/nw_affine.c "
                      f"(flat-array restatement; the Rust original allocates >=3 Rc per cell and is far slower)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port; the Rust crate cannot
    be built in this image) on the same workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = make_batch(args, 0)
    cores = os.cpu_count() or 1
    if args.workload in WFA_WORKLOADS:
        cb = cpu_baseline_wfa(batch, args.cpu_sample or (64 if args.workload == "config4" else 1))
        out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GCUPS", "n_gpus": args.gpus, "steps": 1,
               "warmup": 0, "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "i32", "data": "synthetic", "config": workload_config(args, args.gpus), "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(out), flush=True)
        return
    n_sample = args.cpu_sample or min(args.pairs, 4000 * cores)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(batch, min(n_sample, 2000), cores)
    vals = []
    for _ in range(args.steps):
        vals.append(cpu_baseline(batch, n_sample, cores))
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1]); cb["value"] = v
    ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "i32", "data": "synthetic", "config": workload_config(args, args.gpus), "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the NUMA
    node its GPU hangs off, so that N ranks do not share one socket's memory controllers."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = ids & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa node {node} ({len(allowed)} cpus)"
        return f"numa node {node}: no allowed cpus, not bound"
    except Exception as ex:  # containers often hide sysfs; binding is an optimisation only
        return f"numa: not bound ({type(ex).__name__})"


def int_peak() -> dict:
    """Measured integer issue rate (own microbenchmark, sequencealigning_b200/csrc/microbench)."""
    from sequencealigning_b200.build import INT_PEAK_PATH
    try:
        r = subprocess.run([INT_PEAK_PATH], capture_output=True, text=True, timeout=120)
        d = json.loads(r.stdout)["results"]
        return {"issue_lane_ops_per_s": max(d["iadd"]["lane_ops_per_s"], d["mix_lop3_imad"]["lane_ops_per_s"]),
                "alu_pipe_lane_ops_per_s": d["viaddmnmx_s16x2"]["lane_ops_per_s"], "source": "int_peak microbenchmark, this run"}
    except Exception as ex:  # nominal: 32 lanes/clk/SMSP * 4 * 148 * 1.965 GHz
        return {"issue_lane_ops_per_s": 32 * 4 * 148 * 1.965e9, "alu_pipe_lane_ops_per_s": 16 * 4 * 148 * 1.965e9,
                "source": f"nominal (microbenchmark failed: {ex})"}


def fill_kernel_probe(batch, device: int, reps: int = 7) -> dict:
    """The fill kernel timed alone: a second engine whose segment size is forced to the probe size,
    so the sub-batch is ONE fill launch; duration = CUDA events around that launch on its stream
    (sa_last_timing.fill_ms), median of `reps` after 3 warm-ups."""
    from sequencealigning_b200 import Engine
    n = min(batch.n_pairs, PROBE_PAIRS)
    sub = batch.select(np.arange(n, dtype=np.int64))
    old = os.environ.get("SA_SEG_PAIRS")
    os.environ["SA_SEG_PAIRS"] = str(n)
    try:
        with Engine(device) as eng:
            rb = eng.upload(sub)
            ms = []
            for k in range(3 + reps):
                rb.align()
                if k >= 3:
                    ms.append(eng.timing()["fill_ms"])
            rb.free()
    finally:
        if old is None:
            os.environ.pop("SA_SEG_PAIRS", None)
        else:
            os.environ["SA_SEG_PAIRS"] = old
    return {"pairs": int(n), "cells": int(sub.cells), "residue_bytes": int(sub.q_len.sum()) + int(sub.d_len.sum()), "ms": float(np.median(ms))}


def _max_over_ranks(x: float, world: int, dist, torch) -> float:
    if world <= 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x: float, world: int, dist, torch) -> float:
    if world <= 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def measure(eng, batch, algo, steps: int, warmup: int, ctx, e2e: bool = True, packed: bool = True) -> dict:
    """One workload on one engine: `value` leg (inputs resident in HBM, CUDA events on the engine's
    stream around `steps` passes) and the e2e leg through sa_align_batch with pinned HOST buffers
    (host wall clock, copies inside).  ctx = (torch, dist, world, local)."""
    torch, dist, world, local = ctx
    from sequencealigning_b200.engine import PinnedResult, pin_batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    cells_all = _sum_over_ranks(float(batch.cells), world, dist, torch)
    pairs_all = _sum_over_ranks(float(batch.n_pairs), world, dist, torch)
    rb = eng.upload(batch)
    for _ in range(warmup):
        rb.align(algo=algo)
    eng.synchronize()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    tim = {}
    e0.record(stream)
    for _ in range(steps):
        rb.align(algo=algo)
        tim = eng.timing()
        launches += tim["kernel_launches"]
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_step = _max_over_ranks(e0.elapsed_time(e1), world, dist, torch) / steps
    res_dev = rb.download()
    rb.free()
    out = {"value": cells_all / (ms_step * 1e-3) / 1e9, "ms_per_step": ms_step, "alignments_per_s": pairs_all / (ms_step * 1e-3),
           "gpu_launches": int(launches // max(steps, 1)), "timing": tim, "result": res_dev, "cells_all": cells_all}
    if not e2e:
        return out
    pres = PinnedResult(batch.n_pairs, int(res_dev.cigar.size) + 1024)

    def timed(host_batch, prepare=None):
        for _ in range(2):
            eng.align(prepare(host_batch) if prepare else host_batch, algo=algo, out=pres)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = eng.align(prepare(host_batch) if prepare else host_batch, algo=algo, out=pres)
        torch.cuda.synchronize()
        dt = _max_over_ranks((time.perf_counter() - t0) / steps, world, dist, torch)
        t = eng.timing()
        assert np.array_equal(r.score, res_dev.score)
        return {"value": cells_all / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(t["h2d_bytes"]),
                "d2h_bytes_per_step": int(t["d2h_bytes"]), "alignments_per_s": pairs_all / dt}

    pinned_bytes = pin_batch(batch)
    by_bytes = timed(pinned_bytes)
    by_bytes["input_format"] = "byte per residue (sa_batch_t.packing = 0), pinned"
    if not packed:
        e = by_bytes
    else:
        # The product's input format is what its packer writes (north star: "a packer that writes 2-bit DNA
        # into pinned batches"), sa_batch_t.packing = 1.  Three figures: packed input prepared outside the
        # timed region (like FASTA parsing), the packer INSIDE the timed region (bytes -> sa_pack_2bit_mt ->
        # align, every step), and the byte format.
        from sequencealigning_b200 import PairBatch, _capi
        t0 = time.perf_counter()
        pk = batch.packed()
        pack_s = time.perf_counter() - t0
        pinned_packed = pin_batch(pk)
        e = timed(pinned_packed)
        e["input_format"] = "2-bit packed residues (sa_batch_t.packing = 1), pinned"
        e["packer_host_seconds_untimed"] = pack_s
        lib = _capi.lib()
        src = pinned_bytes.residues

        def pack_then(b):   # the packer inside the timed region: pinned bytes -> pinned 2-bit image, all host threads
            rc = lib.sa_pack_2bit_mt(src.ctypes.data, src.size, b.residues.ctypes.data, 0)
            assert rc == 0
            return b

        e["packer_included"] = timed(pinned_packed, pack_then)
        e["packer_included"]["input_format"] = "bytes in pinned memory -> sa_pack_2bit_mt (all host threads) -> 2-bit, every step, inside the timed region"
        e["packer_included"]["packer_gb_per_s"] = src.size / max(pack_s, 1e-9) / 1e9
        e["byte_per_residue"] = by_bytes
    e["api"] = "sa_align_batch (C ABI) with pinned host buffers; host wall clock, max over ranks"
    out["e2e"] = e
    pres.free()
    return out


def parity_affine(batch, res, n_sample: int = 2000) -> dict:
    """Spot check inside the run: the first pairs of the workload against the CPU oracle, bit for bit."""
    from oracle import binding as ob
    ob.build()
    n = min(n_sample, batch.n_pairs)
    sub = batch.select(np.arange(n))
    stride = int((sub.q_len.astype(np.int64) + sub.d_len).max()) + 1
    ref = ob.affine_batch(sub.residues, sub.q_off, sub.q_len, sub.d_off, sub.d_len, cigar_stride=stride, n_threads=os.cpu_count() or 1)
    ok = bool(np.array_equal(ref.score, res.score[:n]) and np.array_equal(ref.status, res.status[:n])
              and np.array_equal(ref.cigar_len, res.cigar_len[:n])
              and all(list(ref.cigar_pool[p, :ref.cigar_len[p]]) == res.cigar_of(p) for p in range(0, n, max(1, n // 200))))
    return {"ok": ok, "sample": f"first {n} pairs vs oracle/nw_affine.c: score, status, CIGAR length (all), CIGAR words (every {max(1, n // 200)}th)"}


def roofline_fill(batch, ms_step, gcups, world, local, res_dev, args, peak, hbm_peak) -> dict:
    ipeak = peak["issue_lane_ops_per_s"]
    pr = fill_kernel_probe(batch, local)
    k_cups = pr["cells"] / (pr["ms"] * 1e-3)
    cells = batch.cells
    hbm_bytes = cells / 2.0 + batch.residues.size + batch.n_pairs * (24 + 17) + res_dev.cigar.size * 4
    return {
        "bound": "int-issue (integer max-plus DP: neither HBM nor tensor cores bind; DESIGN.md 4.1)",
        "kernel": "nw_affine_fill_s16<K,G> (150 bp: K=19 columns per lane, G=8 lanes per pair-of-pairs; 250 bp: K=16, G=16; single pass)",
        "achieved": k_cups * OPS_PER_CELL_ISSUED / 1e12, "peak": ipeak / 1e12, "unit": "T lane-instr/s",
        "frac": k_cups * OPS_PER_CELL_ISSUED / ipeak,
        "frac_is": "issue-slot share: issued recurrence instructions (8 per cell) / measured issue rate",
        "lane_instr_per_cell": OPS_PER_CELL_ISSUED,
        "launch": {"pairs": pr["pairs"], "cells": pr["cells"], "ms": pr["ms"], "gcups": k_cups / 1e9,
                   "timed": "one fill launch alone, CUDA events on its stream, median of 7 after 3 warm-ups"},
        "whole_step_frac": gcups * 1e9 * OPS_PER_CELL_ISSUED / world / ipeak,
        # SURVEY.md 8d's algorithmic count (16 s32 ops per cell for score + 4-bit traceback); the packed u16x2
        # recurrence does that work in 8 issued instructions, so this ratio can pass 1
        "s32_equivalent": {"ops_per_cell": OPS_PER_CELL_S32_EQUIV, "achieved": k_cups * OPS_PER_CELL_S32_EQUIV / 1e12,
                           "frac": k_cups * OPS_PER_CELL_S32_EQUIV / ipeak,
                           "note": "SURVEY 8d formula (16 s32-equivalent ops per cell / measured issue rate); above 1 because u16x2 does two cells "
                                   "per lane-op and the VIMNMX predicate outputs replace the compare/select ops, not because work is skipped"},
        "traffic": (13920 + 365) * pr["pairs"] if args.length == 150 else None,
        "traffic_detail": {"unit": "DRAM bytes per fill launch", "algorithmic_bytes_per_launch": pr["cells"] / 2.0 + pr["residue_bytes"],
                           "source": "profiles/ncu_fill_r02_final.md (ncu --set full: dram__bytes_write.sum 1.392 GB + dram__bytes_read.sum 36.5 MB per 100 k pairs = 14285 B per 150 bp pair)"},
        "per_gpu": True,
        "note": "achieved = cells x 8 issued lane-instructions per cell (16 per packed pair of cells: 2 adds, 5 VIMNMX, 1 XOR, 8 tie-bit sets) "
                "/ fill-kernel time; peak = measured issue rate, 32 lanes/clk/SMSP (" + peak["source"] + ")",
        "hbm": {"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak, "bytes": "sequences + 0.5 B/cell traceback written + results, whole step"},
    }


def roofline_long(cells: float, tim: dict, peak: dict) -> dict:
    """Tiled long-pair path: the forward launches (nw_long_fwd) timed by CUDA events on their stream
    inside the step (sa_last_timing.long_fwd_ms), 6 issued lane-instructions per cell."""
    ipeak = peak["issue_lane_ops_per_s"]
    fwd_ms = tim.get("long_fwd_ms", 0.0) or float("nan")
    cups = cells / (fwd_ms * 1e-3)
    return {"bound": "int-issue", "kernel": "nw_long_fwd (32-bit, one warp per tile, 16 columns per lane; one launch per tile anti-diagonal)",
            "achieved": cups * 6 / 1e12, "peak": ipeak / 1e12, "unit": "T lane-instr/s", "frac": cups * 6 / ipeak,
            "frac_is": "issue-slot share: 6 issued recurrence instructions per cell (LOP3, VIMNMX, IADD3, VIMNMX3, 2 VIADDMNMX) / measured issue rate",
            "lane_instr_per_cell": 6, "forward_ms": fwd_ms, "traceback_ms": tim.get("long_back_ms"), "forward_gcups": cups / 1e9,
            "s32_equivalent": {"ops_per_cell": 8, "frac": cups * 8 / ipeak, "note": "SURVEY 8d: 8 s32-equivalent ops per cell, score only"},
            "traffic": None, "timed": "all forward launches of the step, CUDA events on their stream (sa_last_timing.long_fwd_ms)"}


def roofline_wfa(peak: dict, wf: dict, ms_step: float) -> dict:
    """WFA: SURVEY 8d's unit is 8 ops per wavefront cell + 4 per 32 extended bases; the kernel counts both."""
    ipeak = peak["issue_lane_ops_per_s"]
    ops = 8.0 * wf["wavefront_cells"] + 4.0 * wf["extended_bases"] / 32.0
    return {"bound": "latency (per-pair score loop) / int-issue", "kernel": "wfa_standard_kernel (one warp per pair)",
            "achieved": ops / (ms_step * 1e-3) / 1e12, "peak": ipeak / 1e12, "unit": "T lane-ops/s", "frac": ops / (ms_step * 1e-3) / ipeak,
            "wavefront_cells_per_s": wf["wavefront_cells"] / (ms_step * 1e-3), "extended_bases_per_s": wf["extended_bases"] / (ms_step * 1e-3),
            "ops": "8 per wavefront cell + 4 per 32 extended bases (SURVEY.md 8d), counted by the kernel", "traffic": None}


def make_sharded_batch(n_pairs: int, length: int, threads: int = 8):
    """config 3 at its stated size: generated in 500 k-pair chunks on a thread pool (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from sequencealigning_b200 import PairBatch, synth
    chunk = 500_000
    counts = [min(chunk, n_pairs - i) for i in range(0, n_pairs, chunk)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda kv: synth.random_pairs(kv[1], length, 0.05, True, seed=synth.SEEDS["config3"] + 104729 * kv[0]), enumerate(counts)))
    base = 0
    res, qo, ql, do, dl = [], [], [], [], []
    for p in parts:
        res.append(p.residues)
        qo.append(p.q_off + np.uint64(base)); do.append(p.d_off + np.uint64(base))
        ql.append(p.q_len); dl.append(p.d_len)
        base += int(p.residues.size)
    return PairBatch(np.concatenate(res), np.concatenate(qo), np.concatenate(ql), np.concatenate(do), np.concatenate(dl))


def sharded_block(args, world: int, steps: int = 3) -> dict:
    """BASELINE.json configs[2]: 10 M x 250 bp pairs from ONE host list through ONE sa_align_batch call on a
    multi-device engine (sa_engine_create_multi) over all `world` GPUs: sharding, H2D, kernels, D2H and the
    gather into input order inside the timed region.  Runs in rank 0's process; the other ranks have closed
    their engines and wait on a CPU (gloo) barrier."""
    from sequencealigning_b200 import Engine, _capi
    from sequencealigning_b200.engine import PinnedResult, pin_batch
    n = args.sharded_pairs
    t0 = time.perf_counter()
    batch = make_sharded_batch(n, 250)
    gen_s = time.perf_counter() - t0
    packed = pin_batch(batch.packed())
    pinned_bytes = pin_batch(batch)
    pres = PinnedResult(n, 40 * n)
    out = {"workload": f"affine NW score+traceback, {n} synthetic 250 bp pairs at 5 % (sub:ins:del 2:1:1), ONE host pair list, "
                       f"ONE sa_align_batch call sharding it over {world} GPU(s) (BASELINE.json configs[2])",
           "n_gpus": world, "pairs": n, "cells": int(batch.cells), "generation_seconds_untimed": gen_s,
           "api": "sa_engine_create_multi + sa_align_batch, pinned host buffers; host wall clock around the call"}
    with Engine(devices=list(range(world))) as eng:
        def timed(hb, prepare=None):
            for _ in range(2):
                eng.align(prepare(hb) if prepare else hb, out=pres)
            ts = []
            for _ in range(steps):
                t0 = time.perf_counter()
                r = eng.align(prepare(hb) if prepare else hb, out=pres)
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            tim, shards = eng.timing(), eng.shards()
            cells = np.array([s["cells"] for s in shards], np.float64)
            return r, {"value": batch.cells / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3, "alignments_per_s": n / dt,
                       "h2d_bytes_per_step": int(tim["h2d_bytes"]), "d2h_bytes_per_step": int(tim["d2h_bytes"]),
                       "gpu_launches": int(tim["kernel_launches"]), "slowest_device_kernels_ms": tim["kernels_ms"],
                       "cell_imbalance_max_over_mean": float(cells.max() / cells.mean()),
                       "shards": [{"device": s["device"], "pairs": s["pairs"], "cells": s["cells"], "contiguous": s["contiguous"],
                                   "device_ms": round(s["device_ms"], 3), "host_ms": round(s["host_ms"], 3)} for s in shards]}
        r, e = timed(packed)
        out.update(e)
        out["input_format"] = "2-bit packed residues (sa_batch_t.packing = 1), pinned"
        lib = _capi.lib()
        src = pinned_bytes.residues

        def pack_then(b):
            assert lib.sa_pack_2bit_mt(src.ctypes.data, src.size, b.residues.ctypes.data, 0) == 0
            return b

        out["packer_included"] = timed(packed, pack_then)[1]
        out["packer_included"].pop("shards", None)
        out["byte_per_residue"] = timed(pinned_bytes)[1]
        out["byte_per_residue"].pop("shards", None)
        if not args.skip_cpu:
            out["parity"] = parity_affine(batch, r, 3000)
            out["cpu_baseline"] = cpu_baseline(batch, 5000, 1)
    pres.free()
    return out


def extra_configs(args, eng, ctx, peak) -> dict:
    """The other BASELINE.json configs at their stated sizes (N = 1): value, e2e, cpu_baseline, parity flag."""
    from oracle import binding as ob
    from sequencealigning_b200 import ALGO_NW_AFFINE, ALGO_WFA_STANDARD, PairBatch, synth
    out = {}
    # ---- config1: linear NW, 1 query x 1 000 db records of ~150 bp (the reference's own CPU-sized case) --------
    from sequencealigning_b200 import ALGO_NW_LINEAR, MODE_LOCAL
    b1 = synth.config1(1000)
    m = measure(eng, b1, ALGO_NW_LINEAR, steps=20, warmup=3, ctx=ctx, packed=True)
    r1 = m.pop("result"); m.pop("timing"); m.pop("cells_all")
    m.update({"unit": "GCUPS", "dtype": "u16x2 (two pairs per 32-bit register; exact integer)",
              "workload": "linear-gap (single-matrix) NW score + first hit, 1 query x 1000 synthetic 150 bp db sequences at 5 % "
                          "(BASELINE.json configs[0]); one sa_align_batch call on a live engine -- the sa_align CLI adds 1.3-3.8 s of "
                          "CUDA context start-up per process (profiles/cli_bench_r02.json)"})
    t0 = time.perf_counter()
    eng.align(b1, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    t0 = time.perf_counter()
    rl = eng.align(b1, algo=ALGO_NW_LINEAR, mode=MODE_LOCAL)
    m["local_mode"] = {"ms_per_call": (time.perf_counter() - t0) * 1e3, "note": "the same pairs with -m local (nw_local.cuh), host buffers, wall clock"}
    if not args.skip_cpu:
        ob.build()
        t0 = time.perf_counter()
        ref = ob.linear_batch(b1.residues, b1.q_off, b1.q_len, b1.d_off, b1.d_len, cigar_stride=320, n_threads=1)
        dt = time.perf_counter() - t0
        refl = ob.linear_batch(b1.residues, b1.q_off, b1.q_len, b1.d_off, b1.d_len, cigar_stride=320, n_threads=4, local=True)
        ok = (np.array_equal(ref.score, r1.score) and np.array_equal(ref.cigar_len, r1.cigar_len)
              and all(ref.cigar(p) == r1.cigar_of(p) for p in range(0, 1000, 10))
              and np.array_equal(refl.score, rl.score) and np.array_equal(refl.end1, rl.end1) and np.array_equal(refl.end2, rl.end2)
              and np.array_equal(refl.cigar_len, rl.cigar_len))
        m["parity"] = {"ok": bool(ok), "sample": "all 1000 pairs vs oracle/nw_linear.c: score, CIGAR length (all), CIGAR words (every 10th); "
                                                 "local mode: score, start cell, CIGAR length (all)"}
        m["cpu_baseline"] = {"value": b1.cells / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "seconds": dt,
                             "sample": "all 1000 pairs, oracle/nw_linear.c (the reference itself walks chars().nth() per cell, O(n) each)"}
    out["config1"] = m
    # ---- config4: WFA, 100 k pairs of 1-10 kbp --------------------------------------------------------
    t0 = time.perf_counter()
    b4 = synth.config4(100_000)
    m = measure(eng, b4, ALGO_WFA_STANDARD, steps=3, warmup=1, ctx=ctx, packed=False)
    r4 = m.pop("result"); m.pop("timing"); m.pop("cells_all")
    # the work counters come back with the host-buffer call (sa_last_timing.wfa_cells / wfa_extended)
    eng.align(b4, algo=ALGO_WFA_STANDARD, cigar=False)
    wt = eng.timing()
    m["roofline"] = roofline_wfa(peak, {"wavefront_cells": wt["wfa_cells"], "extended_bases": wt["wfa_extended"]}, m["ms_per_step"])
    m.update({"unit": "GCUPS (equivalent cells n1*n2)", "dtype": DTYPES["config4"],
              "workload": "gap-affine WFA (standard mode, x=4 o=2 e=6), 100000 synthetic pairs of 1-10 kbp (log-uniform) at 1-15 % error "
                          "(BASELINE.json configs[3]); the reference's own wfa.rs panics or never converges on these inputs: an extension, "
                          "graded against the Gotoh cost DP"})
    if not args.skip_cpu:
        ob.build()
        idx = list(range(0, 100_000, 100_000 // 12))[:12]
        ok = all(int(r4.score[p]) == ob.wfa_standard(b4.query(p), b4.db(p)) for p in idx)
        m["parity"] = {"ok": bool(ok), "sample": "12 pairs spread over the batch vs oracle/wfa.c sao_wfa_standard (itself checked against the Gotoh cost DP in tests)"}
        m["cpu_baseline"] = cpu_baseline_wfa(b4, 48)
    out["config4"] = m
    del b4
    # ---- config5 (WFA half): 1 k pairs of 100 kbp --------------------------------------------------------
    b5 = synth.random_pairs(1000, 100_000, 0.05, True, seed=synth.SEEDS["config5"])
    m = measure(eng, b5, ALGO_WFA_STANDARD, steps=2, warmup=1, ctx=ctx, packed=False)
    r5 = m.pop("result"); m.pop("timing"); m.pop("cells_all")
    eng.align(b5, algo=ALGO_WFA_STANDARD, cigar=False)
    wt = eng.timing()
    m["roofline"] = roofline_wfa(peak, {"wavefront_cells": wt["wfa_cells"], "extended_bases": wt["wfa_extended"]}, m["ms_per_step"])
    m.update({"unit": "GCUPS (equivalent cells n1*n2)", "dtype": DTYPES["config5"],
              "workload": "gap-affine WFA (standard mode), 1000 synthetic pairs of 100 kbp at 5 % (BASELINE.json configs[4], WFA half); score only"})
    if not args.skip_cpu:
        pre = PairBatch.from_pairs([(b5.query(p)[:12000], b5.db(p)[:12000]) for p in range(3)])
        rp = eng.align(pre, algo=ALGO_WFA_STANDARD, cigar=False)
        t0 = time.perf_counter()
        exp = [ob.wfa_standard(pre.query(p), pre.db(p)) for p in range(3)]
        dt = time.perf_counter() - t0
        m["parity"] = {"ok": bool([int(v) for v in rp.score] == exp), "sample": "12 kbp prefixes of the first 3 pairs vs oracle/wfa.c sao_wfa_standard"}
        m["cpu_baseline"] = {"value": pre.cells / dt / 1e9, "unit": "GCUPS (equivalent cells)", "cores": 1, "kind": "port", "seconds": dt,
                             "sample": "12 kbp prefixes of the first 3 pairs, oracle/wfa.c sao_wfa_standard"}
    out["config5_wfa"] = m
    # ---- config5 (NW half): the same 1 k pairs through the tiled long-pair path ----------------------------
    m = measure(eng, b5, ALGO_NW_AFFINE, steps=2, warmup=1, ctx=ctx, packed=True)
    r5n = m.pop("result"); tim = m.pop("timing"); cells = m.pop("cells_all")
    m.update({"unit": "GCUPS", "dtype": DTYPES["config5nw"], "roofline": roofline_long(cells, tim, peak),
              "pairs_fallback_literal_kernel": int(tim.get("pairs_fallback", 0)),
              "workload": "affine NW score + status + CIGAR, 1000 synthetic pairs of 100 kbp at 5 % (BASELINE.json configs[4], NW half), "
                          "tiled long-pair path (nw_long.cuh)"})
    if not args.skip_cpu:
        cb = cpu_baseline_long(b5)
        q, d = b5.query(0)[:20000], b5.db(0)[:20000]
        pre = PairBatch.from_pairs([(q, d)] + [(b5.query(p)[:2600], b5.db(p)[:2500]) for p in range(1, 5)])
        rp = eng.align(pre)
        small = pre.select(np.arange(1, 5))
        ref = ob.affine_batch(small.residues, small.q_off, small.q_len, small.d_off, small.d_len, cigar_stride=5200, n_threads=4)
        ok = (int(rp.score[0]) == ob.affine_score(q, d) and np.array_equal(ref.score, rp.score[1:]) and np.array_equal(ref.status, rp.status[1:])
              and all(list(ref.cigar_pool[p, :ref.cigar_len[p]]) == rp.cigar_of(p + 1) for p in range(4)))
        m["parity"] = {"ok": bool(ok), "sample": "20 kbp prefix of pair 0: score vs the oracle's score-only DP; 2.6 kbp prefixes of pairs 1-4: score, status, "
                                                 "CIGAR vs oracle/nw_affine.c (100 kbp pairs are pinned in tests/test_gpu_affine.py)"}
        m["cpu_baseline"] = cb
    out["config5_nw"] = m
    return out


def main():
    args = parse_args()
    resolve_workload(args)
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE line, the JSON: everything else a library prints there (NCCL's version banner
    # goes to stdout whatever NCCL_DEBUG_FILE says) is sent to stderr by pointing fd 1 at fd 2 for the run.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        os.write(json_fd, (line + "\n").encode())

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        # stdout carries the one JSON line; NCCL's log (NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")   # CPU-side barrier for the sharded block (no kernel waits on a GPU)

    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank"
    from sequencealigning_b200 import ALGO_NW_AFFINE, ALGO_WFA_STANDARD, Engine
    from sequencealigning_b200.build import build_all
    build_all()
    eng = Engine(local)
    batch = make_batch(args, rank)
    ctx = (torch, dist, world, local)
    algo = ALGO_WFA_STANDARD if args.workload in WFA_WORKLOADS else ALGO_NW_AFFINE
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi start-up happens during warm-up, not in the timed region
        time.sleep(0.5)
    m = measure(eng, batch, algo, args.steps, args.warmup, ctx, e2e=not args.skip_e2e, packed=args.workload not in WFA_WORKLOADS)
    clocks = sampler.stop() if rank == 0 else None
    res_dev, tim, cells_all = m.pop("result"), m.pop("timing"), m.pop("cells_all")
    e2e = m.get("e2e")
    if e2e is not None:
        e2e["host_binding"] = numa
    gcups, ms_step = m["value"], m["ms_per_step"]

    out = None
    if rank == 0:
        peak = int_peak()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        out = {
            "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.workload],
            "data": "synthetic", "config": workload_config(args, world), "alignments_per_s": m["alignments_per_s"],
            "gpu_launches": m["gpu_launches"], "pairs_rerun_per_step": int(tim.get("pairs_rerun", 0)),
            "clocks": clocks, "e2e": e2e, "roofline": None,
        }
        if args.workload == "config5nw":
            out["roofline"] = roofline_long(float(batch.cells), tim, peak)
            if not args.skip_cpu:
                out["cpu_baseline"] = cpu_baseline_long(batch)
        elif args.workload in WFA_WORKLOADS:
            out["roofline"] = {"bound": "latency (wavefront dependency chain)", "achieved": None, "peak": None, "unit": None, "frac": None,
                               "traffic": None, "note": "WFA does O(s^2) work, not n1*n2: GCUPS here is equivalent cells"}
            if not args.skip_cpu:
                out["cpu_baseline"] = cpu_baseline_wfa(batch, args.cpu_sample or (64 if args.workload == "config4" else 1))
        else:
            out["roofline"] = roofline_fill(batch, ms_step, gcups, world, local, res_dev, args, peak, hbm_peak)
            if not args.skip_cpu:
                out["parity"] = parity_affine(batch, res_dev)
                n_sample = args.cpu_sample or 20000
                out["cpu_baseline"] = cpu_baseline(batch, n_sample, 1)
                cores = os.cpu_count() or 1
                allc = cpu_baseline(batch, n_sample * min(cores, 8), cores)
                out["cpu_baseline"]["all_cores"] = {"value": allc["value"], "cores": cores, "unit": "GCUPS"}
        # the other BASELINE.json configs at their stated sizes (one GPU; rank 0's engine)
        if args.configs == "all" and args.workload == "config2" and world == 1:
            try:
                out["configs"] = extra_configs(args, eng, ctx, peak)
            except Exception as ex:   # an extra block must never cost the headline line
                out["configs"] = {"error": f"{type(ex).__name__}: {ex}"}
    del res_dev
    eng.close()
    torch.cuda.empty_cache()
    # ---- config 3 from ONE host list through ONE multi-device call (rank 0 drives all GPUs) ----------------
    if args.configs in ("all", "sharded") and args.workload == "config2":
        if world > 1:
            dist.barrier(group=gloo)     # every rank has released its GPU
        if rank == 0:
            try:
                out["sharded"] = sharded_block(args, world)
            except Exception as ex:
                out["sharded"] = {"error": f"{type(ex).__name__}: {ex}"}
        if world > 1:
            dist.barrier(group=gloo)
    if rank == 0:
        sys.stdout.flush()
        emit(json.dumps(out))
    if world > 1:
        dist.barrier(group=gloo)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
