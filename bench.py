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
    "config4": (20_000, 0, 3),
    "config5": (1_000, 100_000, 4),
    "config5nw": (1_000, 100_000, 4),
}
WFA_WORKLOADS = ("config4", "config5")


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
                "note": "score only; the reference's own WFA produces no result on inputs of this size; the affine-NW half of configs[4] "
                        "runs through the exact but untiled long-pair kernel and is not a bench line"}
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
    (sa_last_timing.walk_ms), median of `reps` after 3 warm-ups."""
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
                    ms.append(eng.timing()["walk_ms"])
            rb.free()
    finally:
        if old is None:
            os.environ.pop("SA_SEG_PAIRS", None)
        else:
            os.environ["SA_SEG_PAIRS"] = old
    return {"pairs": int(n), "cells": int(sub.cells), "residue_bytes": int(sub.q_len.sum()) + int(sub.d_len.sum()), "ms": float(np.median(ms))}


def main():
    args = parse_args()
    resolve_workload(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank"
    from sequencealigning_b200 import Engine
    from sequencealigning_b200.build import build_all
    build_all()
    eng = Engine(local)
    batch = make_batch(args, rank)
    cells = batch.cells
    stream = torch.cuda.ExternalStream(eng.stream, device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value` ------------------------------------------
    from sequencealigning_b200 import ALGO_NW_AFFINE, ALGO_WFA_STANDARD
    algo = ALGO_WFA_STANDARD if args.workload in WFA_WORKLOADS else ALGO_NW_AFFINE
    rb = eng.upload(batch)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi start-up happens during warm-up, not in the timed region
        time.sleep(0.5)
    for _ in range(args.warmup):
        rb.align(algo=algo)
    eng.synchronize()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    reruns = 0
    e0.record(stream)
    for _ in range(args.steps):
        rb.align(algo=algo)
        t = eng.timing()
        launches += t["kernel_launches"]
        reruns = t["pairs_rerun"]
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        ct = torch.tensor([float(cells)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        cells_all = float(ct.item())
    else:
        cells_all = float(cells)
    ms_step = ms_total / args.steps
    gcups = cells_all / (ms_step * 1e-3) / 1e9
    res_dev = rb.download()
    rb.free()

    # ---------------- end-to-end leg through the reference-facing call: `e2e` ----------------
    e2e = None
    clocks = None
    if args.skip_e2e and rank == 0:
        clocks = sampler.stop()
    if not args.skip_e2e:
        from sequencealigning_b200.engine import PinnedResult, pin_batch
        cap = int(res_dev.cigar.size) + 1024
        pres = PinnedResult(batch.n_pairs, cap)

        def timed_e2e(host_batch):
            for _ in range(2):
                eng.align(host_batch, algo=algo, out=pres)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r = eng.align(host_batch, algo=algo, out=pres)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            tim = eng.timing()
            if world > 1:
                tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            assert np.array_equal(r.score, res_dev.score)
            return {"value": cells_all / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": int(tim["h2d_bytes"]), "d2h_bytes_per_step": int(tim["d2h_bytes"]),
                    "alignments_per_s": args.pairs * world / dt}

        by_bytes = timed_e2e(pin_batch(batch))
        clocks = sampler.stop() if rank == 0 else None
        if args.workload in WFA_WORKLOADS:
            e2e = by_bytes
            e2e["input_format"] = "byte per residue (sa_batch_t.packing = 0)"
        else:
            # The product's input is what its packer writes (north star: "a packer that writes 2-bit
            # DNA ... into pinned batches"): sa_batch_t.packing = 1, produced by sa_pack_2bit before
            # the timed region, like FASTA parsing.  The byte-per-residue format is timed beside it.
            t0 = time.perf_counter()
            packed = batch.packed()
            pack_s = time.perf_counter() - t0
            e2e = timed_e2e(pin_batch(packed))
            e2e["input_format"] = "2-bit packed residues (sa_batch_t.packing = 1), pinned"
            e2e["packer_host_seconds_untimed"] = pack_s
            e2e["byte_per_residue"] = by_bytes
        e2e["api"] = "sa_align_batch (C ABI) with pinned host buffers; host wall clock, max over ranks"
        e2e["host_binding"] = numa

    if rank == 0:
        peak = int_peak()
        step_issued = gcups * 1e9 * OPS_PER_CELL_ISSUED / world
        tb_bytes = cells / 2.0
        hbm_bytes = tb_bytes + batch.residues.size + batch.n_pairs * (24 + 17) + res_dev.cigar.size * 4
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ipeak = peak["issue_lane_ops_per_s"]
        roof = None
        if args.workload not in WFA_WORKLOADS and args.workload != "config5nw":
            pr = fill_kernel_probe(batch, local)
            k_cups = pr["cells"] / (pr["ms"] * 1e-3)
            roof = {
                "bound": "int-issue (integer max-plus DP: neither HBM nor tensor cores bind; DESIGN.md 4.1)",
                "kernel": "nw_affine_fill_s16<K,G> (150 bp: K=19 columns per lane, G=8 lanes per pair-of-pairs, single pass)",
                "achieved": k_cups * OPS_PER_CELL_ISSUED / 1e12, "peak": ipeak / 1e12, "unit": "T lane-instr/s",
                "frac": k_cups * OPS_PER_CELL_ISSUED / ipeak,
                "lane_instr_per_cell": OPS_PER_CELL_ISSUED,
                "launch": {"pairs": pr["pairs"], "cells": pr["cells"], "ms": pr["ms"], "gcups": k_cups / 1e9,
                           "timed": "one fill launch alone, CUDA events on its stream, median of 7 after 3 warm-ups"},
                "whole_step_frac": step_issued / ipeak,
                # SURVEY.md 8d's algorithmic count (16 s32 ops per cell for score + 4-bit traceback); the packed
                # u16x2 recurrence does that work in 8 issued instructions, so this ratio can pass 1
                "s32_equivalent": {"ops_per_cell": OPS_PER_CELL_S32_EQUIV, "achieved": k_cups * OPS_PER_CELL_S32_EQUIV / 1e12,
                                   "frac": k_cups * OPS_PER_CELL_S32_EQUIV / ipeak},
                "traffic": (14133 + 325) * pr["pairs"] if args.length == 150 else None,
                "traffic_detail": {"unit": "DRAM bytes per fill launch",
                                   "algorithmic_bytes_per_launch": pr["cells"] / 2.0 + pr["residue_bytes"],
                                   "source": "profiles/ncu_fill_r01.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum = 14458 B per 150 bp pair: 96-bit rows of 19 four-bit cells, rows padded to the tile)"},
                "per_gpu": True,
                "note": "achieved = cells x 8 issued lane-instructions per cell (16 per packed pair of cells: 2 adds, 5 VIMNMX, 1 XOR, 8 tie-bit sets) "
                        "/ fill-kernel time, i.e. the share of all issue slots doing recurrence work; peak = measured issue rate, 32 lanes/clk/SMSP (" + peak["source"] + ")",
                "hbm": {"achieved": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak,
                        "bytes": "sequences + 0.5 B/cell traceback written + results, whole step"},
            }
        out = {
            "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16x2 (packed pairs; exact integer)",
            "data": "synthetic", "config": workload_config(args, world),
            "alignments_per_s": args.pairs * world / (ms_step * 1e-3),
            "gpu_launches": int(launches // max(args.steps, 1)), "pairs_rerun_per_step": int(reruns),
            "clocks": clocks, "e2e": e2e,
            "roofline": roof,
        }
        if args.workload == "config5nw":
            out["roofline"] = {"bound": "int-issue", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                               "note": "literal 32-bit recurrences + DFS first-event bookkeeping, ~140 lane-instructions per cell "
                                       "(nw_general.cuh); not the packed hot kernel, no roofline claimed"}
            if not args.skip_cpu:
                out["cpu_baseline"] = cpu_baseline_long(batch)
        elif args.workload in WFA_WORKLOADS:
            out["roofline"] = {"bound": "latency (wavefront dependency chain)", "achieved": None, "peak": None, "unit": None,
                               "frac": None, "traffic": None,
                               "note": "WFA does O(s^2) work, not n1*n2: GCUPS here is equivalent cells; no roofline is claimed this round"}
            if not args.skip_cpu:
                out["cpu_baseline"] = cpu_baseline_wfa(batch, args.cpu_sample or (64 if args.workload == "config4" else 1))
        elif not args.skip_cpu:
            n_sample = args.cpu_sample or 20000
            out["cpu_baseline"] = cpu_baseline(batch, n_sample, 1)
            cores = os.cpu_count() or 1
            allc = cpu_baseline(batch, n_sample * min(cores, 8), cores)
            out["cpu_baseline"]["all_cores"] = {"value": allc["value"], "cores": cores, "unit": "GCUPS"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
