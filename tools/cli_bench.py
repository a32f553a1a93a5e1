#!/usr/bin/env python
"""FASTA file -> results wall time through the sa_align CLI (parse + pack + align [+ print]), beside the CPU
oracle doing the same job (parse + align) on a bounded sample.  Config 1 (1 query x 1 000 db, linear NW, the
reference's CPU case) and 1 query x 1 M db records of 150 bp (affine NW).  Prints one JSON object."""
import json, os, random, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLI = os.path.join(ROOT, "sequencealigning_b200", "_lib", "sa_align")


def make(path_q, path_d, n, seed=3):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 4, 150)
    letters = np.frombuffer(b"ACGT", np.uint8)
    open(path_q, "wb").write(b">q\n" + letters[q].tobytes() + b"\n")
    db = np.tile(q, (n, 1))
    mut = rng.random((n, 150)) < 0.05
    db[mut] = rng.integers(0, 4, int(mut.sum()))
    with open(path_d, "wb") as f:
        chunk = 100000
        for i0 in range(0, n, chunk):
            rows = letters[db[i0:i0 + chunk]]
            f.write(b"".join(b">d%d\n%s\n" % (i0 + k, rows[k].tobytes()) for k in range(rows.shape[0])))


def run(args):
    t0 = time.perf_counter()
    r = subprocess.run([CLI] + args, stdout=subprocess.DEVNULL if "--no-output" in args else subprocess.PIPE, stderr=subprocess.PIPE)
    dt = time.perf_counter() - t0
    timing = [l for l in r.stderr.decode("latin1").splitlines() if l.startswith("timing:")]
    return dt, (timing[-1] if timing else ""), (len(r.stdout) if r.stdout else 0)


def oracle_job(path_q, path_d, algo, sample):
    from oracle import binding as ob
    ob.build()
    t0 = time.perf_counter()
    q = ob.parse_fasta_bytes(open(path_q, "rb").read())
    d = ob.parse_fasta_bytes(open(path_d, "rb").read())
    t_parse = time.perf_counter() - t0
    from sequencealigning_b200 import PairBatch
    pairs = [(q.seqs[0], s) for s in d.seqs[:sample]]
    b = PairBatch.from_pairs(pairs)
    t0 = time.perf_counter()
    if algo == "linear":
        ob.linear_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=320, n_threads=1)
    else:
        ob.affine_batch(b.residues, b.q_off, b.q_len, b.d_off, b.d_len, cigar_stride=320, n_threads=1)
    t_align = time.perf_counter() - t0
    return {"parse_seconds_whole_file": t_parse, "align_seconds_sample": t_align, "sample_pairs": len(pairs),
            "align_seconds_per_pair": t_align / max(len(pairs), 1), "cores": 1}


def main():
    out = {}
    tmp = "/tmp"
    for name, n, algo, cli_algo in (("config1_1x1000_linear", 1000, "linear", "needleman-wunsch-linear"),
                                    ("1x1M_affine", 1_000_000, "affine", "needleman-wunsch")):
        pq, pd = f"{tmp}/q_{n}.fa", f"{tmp}/d_{n}.fa"
        make(pq, pd, n)
        base = ["-q", pq, "-d", pd, "-a", cli_algo, "--timing"]
        run(base + ["--no-output"])  # warm the page cache and the driver
        wall_q, timing_q, _ = run(base + ["--no-output"])
        wall_p, timing_p, nbytes = run(base)
        o = oracle_job(pq, pd, algo, min(n, 20000))
        out[name] = {"pairs": n, "fasta_bytes": os.path.getsize(pd) + os.path.getsize(pq),
                     "cli_wall_seconds_no_output": wall_q, "cli_timing_no_output": timing_q,
                     "cli_wall_seconds_with_text": wall_p, "cli_timing_with_text": timing_p, "stdout_bytes": nbytes,
                     "cpu_oracle": o, "cpu_oracle_projected_seconds": o["parse_seconds_whole_file"] + o["align_seconds_per_pair"] * n}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
