"""The command line with the reference's flags (reference: src/parse.rs:8-34, src/main.rs:19-80)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "sequencealigning_b200", "_lib", "sa_align")


@pytest.fixture(scope="module")
def cli():
    from sequencealigning_b200.build import build_all
    build_all()
    assert os.path.exists(CLI)
    return CLI


def _fa(path, recs):
    with open(path, "wb") as f:
        for name, seq in recs:
            f.write(b">" + name + b"\n" + seq + b"\n")
    return str(path)


def test_flags_and_fasta_errors(cli, tmp_path):
    assert subprocess.run([cli, "-V"], capture_output=True, text=True).stdout.startswith("sa_align 0.1.0")
    assert subprocess.run([cli, "-h"], capture_output=True, text=True).returncode == 0
    assert subprocess.run([cli], capture_output=True).returncode == 2
    q = _fa(tmp_path / "q.fa", [(b"q1", b"ACGT")])
    bad = _fa(tmp_path / "db.txt", [(b"d1", b"ACGT")])
    r = subprocess.run([cli, "-q", q, "-d", bad], capture_output=True, text=True)
    assert r.returncode == 0 and "DB fasta could not be opened" in r.stderr and "aborting" in r.stderr  # main.rs:24-28
    r = subprocess.run([cli, "-q", q, "-d", q, "-m", "nonsense"], capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([cli, "-q", q, "-d", q, "-a", "a-star"], capture_output=True, text=True)
    assert r.returncode == 2 and "a-star" in r.stderr


@pytest.mark.gpu
def test_affine_stdout_matches_reference_text(cli, tmp_path, oracle):
    query = [(b"q1", b"ACGTACGT"), (b"q2", b"AAAA"), (b"q3", b"GACGT")]
    db = [(b"d1", b"ACGGT"), (b"d2", b"AAA"), (b"d3", b"ACGT")]
    q, d = _fa(tmp_path / "q.fasta", query), _fa(tmp_path / "d.fna", db)
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "needleman-wunsch", "-m", "global"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # strip the per-pair Duration line (non-deterministic in the reference too, nw_affine:431)
    out = re.sub(r"^[0-9.]+(ns|µs|ms|s)\n", "", r.stdout, flags=re.M)
    exp = ""
    n_panic = 0
    for dn, ds in db:          # db-major, main.rs:61-62
        for qn, qs in query:
            text, n, pan = oracle.affine_print_all(qs, ds, max_alignments=1)
            exp += text
            n_panic += oracle.affine_align(qs, ds).status in (oracle.REF_PANIC, oracle.REF_PANIC_EARLY)
    assert out == exp
    assert r.stderr.count("the reference panics here") == n_panic and n_panic >= 1
    assert r.stdout.count("\n") - out.count("\n") == len(db) * len(query)  # one Duration line per pair
    # --all: every co-optimal alignment, i.e. the reference's complete stdout for the pair
    r4 = subprocess.run([cli, "-q", q, "-d", d, "--all"], capture_output=True, text=True)
    out4 = re.sub(r"^[0-9.]+(ns|µs|ms|s)\n", "", r4.stdout, flags=re.M)
    exp4 = "".join(oracle.affine_print_all(qs, ds)[0] for dn, ds in db for qn, qs in query)
    assert out4 == exp4 and out4.count("alignment found") > out.count("alignment found")
    # --strict: stop with exit status 101 at the first pair the reference dies on
    r2 = subprocess.run([cli, "-q", q, "-d", d, "--strict"], capture_output=True, text=True)
    assert r2.returncode == 101
    # non-global modes: nw_affine:433-434 via main.rs:68-74
    r3 = subprocess.run([cli, "-q", q, "-d", d, "-m", "local"], capture_output=True, text=True)
    assert r3.stdout == "" and r3.stderr.count("Error in alignment: not implemented") == 9
    assert "An error occured during alignment of >q1 and >d1" in r3.stderr


@pytest.mark.gpu
def test_wfa_and_linear_algos(cli, tmp_path):
    q = _fa(tmp_path / "q.fa", [(b"q", b"AAAATTTTCCCC"), (b"r", b"ACGT")])
    d = _fa(tmp_path / "d.fa", [(b"d", b"AAAATCTCC"), (b"e", b"ACGT")])
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "wfa"], capture_output=True, text=True)
    assert "converged with score 25: \n" in r.stdout          # wfa.rs:36 on the reference's own test pair
    assert "never converges" in r.stderr                      # ACGT vs ACGT (wfa.rs:189)
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "wfa-standard"], capture_output=True, text=True)
    assert r.stdout.count("gap-affine cost") == 4 and ">r vs >e: gap-affine cost 0\n" in r.stdout
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "needleman-wunsch-linear"], capture_output=True, text=True)
    assert r.stdout.count("Alignment between sequences") == 4 and "seq1: ACGT\n      ||||\nseq2: ACGT" in r.stdout


@pytest.mark.gpu
def test_linear_stdout_matches_reference_text(cli, tmp_path, oracle):
    """needleman_wunsch.rs:193-201 header + the first hit (:205-213, Display :155-178), global and local mode."""
    query = [(b"q1", b"ACGTACGT"), (b"q2", b"AAAA"), (b"q3", b"TTGACGTAA")]
    db = [(b"d1", b"ACGGT"), (b"d2", b"CCC"), (b"d3", b"GGACGTCC")]
    q, d = _fa(tmp_path / "q.fasta", query), _fa(tmp_path / "d.fna", db)
    for mode, local in (("global", False), ("local", True)):
        r = subprocess.run([cli, "-q", q, "-d", d, "-a", "needleman-wunsch-linear", "-m", mode], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exp = ""
        for dn, ds in db:          # db-major, main.rs:61-62
            for qn, qs in query:
                exp += f"Alignment between sequences >{qn.decode()} and >{dn.decode()} found\n"
                exp += oracle.linear_print_hits(qs, ds, local, 1)[0]
        assert r.stdout == exp, mode
        # --all: every hit of every start cell = the reference's complete stdout (:106-116)
        r = subprocess.run([cli, "-q", q, "-d", d, "-a", "needleman-wunsch-linear", "-m", mode, "--all"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exp = ""
        for dn, ds in db:
            for qn, qs in query:
                exp += f"Alignment between sequences >{qn.decode()} and >{dn.decode()} found\n"
                exp += oracle.linear_print_hits(qs, ds, local)[0]
        assert r.stdout == exp, mode + " --all"
    # the linear aligner has no semi-global mode (n_w_align takes `local: bool`, :180)
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "needleman-wunsch-linear", "-m", "semi-global"], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist in the reference" in r.stderr


@pytest.mark.gpu
def test_wfa_stdout_matches_reference_text(cli, tmp_path, oracle):
    """`-a wfa`: wfa_align's complete stdout per pair, db-major, against the oracle's printer."""
    query = [(b"q1", b"ACGT"), (b"q2", b"AAAATTTTCCCC"), (b"q3", b"GATTACA")]
    db = [(b"d1", b"ACGA"), (b"d2", b"AAAATCTCC"), (b"d3", b"GCTTAGA"), (b"d4", b"ACGT")]
    q, d = _fa(tmp_path / "q.fasta", query), _fa(tmp_path / "d.fna", db)
    r = subprocess.run([cli, "-q", q, "-d", d, "-a", "wfa"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exp, bad = "", 0
    for dn, ds in db:
        for qn, qs in query:
            text, st = oracle.wfa_print(qs, ds)
            exp += text
            bad += st != 0
    assert r.stdout == exp
    assert r.stderr.count("the reference") == bad and "huhu, diag: 0\nElement {\n\tstate: M" in r.stdout
