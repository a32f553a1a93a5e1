#!/bin/bash
# Pins the parity claim on the REAL reference: given a cargo-built binary of
# Qw11111111111/SequenceAligning (`cargo build --release` -> target/release/a_star_align; there is no Rust
# toolchain in the build image, so this runs wherever one exists), runs it and `sa_align` on the same
# FASTA files, one pair per process (a reference panic then ends only that pair's process, as it would in
# a user's run), and diffs stdout.
#
#   tools/diff_vs_reference.sh /path/to/a_star_align [corpus.json ...]
#
# Corpora: tests/golden/affine_golden.json (-a needleman-wunsch, compared with `sa_align --all`, modulo the
# reference's per-pair Duration line) and tests/golden/wfa_golden.json (-a wfa; pairs on which the
# reference never converges are skipped: its output is endless).  Exit status 0 = every pair identical.
set -u
REF=${1:?usage: $0 <reference binary> [golden.json ...]}
shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CLI="$ROOT/sequencealigning_b200/_lib/sa_align"
[ -x "$CLI" ] || { echo "build sa_align first: python -m sequencealigning_b200.build" >&2; exit 2; }
CORPORA=("$@")
[ ${#CORPORA[@]} -gt 0 ] || CORPORA=("$ROOT/tests/golden/affine_golden.json" "$ROOT/tests/golden/wfa_golden.json")
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
fail=0; total=0
strip_duration() { sed -E '/^[0-9.]+(ns|µs|ms|s)$/d'; }
for corpus in "${CORPORA[@]}"; do
  case "$corpus" in *wfa*) algo=wfa; extra=() ;; *) algo=needleman-wunsch; extra=(--all) ;; esac
  python3 - "$corpus" "$TMP" <<'PY'
import json, sys
vec = json.load(open(sys.argv[1]))["vectors"]
with open(sys.argv[2] + "/pairs.tsv", "w") as f:
    for v in vec:
        if v.get("status") in ("NO_CONVERGENCE", 2):
            continue
        if v["seq1"] and v["seq2"]:
            f.write(v["seq1"] + "\t" + v["seq2"] + "\n")
PY
  while IFS=$'\t' read -r s1 s2; do
    printf '>q\n%s\n' "$s1" > "$TMP/q.fa"; printf '>d\n%s\n' "$s2" > "$TMP/d.fa"
    timeout 20 "$REF" -q "$TMP/q.fa" -d "$TMP/d.fa" -a "$algo" 2> "$TMP/ref.err" | strip_duration > "$TMP/ref.out"
    "$CLI" -q "$TMP/q.fa" -d "$TMP/d.fa" -a "$algo" "${extra[@]}" 2> "$TMP/our.err" | strip_duration > "$TMP/our.out"
    total=$((total + 1))
    if ! cmp -s "$TMP/ref.out" "$TMP/our.out"; then
      fail=$((fail + 1))
      echo "DIFF ($algo) seq1=$s1 seq2=$s2"; diff "$TMP/ref.out" "$TMP/our.out" | head -10
    fi
  done < "$TMP/pairs.tsv"
done
echo "$total pairs compared, $fail differ"
[ "$fail" -eq 0 ]
