#!/bin/bash
# developer check: the sa_align CLI end to end on a generated FASTA pair (1 query x N db records)
N=${1:-100000}
python - <<PY
import random
random.seed(3)
q = "".join(random.choice("ACGT") for _ in range(150))
open("/tmp/q.fa", "w").write(">q\n" + q + "\n")
with open("/tmp/d.fa", "w") as f:
    for i in range($N):
        s = list(q)
        for k in range(8):
            s[random.randrange(150)] = random.choice("ACGT")
        f.write(">d%d\n%s\n" % (i, "".join(s)))
PY
T0=$(date +%s.%N); sequencealigning_b200/_lib/sa_align -q /tmp/q.fa -d /tmp/d.fa -a needleman-wunsch -m global > /tmp/out.txt 2> /tmp/err.txt
echo "wall $(echo "$(date +%s.%N) - $T0" | bc) s"; tail -2 /tmp/err.txt; wc -c /tmp/out.txt; head -8 /tmp/out.txt
