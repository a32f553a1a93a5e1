run() { env "$@" python bench.py --steps 20 --warmup 3 --skip-cpu --configs none $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$* $EXTRA value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'bytes', round(d['e2e']['byte_per_residue']['value'],1), 'packer', round(d['e2e']['packer_included']['value'],1))"; }
EXTRA=""
run SA_SEG_HEAD=65536
run SA_SEG_HEAD=32768
run SA_SEG_HEAD=16384
run SA_SEG_HEAD=49152
