mkdir -p gpurun_out
for head in 65536 32768 16384; do SA_SEG_HEAD=$head python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('head', $head, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'bytes', round(d['e2e']['byte_per_residue']['value'],1), 'packer', round(d['e2e']['packer_included']['value'],1))"; done
