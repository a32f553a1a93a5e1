# quick check after a host-side change: all GPU tests, then the headline with and without the threaded head scan
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for t in 1 4 4 1; do SA_SCAN_THREADS=$t python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('scan threads $t value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'bytes', round(d['e2e']['byte_per_residue']['value'],1), 'packer', round(d['e2e']['packer_included']['value'],1))"; done
