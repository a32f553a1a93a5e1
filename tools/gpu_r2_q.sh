for c in 1 2; do SA_LONG_CELL=$c timeout 600 python -m pytest tests/test_gpu_affine.py -m gpu -x -q -k "tiled or pinned_at or sentinel" 2>&1 | tail -1; done
for c in 0 1 2; do SA_LONG_CELL=$c python bench.py --workload config5nw --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cell $c value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), (d.get('roofline') or {}).get('frac'))"; done
