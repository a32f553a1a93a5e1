SA_FILL_CHAIN=4 SA_FILL_CHAIN_MIN=0 timeout 900 python -m pytest tests/test_gpu_affine.py -m gpu -x -q -k "not long and not tiled and not sentinel and not cooptimal and not checkpoint" 2>&1 | tail -5
for ch in 1 2 4 8; do SA_FILL_CHAIN=$ch python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chain $ch value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1), 'frac', round(d['roofline']['frac'],4))"; done
