mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1), 'frac', round(d['roofline']['frac'],4), 'step frac', round(d['roofline']['whole_step_frac'],4))"; done
python tools/local_bench.py 2>&1 | grep "global linear"
