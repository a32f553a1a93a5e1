# round 2: taint queue in the fill, refill beside the walk
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'probe', round(d['roofline']['launch']['gcups'],1), 'step frac', round(d['roofline']['whole_step_frac'],4), 'rerun', d['pairs_rerun_per_step'])"; done
python bench.py --workload config3 --steps 10 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c3 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'probe', round(d['roofline']['launch']['gcups'],1), 'step frac', round(d['roofline']['whole_step_frac'],4))"
