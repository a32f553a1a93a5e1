for sortm in 0 1; do SA_SORT=$sortm python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sort', $sortm, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1))"; done
SA_SORT=1 python bench.py --workload config3 --steps 10 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c3 sort1 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'probe', round(d['roofline']['launch']['gcups'],1))"
python bench.py --workload config3 --steps 10 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c3 sort0 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'probe', round(d['roofline']['launch']['gcups'],1))"
