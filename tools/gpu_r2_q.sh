for mb in 1 16; do SA_FILL_MINB=$mb python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('minb $mb value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1), 'frac', round(d['roofline']['frac'],4))"; done
