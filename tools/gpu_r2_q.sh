for m in 0x11201 0x1201 0x01 0x11201; do SA_ORMASK=$m python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ormask $m value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1), 'frac', round(d['roofline']['frac'],4))"; done
