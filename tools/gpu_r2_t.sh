mkdir -p gpurun_out
for pf in 6 262 264 268 6; do SA_WALK_PF=$pf python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pf $pf value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'probe', round(d['roofline']['launch']['gcups'],1), 'step frac', round(d['roofline']['whole_step_frac'],4))"; done
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 300000"
for pf in 262 268; do
SA_WALK_PF=$pf ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file gpurun_out/launches_pf$pf.csv $CMD > gpurun_out/ncu_pf.log 2>&1
done
