run() { env "$@" python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --configs none $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$* $EXTRA value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'probe', round(d['roofline']['launch']['gcups'],1), 'frac', round(d['roofline']['frac'],4), 'step', round(d['roofline']['whole_step_frac'],4))"; }
EXTRA=""
run SA_FILL_MINB=16 SA_ORMASK=0xFF
run SA_FILL_MINB=16 SA_ORMASK=0x01
run SA_FILL_MINB=16 SA_ORMASK=0x1201
run SA_FILL_MINB=16 SA_ORMASK=0x11201
run SA_FILL_MINB=1
for L in 100 200 300; do EXTRA="--length $L --pairs 600000"; run SA_FILL_MINB=1; run SA_FILL_MINB=16; done
SA_FILL_MINB=16 timeout 600 python -m pytest tests/test_gpu_affine.py -m gpu -x -q -k "random_ragged or config_shapes or every_kernel_form or two_bit" 2>&1 | tail -2
