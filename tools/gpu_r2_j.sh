mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_affine.py -m gpu -x -q -k "full_size or config_shapes" 2>&1 | tail -3
timeout 300 python tools/trace_e2e.py 2> gpurun_out/trace_e2e.err; grep -v "count of segment" gpurun_out/trace_e2e.err | tail -40
for seg in 131072 65536; do SA_SEG_PAIRS=$seg python bench.py --steps 20 --warmup 3 --skip-cpu --configs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('seg', $seg, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'bytes', round(d['e2e']['byte_per_residue']['value'],1))"; done
