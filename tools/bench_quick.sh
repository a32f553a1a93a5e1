#!/bin/bash
# developer tool: bench.py value / e2e for a few SA_SEG_PAIRS settings
for seg in "$@"; do
  SA_SEG_PAIRS=$seg python bench.py --steps 10 --warmup 3 --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('seg', $seg, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])"
done
