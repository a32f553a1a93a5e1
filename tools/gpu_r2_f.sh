mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_linear.py -m gpu -x -q ) 2>&1 | tail -6
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/bench_r02_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_n2.json'))
print('N=2 value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
s=d.get('sharded',{})
print('sharded', {k:s.get(k) for k in ('value','ms_per_step','cell_imbalance_max_over_mean','error')}, 'byte', (s.get('byte_per_residue') or {}).get('value'), 'packer', (s.get('packer_included') or {}).get('value'))
for sh in s.get('shards',[]): print('  ', sh)
PY
