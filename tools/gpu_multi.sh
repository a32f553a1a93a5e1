# N-GPU check (run as: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): the multi-device tests on all N devices, then the
# bench under torchrun with the sharded block (configs[2]: ONE 10 M x 250 bp list through ONE multi-device call)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) 2>&1 | tail -4
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu --configs sharded > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json').read().splitlines() if l.startswith('{')][-1])
print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'e2e bytes', d['e2e']['byte_per_residue']['value'])
s=d.get('sharded',{})
print('sharded', {k:s.get(k) for k in ('value','ms_per_step','cell_imbalance_max_over_mean','error','slowest_device_kernels_ms')}, 'byte', (s.get('byte_per_residue') or {}).get('value'), 'packer', (s.get('packer_included') or {}).get('value'))
for sh in s.get('shards',[]): print('  ', sh)
PY
