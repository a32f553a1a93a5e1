mkdir -p gpurun_out
timeout 600 python tools/sweep_kernel.py 524288 150 8 0x00,0x01,0x05,0x11 2>&1 | tail -6
SA_FILL_MINB=16 timeout 300 python tools/sweep_kernel.py 524288 150 8 0x00 2>&1 | tail -1
timeout 300 python tools/sweep_kernel.py 524288 250 16 0x00 2>&1 | tail -1
bash tools/bench_quick.sh 524288 262144 131072 2>&1 | tail -3
python bench.py --workload config4 --steps 2 --warmup 1 --skip-cpu --configs none > gpurun_out/e_c4.json 2> gpurun_out/e.err; python -c "import json;d=json.load(open('gpurun_out/e_c4.json'));print('c4 value',d['value'],'ms',d['ms_per_step'],'aln/s',d['alignments_per_s'])"
timeout 900 python tools/cli_bench.py > gpurun_out/cli_bench_r02.json 2> gpurun_out/cli_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/cli_bench_r02.json'))
for k,v in d.items():
    print(k, 'wall no-output', round(v['cli_wall_seconds_no_output'],3), 'with text', round(v['cli_wall_seconds_with_text'],3), 'cpu projected', round(v['cpu_oracle_projected_seconds'],2))
    print('  ', v['cli_timing_no_output']); print('  ', v['cli_timing_with_text'])
PY
tail -3 gpurun_out/cli_bench.err
python tests/../sequencealigning_b200/_lib/sa_align -V
( cd /tmp && /root/repo/sequencealigning_b200/_lib/sa_align -q /tmp/q_1000000.fa -d /tmp/d_1000000.fa -a needleman-wunsch --timing --no-output --pageable 2>&1 | tail -1 )
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 200000"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_affine_fill -s 0 -c 1 -o gpurun_out/fill_r02 $CMD > gpurun_out/ncu_fill.log 2>&1
ls -la gpurun_out | tail -4
