mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
B="python bench.py --steps 10 --warmup 3 --skip-cpu --configs none"
$B > gpurun_out/d_c2.json 2> gpurun_out/d_c2.err; python -c "import json;d=json.load(open('gpurun_out/d_c2.json'));print('c2 value',d['value'],'e2e',d['e2e']['value'],'probe',d['roofline']['launch'],'frac',d['roofline']['frac'])"
SA_FILL_MINB=16 $B > gpurun_out/d_c2_minb16.json 2>> gpurun_out/d_c2.err; python -c "import json;d=json.load(open('gpurun_out/d_c2_minb16.json'));print('c2 minb16 value',d['value'],'e2e',d['e2e']['value'],'probe',d['roofline']['launch'],'frac',d['roofline']['frac'])"
$B --workload config3 > gpurun_out/d_c3.json 2>> gpurun_out/d_c2.err; python -c "import json;d=json.load(open('gpurun_out/d_c3.json'));print('c3 value',d['value'],'e2e',d['e2e']['value'],'probe',d['roofline']['launch'],'frac',d['roofline']['frac'])"
python bench.py --workload config4 --steps 2 --warmup 1 --skip-cpu --configs none > gpurun_out/d_c4.json 2>> gpurun_out/d_c2.err; python -c "import json;d=json.load(open('gpurun_out/d_c4.json'));print('c4 value',d['value'],'ms',d['ms_per_step'],'aln/s',d['alignments_per_s'])"
timeout 900 python tools/cli_bench.py > gpurun_out/cli_bench_r02.json 2> gpurun_out/cli_bench.err; tail -c 1500 gpurun_out/cli_bench_r02.json; tail -3 gpurun_out/cli_bench.err
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 200000"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_affine_fill -s 2 -c 1 -o gpurun_out/fill_r02 $CMD > gpurun_out/ncu_fill.log 2>&1
ls -la gpurun_out | tail -5
