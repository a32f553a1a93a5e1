# round 2, call B: gpu tests, the full default bench line, launch list, ncu captures of the long and WFA kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
( time timeout 1200 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/bench_r02_n1.err
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_l.log 2>&1
CMD="python bench.py --workload config5nw --pairs 64 --length 30000 --steps 1 --warmup 1 --skip-cpu --skip-e2e --configs none"
$CMD > gpurun_out/plain_long.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_long_fwd -s 40 -c 1 -o gpurun_out/long_fwd_r02 $CMD > gpurun_out/ncu_long2.log 2>&1
CMD="python bench.py --workload config4 --pairs 20000 --steps 1 --warmup 1 --skip-cpu --skip-e2e --configs none"
$CMD > gpurun_out/plain_wfa.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wfa_standard -s 1 -c 1 -o gpurun_out/wfa_std_r02 $CMD > gpurun_out/ncu_wfa.log 2>&1
ls -la gpurun_out | tail -12
