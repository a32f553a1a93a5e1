# the evidence run (gpurun -- bash tools/gpu_evidence.sh) -- gpu tests, smoke, the default bench line, the reference arm, launch list, ncu captures
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 1500 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/bench_r02_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2>> gpurun_out/bench_r02_n1.err
python bench.py --workload config3 --steps 10 --warmup 3 --configs none > gpurun_out/bench_r02_c3.json 2>> gpurun_out/bench_r02_n1.err
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_l.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 200000"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_affine_fill -s 0 -c 1 -o gpurun_out/fill_r02_final $CMD > gpurun_out/ncu_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nw_affine_walk -s 0 -c 1 -o gpurun_out/walk_r02_final $CMD > gpurun_out/ncu_walk.log 2>&1
CMD="python bench.py --workload config3 --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 100000"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_affine_fill -s 0 -c 1 -o gpurun_out/fill250_r02_final $CMD > gpurun_out/ncu_fill250.log 2>&1
timeout 600 python tools/local_bench.py 2>&1 | tail -12 > gpurun_out/local_bench_r02.txt
ls -la gpurun_out | tail -8
