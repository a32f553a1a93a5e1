mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 1500 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err ) 2>&1 | tail -3
python bench.py --workload config3 --steps 10 --warmup 3 --configs none > gpurun_out/bench_r02_c3.json 2>> gpurun_out/bench_r02_n1.err
tail -c 300 gpurun_out/bench_r02_n1.err
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_l.log 2>&1
