mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --configs none --pairs 200000"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nw_affine_fill -s 0 -c 1 -o gpurun_out/fill_r02b $CMD > gpurun_out/ncu_fill.log 2>&1
ls -la gpurun_out/fill_r02b.ncu-rep
