mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15
timeout 600 python tools/local_bench.py 2>&1 | tail -20
