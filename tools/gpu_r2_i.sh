mkdir -p gpurun_out
timeout 600 python tools/trace_shard.py 1250000 0,0,0,0,0,0,0,0 2> gpurun_out/trace_shard8.err; grep -v "queued\|prepared\|count of segment\|main stream" gpurun_out/trace_shard8.err | tail -32
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_affine.py -m gpu -x -q 2>&1 | tail -3
