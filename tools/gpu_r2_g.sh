mkdir -p gpurun_out
timeout 600 python tools/trace_shard.py 1250000 0 2> gpurun_out/trace_shard.err; grep -v "queued\|prepared\|count of segment" gpurun_out/trace_shard.err | tail -20; echo ...; grep "sa trace" gpurun_out/trace_shard.err | head -12; echo ...; grep "sa trace" gpurun_out/trace_shard.err | tail -14
