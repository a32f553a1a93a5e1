mkdir -p gpurun_out
timeout 300 python tools/trace_e2e.py 2> gpurun_out/trace_e2e.err; grep -v "count of segment" gpurun_out/trace_e2e.err | grep -v "prepared\|queued A" | tail -22
