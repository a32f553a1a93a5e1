mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_wfa.py tests/test_cli.py -m gpu -x -q ) 2>&1 | tail -5
python bench.py --workload config4 --steps 2 --warmup 1 --skip-cpu --configs none > gpurun_out/h_c4.json 2> gpurun_out/h.err; python -c "import json;d=json.load(open('gpurun_out/h_c4.json'));print('c4 value',d['value'],'ms',d['ms_per_step'],'aln/s',d['alignments_per_s'])"
python bench.py --workload config5 --steps 2 --warmup 1 --skip-cpu --configs none > gpurun_out/h_c5.json 2>> gpurun_out/h.err; python -c "import json;d=json.load(open('gpurun_out/h_c5.json'));print('c5 value',d['value'],'ms',d['ms_per_step'],'aln/s',d['alignments_per_s'])"
tail -3 gpurun_out/h.err
