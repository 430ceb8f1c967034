bash tools/run_gpu_tests.sh encode > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
python tools/encode_clip.py > gpurun_out/encode_720p.log 2>&1; tail -1 gpurun_out/encode_720p.log | cut -c1-600
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench1.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['encode'])"
