bash tools/run_gpu_tests.sh motion reconstruct encode > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
python tools/encode_clip.py > gpurun_out/encode_720p.log 2>&1; tail -1 gpurun_out/encode_720p.log
