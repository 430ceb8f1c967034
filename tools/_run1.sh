timeout 300 bash tools/run_gpu_tests.sh motion reconstruct encode > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt; tail -15 gpurun_out/test_motion.log | head -30
timeout 300 python tools/encode_clip.py > gpurun_out/encode_720p.log 2>&1; tail -1 gpurun_out/encode_720p.log | cut -c1-1100
