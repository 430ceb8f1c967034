bash tools/run_gpu_tests.sh knn kmeans matcher dropin > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
python tools/knn_timing.py 65536 432000 2>&1 | tail -1 > gpurun_out/knn_timing.log; cat gpurun_out/knn_timing.log
