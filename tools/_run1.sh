bash tools/run_gpu_tests.sh knn > gpurun_out/run1.log 2>&1; tail -2 gpurun_out/run1.log
python tools/knn_timing.py 65536 432000 2>&1 | tail -1 > gpurun_out/knn_timing.log; cat gpurun_out/knn_timing.log
TM_TK_DBG=1 python tools/knn_timing.py 65536 432000 2>&1 | tail -1
