bash tools/run_gpu_tests.sh knn matcher kmeans > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
python tools/knn_timing.py 65536 432000 2>&1 | tail -1 | cut -c1-420
TM_TK_DBG=16 python tools/knn_timing.py 65536 432000 2>&1 | tail -1 | cut -c1-420
