bash tools/run_gpu_tests.sh knn > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
python tools/knn_timing.py 65536 432000 adv 2>&1 | tail -1 | cut -c1-420
python tools/knn_timing.py 65536 432000 2>&1 | tail -1 | cut -c1-420
