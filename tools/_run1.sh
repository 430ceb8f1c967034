bash tools/run_gpu_tests.sh knn matcher dropin kmeans > gpurun_out/run1.log 2>&1
python tools/knn_timing.py 65536 432000 > gpurun_out/knn_timing.log 2>&1
python bench.py --steps 6 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/run1.log; cat gpurun_out/knn_timing.log; tail -c 1500 gpurun_out/bench.log
