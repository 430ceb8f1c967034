bash tools/run_gpu_tests.sh > gpurun_out/run_all.log 2>&1; cat gpurun_out/summary.txt
timeout 600 python tools/kmeans_c_timing.py 4194304 262144 2 > gpurun_out/kmeans_c.log 2>&1; tail -1 gpurun_out/kmeans_c.log | cut -c1-900
python __graft_entry__.py smoke 2>&1 | tail -1
