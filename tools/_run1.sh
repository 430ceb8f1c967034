bash tools/run_gpu_tests.sh tile_classes encode > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt; tail -30 gpurun_out/test_encode.log; tail -15 gpurun_out/test_tile_classes.log
