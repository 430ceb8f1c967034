bash tools/run_gpu_tests.sh "sliding or motion or reconstruct" > gpurun_out/run1.log 2>&1; tail -3 gpurun_out/run1.log; tail -30 "gpurun_out/test_sliding or motion or reconstruct.log"
