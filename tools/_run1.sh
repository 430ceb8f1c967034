bash tools/run_gpu_tests.sh > gpurun_out/run_all.log 2>&1; cat gpurun_out/summary.txt
TM_MOTION_SCALAR=1 timeout 600 python -m pytest tests/test_gpu_core.py -m gpu -q -x -k "motion or reconstruct" 2>&1 | tail -1
python __graft_entry__.py smoke 2>&1 | tail -1
