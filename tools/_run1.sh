bash tools/run_gpu_tests.sh features sliding > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
echo "intcvt1: $(python tools/feat_timing.py 2>&1 | tail -1)"
for v in ic0 ic2; do echo "$v: $(TM_LIB_PATH=$PWD/gpurun_variants/libtm_$v.so python tools/feat_timing.py 2>&1 | tail -1)"; done
TM_LIB_PATH=$PWD/gpurun_variants/libtm_ic2.so python -m pytest tests/test_gpu_core.py -m gpu -q -x -k "features or sliding" 2>&1 | tail -1
