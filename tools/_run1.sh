for v in np2 np3; do echo "$v: $(TM_LIB_PATH=$PWD/gpurun_variants/libtm_$v.so python tools/feat_timing.py 2>&1 | tail -1)"; done
echo "np4: $(python tools/feat_timing.py 2>&1 | tail -1)"
bash tools/run_gpu_tests.sh features sliding > gpurun_out/run1.log 2>&1; cat gpurun_out/summary.txt
