bash tools/run_gpu_tests.sh knn > gpurun_out/run1.log 2>&1; tail -2 gpurun_out/run1.log
for cfg in "16 0" "16 1"; do
  set -- $cfg
  echo "slack=$1 dbg=$2: $(TM_TK_SLACK=$1 TM_TK_DBG=$2 python tools/knn_timing.py 65536 432000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['k64']['kernel_ms'], d['k1']['kernel_ms'])")"
done > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
for dbg in 0 1; do
echo "== dbg=$dbg"; TM_LIB_PATH=$PWD/gpurun_variants/libtm_T.so TM_TK_DBG=$dbg python tools/knn_probe.py 65536 75776 64 2>&1 | sort | uniq | awk 'NR%4==1' | head -12
done > gpurun_out/timing.log 2>&1
cat gpurun_out/timing.log
