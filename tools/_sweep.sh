for dbg in 0 1; do
echo "== dbg=$dbg"; TM_LIB_PATH=$PWD/gpurun_variants/libtm_T.so TM_TK_DBG=$dbg python tools/knn_probe.py 65536 75776 64 2>&1 | sort | uniq | grep -E "mma|warp 0 |warp 5 " | awk 'NR%4==1' | head -6
done > gpurun_out/timing.log 2>&1
cat gpurun_out/timing.log
