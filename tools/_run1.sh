bash tools/run_gpu_tests.sh knn > gpurun_out/run1.log 2>&1; tail -1 gpurun_out/summary.txt
python tools/knn_timing.py 65536 432000 2>&1 | tail -1 | cut -c1-400
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_i8_topk -s 1 -c 1 python tools/knn_probe.py 65536 432000 64 2>&1 | grep -E "dram__|gpu__time" 
