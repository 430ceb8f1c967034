timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_i8_topk -s 2 -c 1 -o gpurun_out/prof_topk_bench_r01c -f python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log
