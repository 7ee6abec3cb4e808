timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 300 gpurun_out/bench_r01.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_v3.csv $CMD > gpurun_out/ncu1.log 2>&1
ls -la gpurun_out | grep -E "v3|bench_"
