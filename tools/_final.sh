set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 400 gpurun_out/bench_r01.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_r01.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches_v2.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dgrad_dense -s 56 -c 2 -o gpurun_out/prof3_dgrad_dense -f $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | grep -E "v2|prof3|bench_"
