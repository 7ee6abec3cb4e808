#!/bin/bash
# One parameterised GPU-box script (replaces the per-experiment scratch scripts of round 1):
#   tools/gpu_check.sh [tests|ops|bench|launches|all] ...
# Everything it writes goes to gpurun_out/.
set -u
mkdir -p gpurun_out
for what in "$@"; do
  case "$what" in
    ops)      python -m pytest tests/test_ops_gpu.py -x -q -s 2>&1 | tail -150 > gpurun_out/ops.log; tail -40 gpurun_out/ops.log ;;
    opsall)   python -m pytest tests/test_ops_gpu.py -q -rP 2>&1 > gpurun_out/ops.log; grep -E "^(PASSED|FAILED|ERROR)|rel err|within|passed|failed|differ" gpurun_out/ops.log | tail -220 ;;
    tests)    python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/tests.log; tail -30 gpurun_out/tests.log ;;
    testsall) python -m pytest tests -m gpu -q 2>&1 | tail -120 > gpurun_out/tests.log; tail -60 gpurun_out/tests.log ;;
    smoke)    python __graft_entry__.py smoke 2>&1 | tail -5 ;;
    bench)    python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    benchq)   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    infer)    python bench.py --workload infer --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; cat gpurun_out/bench_infer.json; tail -3 gpurun_out/bench_infer.err ;;
    mme)      python bench.py --workload mme --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mme.json 2> gpurun_out/bench_mme.err; cat gpurun_out/bench_mme.json; tail -3 gpurun_out/bench_mme.err ;;
    nograph)  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; cat gpurun_out/bench_nograph.json; tail -3 gpurun_out/bench_nograph.err ;;
    bar)      for a in "train" "train --graph" "infer"; do python tools/stock_torch_bar.py $a 2>&1 | tail -1; done | tee gpurun_out/stock_torch_bar.jsonl ;;
    graphtests) python -m pytest tests/test_graph_ddp_gpu.py -q -x -rP 2>&1 | tail -40 ;;
    perlaunch) python tools/per_launch.py > gpurun_out/per_launch.txt 2>&1; cat gpurun_out/per_launch.txt ;;
    perlaunch0) BWD_FUSED=0 python tools/per_launch.py > gpurun_out/per_launch_unfused.txt 2>&1; cat gpurun_out/per_launch_unfused.txt ;;
    launches) # kernel list of one eager step (the graph replays the same kernels): skip the 3 warm-up steps, capture > 1 step
              python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
              ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-730} -c 300 --csv --log-file gpurun_out/launches.csv \
                  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log ;;
    ncu:*)    k="${what#ncu:}"   # one launch of kernel regex k (first occurrence after 3 warm-up steps) with the full metric set
              python tools/one_step.py > gpurun_out/plain_one_step.log 2>&1 &&
              ncu --set full --clock-control none --import-source on -k "regex:$k" -s ${NCU_SKIP:-0} -c 1 -f -o "gpurun_out/prof_$k" \
                  python tools/one_step.py > "gpurun_out/ncu_$k.log" 2>&1; tail -3 "gpurun_out/ncu_$k.log" ;;
    refarm)   python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json ;;
    final)    # everything the round-end driver runs, plus the secondary workloads; JSON lines land in gpurun_out/
              python -m pytest tests -m gpu -q 2>&1 | tail -5
              python __graft_entry__.py smoke 2>&1 | tail -1
              python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
              python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; cut -c1-300 gpurun_out/bench_infer.json
              python bench.py --workload mme --no-cpu-baseline > gpurun_out/bench_mme.json 2> gpurun_out/bench_mme.err; cut -c1-300 gpurun_out/bench_mme.json
              python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; cut -c1-300 gpurun_out/bench_nograph.json
              python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json ;;
    *) echo "unknown step $what" ;;
  esac
done
