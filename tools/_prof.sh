set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_fwd -s 105 -c 5 -o gpurun_out/prof_dense_fwd -f $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgrad_dense -s 55 -c 4 -o gpurun_out/prof_dgrad_dense -f $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad -s 68 -c 4 -o gpurun_out/prof_wgrad -f $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/
