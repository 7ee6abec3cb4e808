CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 300 gpurun_out/plain.log
ncu --set full --clock-control none --import-source on -k regex:dense_fwd -s 107 -c 2 -o gpurun_out/prof2_dense_fwd -f $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgrad_dense -s 56 -c 2 -o gpurun_out/prof2_dgrad_dense -f $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad -s 69 -c 2 -o gpurun_out/prof2_wgrad -f $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/ | grep prof2
