CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:dense_fwd -s 108 -c 1 -o gpurun_out/prof4_dense_fwd -f $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/ | grep prof4
