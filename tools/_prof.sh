CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
for k in dense_fwd dgrad_dense wgrad; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 108 -c 1 -o gpurun_out/prof5_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/ | grep prof5
