CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 200 -c 1 -o gpurun_out/prof9_wgrad -f $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | grep prof9
