CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgrad_td_kernel -s 19 -c 1 -o gpurun_out/prof12_dgtd -f $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | grep prof12
