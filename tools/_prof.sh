CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 198 -c 6 -o gpurun_out/prof10_wgrad -f $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | grep prof10
