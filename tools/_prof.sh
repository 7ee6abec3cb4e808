CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgrad_dense_kernel -s 165 -c 3 -o gpurun_out/prof13_dgrad -f $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | grep prof13
