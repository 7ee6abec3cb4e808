timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_modules_gpu.py tests/test_shapes_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_step']; print(d['value'], d['ms_per_step'], c['wgrad'])"
