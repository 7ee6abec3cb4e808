timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_modules_gpu.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['class_ms_per_step'])"
