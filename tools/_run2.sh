for v in 3; do
RL_VARIANT=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_step']; print('variant $v', d['value'], c)"
done
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q 2>&1 | tail -2
