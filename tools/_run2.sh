for v in 3 7 11 19 35 63; do
RL_VARIANT=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_step']; print('variant $v', d['value'], 'dgrad_dense', c['dgrad_dense'])"
done
