timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E|FAILED|passed|failed" | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
