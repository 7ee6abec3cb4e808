timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
