for cfg in "3 2" "2 1" "2 2" "3 1" "4 2"; do
set -- $cfg
RL_DG_CTAS=$1 RL_WG_CTAS=$2 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dg $1 wg $2', d['value'], d['ms_per_step'])"
done
