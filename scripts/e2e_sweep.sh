# sweep of the pnc_wbc_tick_host chunk schedule (diagnostic): prints e2e solves/s, ms per step, device solves/s
# usage: bash scripts/e2e_sweep.sh "4 0.125" "3 0.15" ...   (chunks, fraction of the batch in the first chunk; 0 = equal)
for cfg in "$@"; do set -- $cfg; echo -n "chunks $1 first $2: "; PNC_TICK_CHUNKS=$1 PNC_TICK_FIRST=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), round(d['value']))"; done
