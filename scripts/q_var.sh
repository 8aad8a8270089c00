# A/B of library variants: bash scripts/q_var.sh robot lib1 lib2 ...   (device-only headline bench, 10 steps)
R=$1; shift
for L in "$@"; do echo -n "$L: "; PNC_B200_LIB=$(realpath $L) timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'])"; done
