# A/B of library builds: bash scripts/q_ab.sh tag "robots" lib...   (parity subset of the default build first)
T=${1:-qab}; shift; RS=$1; shift; O=gpurun_out/$T; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tick_matches_oracle or golden_vectors or large_sample or transmission or generic or tick_host_equals" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for R in $RS; do
echo -n "$R default: "; timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'])"
for L in "$@"; do echo -n "$R $L: "; PNC_B200_LIB=$(realpath $L) timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'])"; done; done
