# A/B of k_update_system builds: bash scripts/q_dyn.sh tag lib...   (parity of the default build first)
T=${1:-qdyn}; shift; O=gpurun_out/$T; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_first_principles.py -x -q -m gpu -k "update_system or centroidal or first_principles or manipulator" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for R in atlas laikago valkyrie; do
echo -n "$R default: "; timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'], d.get('fp32_mode',{}).get('kernels_ms'))"
for L in "$@"; do echo -n "$R $L: "; PNC_B200_LIB=$(realpath $L) timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'], d.get('fp32_mode',{}).get('kernels_ms'))"; done; done
