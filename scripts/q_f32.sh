# fp32-mode check on the GPU box: fp64 parity subset (templated kernel), fp32 report, bench A/B.  Usage: $0 tag
T=${1:-qf32}; O=gpurun_out/$T; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "update_system or tick_matches_oracle or golden_vectors or centroidal" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python scripts/fp32_mode_report.py 16384 > $O/fp32_report.jsonl 2> $O/fp32_report.err; tail -2 $O/fp32_report.err
PNC_B200_LIB=$(realpath scripts/lib_f32c3.bin) timeout 300 python scripts/fp32_mode_report.py 16384 2>/dev/null | python -c "
import json,sys
for l in sys.stdin: d=json.loads(l); print('c3', d['robot'], d['update_ms'])"
for P in fp64 fp32; do echo -n "$P: "; PNC_PRECISION=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'])"; done
