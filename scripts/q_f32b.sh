T=${1:-qf32b}; O=gpurun_out/$T; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fp32_mode.py -x -q -m gpu -s > $O/pytest.log 2>&1; tail -12 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['kernels_ms'],d['e2e']['value']);print(d['fp32_mode'])"
