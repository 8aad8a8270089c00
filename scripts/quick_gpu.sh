#!/bin/bash
# quick A/B on the GPU box: parity subset, headline bench (device only), per-kernel launch list.  Usage: $0 tag [robot]
T=${1:-q}
R=${2:-atlas}
O=gpurun_out/$T
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tick_matches_oracle or golden_vectors or large_sample" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python bench.py --robot $R --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/bench.json 2> $O/bench.err && python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['kernels_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --robot $R --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu.log 2>&1
python scripts/launch_times.py $O/launches.csv
