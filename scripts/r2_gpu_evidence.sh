#!/bin/bash
# Round-2 evidence run (one gpurun call): tests, smoke, the bench line (all five BASELINE configs), the reference arm,
# the ncu launch list of the bench command and one full-set capture of the step's kernels.  Usage: $0 [tag]
T=${1:-r2}
O=gpurun_out/$T
mkdir -p $O
set -x
timeout 1500 python -m pytest tests -x -q -m gpu -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/plain2.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ihwbc_reg|k_ihwbc_setup|k_update_system|k_ihwbc_finish|k_task_prepare|k_wbc" -c 8 -o $O/${T}_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_full.log 2>&1
timeout 600 python scripts/fp32_mode_report.py 16384 > $O/fp32_report.jsonl 2> $O/fp32_report.err
ls -la $O
