#!/bin/bash
# Round-1 evidence run (one gpurun call): tests, smoke, bench lines, ncu launch list and full captures.
set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for r in laikago valkyrie draco3 draco3_lb; do timeout 300 python bench.py --robot $r --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_$r.json 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ihwbc_reg|k_ihwbc_setup|k_update_system|k_ihwbc_finish|k_task_prepare" -c 8 -o $O/r1_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
ls -la $O
