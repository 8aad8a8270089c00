O=gpurun_out/q3; mkdir -p $O
timeout 300 python bench.py --robot draco3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_draco3.csv python bench.py --robot draco3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu.log 2>&1
python scripts/launch_times.py $O/launches_draco3.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ic_prepare|k_ihwbc_setup" -c 2 -o $O/draco3_full -f python bench.py --robot draco3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ihwbc_setup" -c 1 -o $O/atlas_setup -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_full2.log 2>&1
ls -la $O
