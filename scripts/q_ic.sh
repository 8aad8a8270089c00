O=gpurun_out/qic; mkdir -p $O
timeout 300 python bench.py --robot draco3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ic_prepare" -c 1 -o $O/ic_full -f python bench.py --robot draco3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_full.log 2>&1
ls -la $O
