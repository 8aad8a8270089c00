O=gpurun_out/qreg; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tick_matches_oracle or golden or edge" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ihwbc_reg|k_ihwbc_setup" -c 2 -o $O/reg_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_full.log 2>&1
ls -la $O
