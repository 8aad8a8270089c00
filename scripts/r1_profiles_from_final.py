"""Turn gpurun_out/final/ (written by scripts/r1_final_gpu.sh on the GPU box) into the committed profiles/ summaries."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, P = os.path.join(ROOT, 'gpurun_out', 'final'), os.path.join(ROOT, 'profiles')
R = os.path.join(F, 'r1_full.ncu-rep')
shutil.copy(os.path.join(F, 'launches.csv'), os.path.join(P, 'r1_launches.csv'))
shutil.copy(os.path.join(F, 'bench_full.json'), os.path.join(P, 'r1_bench_full.json'))
shutil.copy(os.path.join(F, 'bench_reference.json'), os.path.join(P, 'r1_bench_reference.json'))
last = lambda f: json.loads(open(os.path.join(F, f)).read().strip().splitlines()[-1])
out = {}
for r in ['laikago', 'valkyrie', 'draco3', 'draco3_lb']:
    d = last('bench_%s.json' % r)
    out[r] = {'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernels_ms': d['kernels_ms'], 'gpu_launches': d['gpu_launches'],
              'qp_dims_n_p_m': d['config']['qp_dims_n_p_m'], 'mean_active_set_iterations': d['solver']['mean_active_set_iterations']}
json.dump(out, open(os.path.join(P, 'r1_bench_other_robots.json'), 'w'), indent=1)
rows = list(csv.reader(subprocess.run(['ncu', '-i', R, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
kern, total = {}, 0.0
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d['Kernel Name'].split('(')[0].replace('void ', '').strip()
    if name in kern:
        continue
    rd, wr = float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])  # Gbyte
    kern[name] = {'ms': round(float(d['gpu__time_duration.sum']), 3), 'dram_read_gb': round(rd, 4), 'dram_write_gb': round(wr, 4),
                  'warp_instructions': int(float(d['smsp__inst_executed.sum'])), 'registers': int(float(d['launch__registers_per_thread']))}
    if not name.startswith('k_update_system'):
        total += rd + wr
full = last('bench_full.json')
t = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full --clock-control none` capture of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e` (scripts/r1_final_gpu.sh -> gpurun_out/final/r1_full.ncu-rep, summarised in r1_ncu_details.txt); Atlas, 65536 states, one B200. ms = gpu__time_duration of that (serialised, cold-cache) capture.",
     "robot": "atlas", "batch": 65536, "kernels": kern, "solve_stage_bytes": int(total * 1e9),
     "flops_per_solve": round(full['roofline']['flops_per_solve']),
     "flops_note": "mean floating-point operations per solve of the reference algorithm on this workload, counted by the oracle (every multiply/add of assembly, Cholesky, inverse and the Goldfarb-Idnani iterations); bench.py recounts it whenever the CPU baseline leg runs and falls back to this figure otherwise"}
json.dump(t, open(os.path.join(P, 'r1_traffic.json'), 'w'), indent=1)
open(os.path.join(P, 'r1_ncu_details.txt'), 'w').write(subprocess.run(['ncu', '-i', R, '--page', 'details'], capture_output=True, text=True).stdout)
src = os.path.join(ROOT, 'pypnc_b200', 'csrc', 'ihwbc_reg.cu')
for k, top in (('k_ihwbc_reg', '40'), ('k_ihwbc_setup', '30')):
    for script, extra, suffix in (('ncu_regions.py', [src, '65536'], 'regions'), ('ncu_lines.py', ['65536', top], 'top_lines')):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', script), R] + extra + [k], capture_output=True, text=True).stdout
        open(os.path.join(P, 'r1_%s_%s.txt' % (k, suffix)), 'w').write(txt)
print(json.dumps({k: v['ms'] for k, v in kern.items()}), t['solve_stage_bytes'])
for f in ['bench_full', 'bench_reference', 'bench_laikago', 'bench_valkyrie', 'bench_draco3', 'bench_draco3_lb']:
    d = last(f + '.json')
    print(f, '%.4g' % d['value'], round(d['ms_per_step'], 3), d.get('kernels_ms'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
          'roof', (d.get('roofline') or {}).get('frac'))
