"""Turn gpurun_out/<tag>/ (written by scripts/r2_gpu_evidence.sh on the GPU box) into the committed profiles/r2_* summaries.
Usage: python scripts/r2_profiles.py [tag]   (default r2)"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'
F, P = os.path.join(ROOT, 'gpurun_out', TAG), os.path.join(ROOT, 'profiles')
R = os.path.join(F, TAG + '_full.ncu-rep')
shutil.copy(os.path.join(F, 'launches.csv'), os.path.join(P, 'r2_launches.csv'))
shutil.copy(os.path.join(F, 'bench_full.json'), os.path.join(P, 'r2_bench_full.json'))
shutil.copy(os.path.join(F, 'bench_reference.json'), os.path.join(P, 'r2_bench_reference.json'))
shutil.copy(os.path.join(F, 'pytest_gpu.log'), os.path.join(P, 'r2_pytest_gpu.log'))
last = lambda f: json.loads(open(os.path.join(F, f)).read().strip().splitlines()[-1])
rows = list(csv.reader(subprocess.run(['ncu', '-i', R, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
kern, total = {}, 0.0
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d['Kernel Name'].split('(')[0].replace('void ', '').strip()
    if name in kern:
        continue
    rd, wr = float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])
    unit_r, unit_w = rows[1][h.index('dram__bytes_read.sum')], rows[1][h.index('dram__bytes_write.sum')]
    scale = {'Gbyte': 1.0, 'Mbyte': 1e-3, 'Kbyte': 1e-6, 'byte': 1e-9}
    rd, wr = rd * scale.get(unit_r, 1.0), wr * scale.get(unit_w, 1.0)
    st = {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(float(v), 2)
          for k, v in d.items() if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and v}
    top = dict(sorted(st.items(), key=lambda x: -x[1])[:6])
    kern[name] = {'ms': round(float(d['gpu__time_duration.sum']), 3), 'dram_read_gb': round(rd, 4), 'dram_write_gb': round(wr, 4),
                  'warp_instructions': int(float(d['smsp__inst_executed.sum'])), 'registers': int(float(d['launch__registers_per_thread'])),
                  'fp64_pipe_pct': round(float(d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0) or 0), 1),
                  'warps_per_sm_theoretical': round(float(d.get('sm__maximum_warps_per_active_cycle_pct', 0) or 0) * 0.64, 1),
                  'stall_cycles_per_issue': top}
    if not name.startswith('k_update_system'):
        total += rd + wr
full = last('bench_full.json')
t = {"_comment": "per robot: dram__bytes_read.sum + dram__bytes_write.sum per launch of the solve stage (solve_stage_bytes) from one `ncu --set full --clock-control none` capture of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs` (scripts/r2_gpu_evidence.sh -> gpurun_out/%s/%s_full.ncu-rep, summarised in r2_ncu_details.txt; one B200; ms = gpu__time_duration of that serialised, cold-cache capture), and flops_per_solve = mean floating-point operations per solve of the reference algorithm on the bench workload, counted by the oracle in the bench run of the same call (bench.py recounts it whenever the CPU baseline leg runs and falls back to this figure otherwise)" % (TAG, TAG),
     "atlas": {"batch": 65536, "kernels": kern, "solve_stage_bytes": int(total * 1e9), "flops_per_solve": round(full['roofline']['flops_per_solve'])}}
for k, v in full.get('configs', {}).items():
    rf = v.get('roofline') or {}
    rb = {'config2_laikago_16k': 'laikago', 'config4_draco3_64k': 'draco3', 'config5_valkyrie_256k': 'valkyrie'}.get(k)
    if rb and 'flops_per_solve' in rf and 'oracle' in str(rf.get('flops_source', '')):
        t[rb] = {"batch": v.get('batch_per_gpu'), "solve_stage_bytes": None, "flops_per_solve": round(rf['flops_per_solve'])}
json.dump(t, open(os.path.join(P, 'r2_traffic.json'), 'w'), indent=1)
open(os.path.join(P, 'r2_ncu_details.txt'), 'w').write(subprocess.run(['ncu', '-i', R, '--page', 'details'], capture_output=True, text=True).stdout)
src = os.path.join(ROOT, 'pypnc_b200', 'csrc', 'ihwbc_reg.cu')
for k, top in (('k_ihwbc_reg', '40'), ('k_ihwbc_setup', '30')):
    for script, extra, suffix in (('ncu_regions.py', [src, '65536'], 'regions'), ('ncu_lines.py', ['65536', top], 'top_lines')):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', script), R] + extra + [k], capture_output=True, text=True).stdout
        open(os.path.join(P, 'r2_%s_%s.txt' % (k, suffix)), 'w').write(txt)
print(json.dumps({k: v['ms'] for k, v in kern.items()}), t['atlas']['solve_stage_bytes'])
d = full
print('headline %.4g solves/s, %.3f ms/step, e2e %.4g, roofline %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
for k, v in d.get('configs', {}).items():
    print(k, '%.4g' % v['value'], round(v['ms_per_step'], 3), 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'))
