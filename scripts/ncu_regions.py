"""Executed-instruction and stall-sample share per code region (marker comments) of a kernel file."""
import csv, subprocess, sys, re, bisect
rep, srcfile, nstates = sys.argv[1], sys.argv[2], int(sys.argv[3])
src = open(srcfile).read().split('\n')
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(src) if re.match(r'\s*// (----|====)', l)]
KFILTER = ['-k', 'regex:' + sys.argv[4]] if len(sys.argv) > 4 else []  # optional kernel-name filter
txt = subprocess.run(['ncu', '-i', rep] + KFILTER + [ '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; cur_file = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != '-': continue
    try: samp = float(r[6]); inst = float(r[7]); ln = int(r[0])
    except ValueError: continue
    if cur_file and cur_file.endswith(srcfile.split('/')[-1]):
        k = bisect.bisect_right([m[0] for m in marks], ln) - 1
        key = '%4d %s' % marks[k] if k >= 0 else 'prologue'
    else:
        key = 'other: ' + (cur_file or '?').split('/')[-1]
    a = agg.setdefault(key, [0.0, 0.0]); a[0] += samp; a[1] += inst
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print('instr/state %.0f' % (ti / nstates))
for k, a in sorted(agg.items()):
    print('%5.1f%% samp %5.1f%% inst (%6.0f/state)  %s' % (100 * a[0] / ts, 100 * a[1] / ti, a[1] / nstates, k))
