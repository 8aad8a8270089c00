"""Summarise an ncu report by CUDA source line: stall samples, executed instructions, stall reasons."""
import csv, subprocess, sys
rep, nstates = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
KFILTER = ['-k', 'regex:' + sys.argv[4]] if len(sys.argv) > 4 else []  # optional kernel-name filter
txt = subprocess.run(['ncu', '-i', rep] + KFILTER + [ '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'Line No': hdr = r; stallk = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]; continue
    if hdr is None or len(r) < 8 or r[2] != '-': continue
    try: samp = float(r[6]); inst = float(r[7])
    except ValueError: continue
    out.append((samp, inst, r[0], r[1][:105], dict(zip(hdr, r))))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print('total samples %.0f, warp instructions %.0f = %.0f per state' % (tot, toti, toti / nstates))
tots = {k: 0.0 for k in stallk}
for o in out:
    for k in stallk: tots[k] += float(o[4][k] or 0)
print('stalls:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / tot) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]))
out.sort(key=lambda o: -o[0])
for o in out[:top]:
    d = o[4]
    st = sorted([(float(d[k] or 0), k) for k in stallk], reverse=True)[:2]
    print('%5.1f%% samp %5.1f%% inst  L%s  %s   [%s]' % (100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], ', '.join('%s %.0f%%' % (k[6:], 100 * v / max(o[0], 1)) for v, k in st)))
