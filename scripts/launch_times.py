"""Mean duration per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r'\(.*', '', r[4]).replace('void ', '')
    agg.setdefault(name, []).append(float(r[14].replace(',', '')))
for k, v in agg.items():
    v2 = v[len(v) // 3:] if len(v) >= 3 else v  # skip the warm-up launches
    print("%-50s n=%3d mean %9.1f us  min %9.1f us" % (k[:50], len(v), sum(v2) / len(v2) / 1e3, min(v) / 1e3))
