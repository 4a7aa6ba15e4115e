"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms and share per kernel."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v * 1e3 if u == 's' else v
    k = r[ki].split('(')[0][:60]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("%-50s %6s %12s %7s %10s" % ("kernel", "n", "total ms", "share", "ms/launch"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s %6d %12.3f %6.1f%% %10.4f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
