"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / share."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(',', ''))
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1e-3)
        d[r[ki].split('(')[0][-60:]].append(v * scale)
tot = sum(sum(v) for v in d.values())
print("%-62s %6s %12s %8s" % ("kernel", "count", "total_us", "share"))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-62s %6d %12.1f %7.1f%%" % (k, len(v), sum(v), 100 * sum(v) / tot))
