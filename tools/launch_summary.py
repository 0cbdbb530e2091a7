"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (developer tool)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[ci['Metric Value']].replace(',', ''))
    except Exception:
        continue
    unit = r[ci['Metric Unit']]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
    agg[r[ci['Kernel Name']].split('(')[0]][0] += 1
    agg[r[ci['Kernel Name']].split('(')[0]][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} n={n:5d} total={t:11.1f} us avg={t/n:9.1f} us {100*t/tot:5.1f}%")
print(f"total {tot:.1f} us")
