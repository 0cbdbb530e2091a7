# clean run, then the ncu launch list of the same command, summarised per kernel.  usage: bash tools/fit_launches.sh <tag> [N]
TAG=${1:-x}; N=${2:-4096}
python tools/fit_launches.py $N 5 | tee gpurun_out/${TAG}_fit.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/fit_launches.py $N 1 > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict(); seq = []
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000 if r[ui] in ("ns", "nsecond") else v
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; seq.append((name, v))
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/${TAG}_launch_summary.txt", "w") as f:
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-60s n=%5d total=%10.1f us avg=%8.1f us %5.1f%%\n" % (k[:60], n, t, t / n, 100 * t / tot))
    f.write("total %.1f us\n" % tot)
    f.write("sequence: " + " ".join("%s:%.0f" % (k[:14], v) for k, v in seq[:400]) + "\n")
print(open("gpurun_out/${TAG}_launch_summary.txt").read()[:6000])
PY
rm -f gpurun_out/${TAG}_launches.csv
