"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_<tag>.txt"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    a = agg.setdefault(name, [0, 0.0, []])
    a[0] += 1
    a[1] += v
    a[2].append(v)
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
import statistics
med_tot = sum(statistics.median(a[2]) * a[0] for a in agg.values())
print("# 'median share' uses n x median(us) per kernel: warm-up launches (e.g. the first render of a mesh, whose tile lists "
      "overflow their estimated capacity and fall back to whole-mesh scans) do not distort it")
print(f"{'total us':>10} {'n':>5} {'mean us':>10} {'median us':>10} {'share':>6} {'median share':>12}  kernel")
for k, (n, t, vals) in sorted(agg.items(), key=lambda x: -statistics.median(x[1][2]) * x[1][0]):
    m = statistics.median(vals)
    print(f"{t:10.1f} {n:5d} {t / n:10.1f} {m:10.1f} {100 * t / tot:5.1f}% {100 * m * n / med_tot:11.1f}%  {k[:100]}")
