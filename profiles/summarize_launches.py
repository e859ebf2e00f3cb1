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
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
print(f"{'total us':>10} {'n':>5} {'us/launch':>10} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} {n:5d} {t / n:10.1f} {100 * t / tot:5.1f}%  {k[:110]}")
