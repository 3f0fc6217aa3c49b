"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, total, share."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(list)
unit = "ns"
for row in csv.DictReader(lines):
    agg[row["Kernel Name"].split("(")[0][:48]].append(float(row["Metric Value"].replace(",", "")))
    unit = row["Metric Unit"]
tot = sum(sum(v) for v in agg.values())
print(f"# {path}: {sum(len(v) for v in agg.values())} launches, total {tot/1e3:.1f} us ({unit})")
print(f"{'kernel':50s} {'n':>5s} {'mean_us':>10s} {'total_us':>12s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} {len(v):5d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/1e3:12.1f} {sum(v)/tot:7.3f}")
