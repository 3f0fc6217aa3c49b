"""Split an ncu source-page CSV of k_sweep into code regions by execution count and summarise instructions / stalls.
usage: python tools/ncu_regions.py src.csv [ntiles_x_slices]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def num(x):
    try: return float(x)
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
segs = []; cur = None
for r in data:
    n = num(r[idx["Instructions Executed"]])
    key = "tile" if n >= 150000 else ("mid" if 2000 < n < 150000 else ("field" if 500 <= n <= 2000 else "rare"))
    if cur is None or cur["key"] != key:
        cur = dict(key=key, addr=r[idx["Address"]][-6:], static=0, dyn=0.0, samples=0.0, st=collections.Counter(), ops=collections.Counter(), line=r[idx["Source"]][:40])
        segs.append(cur)
    cur["static"] += 1; cur["dyn"] += n
    src = r[idx["Source"]].split()
    op = (src[1] if src and src[0].startswith("@") else (src[0] if src else "?")).split(".")[0]
    cur["ops"][op] += n
    for s in stalls:
        v = num(r[idx[s]])
        if s != "stall_barrier": cur["samples"] += v
        cur["st"][s] += v
tot = sum(s["dyn"] for s in segs)
print("dyn instr total %.4g" % tot)
for s in segs:
    if s["dyn"] > tot * 0.01 or s["samples"] > 800:
        top = ", ".join("%s %.0f" % (k[6:], v) for k, v in s["st"].most_common(5) if k != "stall_barrier")
        ops = ", ".join("%s %.2g" % kv for kv in s["ops"].most_common(6))
        print(f'{s["key"]:5s} {s["addr"]} static {s["static"]:5d} dyn {s["dyn"]:.3g} ({100*s["dyn"]/tot:4.1f}%) samples {s["samples"]:7.0f} | {top} | {ops}')
