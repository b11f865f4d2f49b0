"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
f = sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{f}: total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k}")
