"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, total and share."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':40s} {'n':>4s} {'mean us':>10s} {'total ms':>10s} {'share':>7s}")
for k, v in agg.items():
    print(f"{k[:40]:40s} {len(v):4d} {sum(v)/len(v):10.1f} {sum(v)/1e3:10.3f} {sum(v)/tot*100:6.1f}%")
print(f"{'total':40s} {'':4s} {'':10s} {tot/1e3:10.3f}")
