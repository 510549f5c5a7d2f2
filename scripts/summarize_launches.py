"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of total device time)."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    tot[name] += v * scale; cnt[name] += 1
T = sum(tot.values())
print(f"total device time {T:.2f} ms over {sum(cnt.values())} launches")
print(f"{'kernel':60s} {'launches':>9s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:60]:60s} {cnt[k]:9d} {v:10.2f} {100*v/T:6.1f}% {1e3*v/cnt[k]:10.1f}")
