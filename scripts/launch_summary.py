"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row.get("Metric Unit", "")
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("kernel, launches, total_us, avg_us, share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]}, {n}, {t:.1f}, {t / n:.2f}, {t / tot:.3f}")
