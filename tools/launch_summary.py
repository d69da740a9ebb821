"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (k, n, v / 1e3, 100 * v / tot, v / n))
print("\nTotal %.1f ms over %d launches." % (tot / 1e3, sum(n for n, _ in agg.values())))
