"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    name = row["Kernel Name"].replace("unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%-64s %6s %12s %10s %6s" % ("kernel", "n", "total_us", "avg_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %6d %12.1f %10.2f %5.1f%%" % (k[:64], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
