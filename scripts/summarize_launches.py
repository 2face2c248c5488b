"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print(f"total {tot / 1e3:.2f} ms over {sum(c for c, _ in agg.values())} launches")
print(f"{'ms':>10} {'share':>7} {'launches':>9} {'avg us':>9}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{t / 1e3:10.3f} {100 * t / tot:6.2f}% {c:9d} {t / c:9.1f}  {k}")
