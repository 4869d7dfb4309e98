"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (development tool)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:100]}` | {v[0]} | {v[1]:.3f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]*1e3:.1f} |")
