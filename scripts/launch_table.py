"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table (markdown): python scripts/launch_table.py in.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("hos::", "")[:60]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui].startswith("ns") else (v * 1e3 if r[ui].startswith("ms") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
print(f"| all | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100 % |")
