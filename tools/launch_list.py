"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the launches above a threshold.
python tools/launch_list.py file.csv [min_us=150]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
min_ns = float(sys.argv[2]) * 1e3 if len(sys.argv) > 2 else 150e3
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
agg, seq = collections.OrderedDict(), []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name, t = r[ki].split("(")[0], float(r[vi].replace(",", ""))
    seq.append((name, t, r[gi], r[bi]))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"total {sum(v[1] for v in agg.values()) / 1e6:.3f} ms over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{v[1] / 1e6:9.3f} ms {v[0]:5d}  {k}")
for s in seq:
    if s[1] >= min_ns:
        print(f"  {s[1] / 1e6:8.3f} ms {s[2]:>16} x {s[3]:<14} {s[0]}")
