"""aggregate an ncu --csv launch list (gpu__time_duration + dram bytes) per kernel"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > vi:
        d.setdefault((r[0], r[ki]), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for (_, k), m in d.items():
    a = agg[k.split("(")[0][:64]]
    a[0] += 1
    t, u = m["gpu__time_duration.sum"]
    a[1] += t * T.get(u, 1)
    for j, n in ((2, "dram__bytes_read.sum"), (3, "dram__bytes_write.sum")):
        if n in m:
            v, u = m[n]
            a[j] += v * B.get(u, 1)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
print(f"{'kernel':64s} {'n':>4s} {'ms/launch':>10s} {'rd GB':>8s} {'wr GB':>8s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k:64s} {a[0]:4d} {a[1] / a[0]:10.3f} {a[2] / a[0] / 1e9:8.3f} {a[3] / a[0] / 1e9:8.3f} {(a[2] + a[3]) / a[1] / 1e6:7.0f}")
