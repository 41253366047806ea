"""Per-kernel summary of an ncu launch list (--csv with gpu__time_duration.sum [+ dram bytes])."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3,
         "nsecond": 1e-9, "ns": 1e-9, "second": 1, "s": 1}
for r in rows[1:]:
    name = r[ki].split("(")[0][:70]
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    d[name][r[mi]] += v
    if r[mi] == "gpu__time_duration.sum":
        cnt[name] += 1
tot = sum(x["gpu__time_duration.sum"] for x in d.values())
print(f"total {tot * 1e3:.3f} ms over {sum(cnt.values())} launches")
for k, x in sorted(d.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = x["gpu__time_duration.sum"]
    print(f"{k:70s} n={cnt[k]:4d} t={t * 1e3:8.3f} ms share={t / tot:.3f} rd={x['dram__bytes_read.sum'] / 1e6:9.1f}MB "
          f"wr={x['dram__bytes_write.sum'] / 1e6:9.1f}MB")
