"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/launch_shares.py gpurun_out/launches.csv [header line ...]"""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
rows = list(csv.reader(l for l in (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) if l.startswith('"')))
head, rows = rows[0], rows[1:]
ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[ki].split("(")[0]
    tot[name] += float(r[vi].replace(",", "")) * scale[r[ui]]
    cnt[name] += 1
allms = sum(tot.values())
for h in sys.argv[2:]:
    print("# " + h)
print("%-64s %6s %12s %7s" % ("kernel", "n", "total_ms", "share"))
for name, ms in tot.most_common():
    print("%-64s %6d %12.3f %6.1f%%" % (name[:64], cnt[name], ms, 100 * ms / allms))
print("%-64s %6d %12.3f" % ("total", sum(cnt.values()), allms))
