"""Summarise an `ncu --metrics gpu__time_duration.sum[,...] --csv` launch list: per kernel name, launches, total and mean time.
usage: python scripts/launch_times.py file.csv [last_n_launches]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launch = OrderedDict()
for r in rows[1:]:
    launch.setdefault(r[idi], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launch, key=int)
if len(sys.argv) > 2:
    ids = ids[-int(sys.argv[2]):]
agg = OrderedDict()
for i in ids:
    d = launch[i]
    a = agg.setdefault(d["k"].split("(")[0], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("smsp__inst_executed.sum", 0.0)
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{k[:48]:48s} n={a[0]:4d} total {a[1] / 1e3:10.1f} us  mean {a[1] / a[0] / 1e3:9.1f} us  share {100 * a[1] / tot:5.1f}%  inst/launch {a[2] / a[0]:.3g}")
