"""Join scripts/layer_paths.py output with the ncu launch list of the same run: real per-layer kernel durations."""
import csv
import sys
paths = [l.split() for l in open(sys.argv[1]) if l.strip()]
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10 and r[0].isdigit()]
rows = rows[-len(paths):]
tot = sum(float(r[-1]) for r in rows) / 1e3
print(f"{len(paths)} layers, {tot:.1f} us of kernel time (ncu, serialised)")
for (p, impl, fl), r in zip(paths, rows):
    us = float(r[-1]) / 1e3
    print(f"{p:46s} grid={r[8]:14s} {float(fl)/1e9:9.2f} GFLOP {us:9.1f} us {float(fl)/us/1e6:8.1f} TFLOP/s {100*us/tot:5.2f}%")
