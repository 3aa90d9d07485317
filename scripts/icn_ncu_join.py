"""Join scripts/icn_ncu.py's launch names with the ncu metric list of the same run (see icn_ncu.py)."""
import collections
import csv
import sys

names = [l.split() for l in open(sys.argv[1]) if l.strip()]
lines = open(sys.argv[2]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
by_id = collections.OrderedDict()
for r in csv.DictReader(lines[start:]):
    by_id.setdefault(r["ID"], {"kernel": r["Kernel Name"].split("(")[0]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
launches = list(by_id.values())[-len(names):]
tot = sum(l["gpu__time_duration.sum"] for l in launches) / 1e6
print(f"{len(names)} launches of one 64-crop ICN forward, {tot:.3f} ms of kernel time (ncu, cold-cache serialised, --clock-control none)")
print(f"{'launch':44s} {'kernel':34s} {'ms':>7s} {'TFLOP/s':>8s} {'tensor %':>8s} {'DRAM rd MB':>10s} {'wr MB':>8s} {'GB/s':>6s} {'SM GHz':>6s}")
for (name, kind, amount), l in zip(names, launches):
    ms = l["gpu__time_duration.sum"] / 1e6
    rd, wr = l.get("dram__bytes_read.sum", 0.0), l.get("dram__bytes_write.sum", 0.0)
    tf = float(amount) / ms / 1e9 if kind == "flops" else 0.0
    print(f"{name:44s} {l['kernel'][-34:]:34s} {ms:7.3f} {tf:8.1f} {l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0.0):8.1f} "
          f"{rd / 1e6:10.1f} {wr / 1e6:8.1f} {(rd + wr) / ms / 1e6:6.0f} {l.get('sm__cycles_elapsed.avg.per_second', 0.0) / 1e9:6.2f}")
