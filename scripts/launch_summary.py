"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/launch_summary.py launches.csv [skip_first_n]"""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
spin = [r for r in rows if "spin_kernel" in r["Kernel Name"]]
rows = [r for r in rows if "spin_kernel" not in r["Kernel Name"]]
if spin:
    # torch.cuda._sleep: bench.py parks the stream behind it before its per-launch roofline pass (outside the timed steps)
    print(f"(excluded: {len(spin)} torch.cuda._sleep spin launches, {sum(float(r['Metric Value']) for r in spin) / 1e6:.1f} ms -- bench.py's per-launch timing pass)")
for r in rows:
    name = r["Kernel Name"].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.3f} ms of kernel time (ncu, cold-cache serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} n={v[0]:5d} total={v[1]:9.3f} ms  share={100 * v[1] / tot:5.1f}%")
