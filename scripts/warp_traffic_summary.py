"""Per-kernel DRAM traffic and time of ONE fused-warp call at BASELINE config 3 (16384 crops) from an ncu launch list.
usage: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fusg --csv \
           --log-file profiles/r2_warp_config3_launches.csv python scripts/bench_warp.py 16384 1
       python scripts/warp_traffic_summary.py profiles/r2_warp_config3_launches.csv 16384 > profiles/r2_warp_traffic_summary.json"""
import csv
import json
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
B = int(sys.argv[2])
per = OrderedDict()
for r in rows[1:]:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
    per.setdefault(int(r[idi]), {"k": r[ki].split("(")[0].replace("void ", "")})[r[mi]] = float(r[vi].replace(",", "")) * scale
ids = sorted(per)
# the last call = the last run of kernels starting at k_visibility
starts = [i for i in ids if "k_visibility" in per[i]["k"]]
ids = [i for i in ids if i >= starts[-1]]
kernels = [{"kernel": per[i]["k"], "duration_us": per[i]["gpu__time_duration.sum"] / 1e3, "dram_read_bytes": int(per[i]["dram__bytes_read.sum"]),
            "dram_write_bytes": int(per[i]["dram__bytes_write.sum"])} for i in ids]
rd, wr = sum(k["dram_read_bytes"] for k in kernels), sum(k["dram_write_bytes"] for k in kernels)
alg = B * 256 * 256 * 3 * 6
print(json.dumps({"source": sys.argv[1] + " (ncu, serialised launches of one fusg_warp_fused call)", "crops": B, "algorithmic_bytes": alg,
                  "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_over_algorithmic": (rd + wr) / alg,
                  "kernel_time_us_sum": sum(k["duration_us"] for k in kernels), "kernels": kernels}, indent=1))
