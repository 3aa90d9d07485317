"""Summarise an ncu launch list with dram bytes + duration per k_conv_tc launch (one VUNet forward, B = 64).
usage: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv \
           --csv --log-file profiles/r1_conv_traffic.csv python scripts/layer_paths.py 64 > gpurun_out/paths.txt
       python scripts/traffic_summary.py gpurun_out/paths.txt profiles/r1_conv_traffic.csv > profiles/r1_conv_traffic_summary.json
Also writes the per-layer table next to it (r1_layers_ncu.txt)."""
import csv
import json
import sys

paths = [l.split() for l in open(sys.argv[1]) if l.strip()]
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10 and r[0].isdigit()]
hdr_ok = {}
per = {}
for r in rows:
    lid, name, grid, metric, unit, val = int(r[0]), r[4], r[8], r[-3], r[-2], float(r[-1].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per.setdefault(lid, {"grid": grid})[metric] = val * scale
ids = sorted(per)[-len(paths):]
rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)
tt = sum(per[i]["gpu__time_duration.sum"] for i in ids)
top = max(ids, key=lambda i: per[i]["gpu__time_duration.sum"])
tp = paths[ids.index(top)]
out = {"source": sys.argv[2] + ": ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the "
                 f"{len(ids)} convolution launches of one VUNet forward, B=64 (scripts/layer_paths.py)",
       "launches": len(ids), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
       "avg_traffic_bytes_per_launch": int((rd + wr) / len(ids)), "kernel_time_ms_sum": tt / 1e6,
       "top_launch": {"layer": tp[0], "dram_read_bytes": int(per[top]["dram__bytes_read.sum"]),
                      "dram_write_bytes": int(per[top]["dram__bytes_write.sum"]), "duration_us": per[top]["gpu__time_duration.sum"] / 1e3,
                      "flops": float(tp[2])}}
print(json.dumps(out, indent=1))
with open(sys.argv[2].replace("_conv_traffic.csv", "_layers_ncu.txt"), "w") as f:
    f.write(f"{len(ids)} convolution launches of one forward (B=64), ncu serialised/cold-cache: {tt/1e6:.3f} ms, "
            f"{rd/1e9:.2f} GB read + {wr/1e9:.2f} GB written\n")
    f.write(f"{'layer':46s} {'grid':>5s} {'GFLOP':>9s} {'us':>8s} {'TFLOP/s':>8s} {'GB moved':>9s} {'TB/s':>6s}\n")
    for i, (p, impl, fl) in zip(ids, paths):
        us = per[i]["gpu__time_duration.sum"] / 1e3
        gb = (per[i]["dram__bytes_read.sum"] + per[i]["dram__bytes_write.sum"]) / 1e9
        g = per[i]["grid"].strip("()").split(",")[0]
        f.write(f"{p:46s} {g:>5s} {float(fl)/1e9:9.2f} {us:8.1f} {float(fl)/us/1e6:8.1f} {gb:9.3f} {gb/us*1e3:6.2f}\n")
