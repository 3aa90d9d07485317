"""Top instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output.
usage: python scripts/ncu_hot.py source.csv [N]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ci["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
stall_cols = [h for h in hdr if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
top = sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:n]
for r in top:
    extra = ""
    for h in hdr:
        if h.startswith("stall") and r[ci[h]] not in ("0", ""):
            extra += f" {h}={r[ci[h]]}"
    print(f"{int(r[ci['# Samples']]):7d} {100*int(r[ci['# Samples']])/tot:5.1f}% exec={r[ci['Instructions Executed']]:>9s} {r[ci['Source']].strip()[:90]}{extra[:200]}")
# instruction mix for global stores/loads
for key in ("STG", "LDG", "LDTM", "MUFU", "UTMALDG", "UTCHMMA"):
    c = sum(int(r[ci["Instructions Executed"]]) for r in body if key in r[ci["Source"]])
    print(key, c)
