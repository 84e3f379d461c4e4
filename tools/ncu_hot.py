"""Reads `ncu -i X.ncu-rep --page source --csv` output and prints the SASS instructions with the most stall samples.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_spmv | python tools/ncu_hot.py [N]
"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[start + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # next kernel instance in the report
    if len(r) == len(hdr):
        body.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_")]
print(f"total samples {tot}; instructions {len(body)}")
pos = {id(r): i for i, r in enumerate(body)}
top = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]
for r in sorted(top, key=lambda r: pos[id(r)]):
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{pos[id(r)]:5d} {100.0 * s / max(tot, 1):5.1f}%  {r[ix['Source']].strip():70s} {why}")
