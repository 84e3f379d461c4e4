"""Instruction and stall-sample share per barrier-delimited SASS segment (input: ncu --page source --csv)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[start + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        body.append(r)
tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tots = sum(int(r[ix["# Samples"]] or 0) for r in body)
seg = acc = accs = first = 0
print("total warp-instr", tot, "samples", tots)
for i, r in enumerate(body):
    src = r[ix["Source"]].strip()
    acc += int(r[ix["Instructions Executed"]] or 0)
    accs += int(r[ix["# Samples"]] or 0)
    if "BAR.SYNC" in src or "EXIT" in src or i == len(body) - 1:
        print(f"seg {seg:2d} sass[{first:4d}-{i:4d}] instr {100 * acc / max(tot, 1):5.1f}%  samples "
              f"{100 * accs / max(tots, 1):5.1f}%  ends with {src[:50]}")
        seg += 1
        acc = accs = 0
        first = i + 1
