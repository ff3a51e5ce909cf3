"""Summarise the source page of an ncu report (ncu -i X.ncu-rep --page source --csv > X.csv): warp-stall samples per
opcode and per stall reason, and the hottest instructions.  Usage: python tools/ncu_stalls.py X.csv [min_share]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
kernel = rows[0][1] if len(rows[0]) > 1 else "?"
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
by_op, by_reason = collections.Counter(), collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    by_op[op] += int(r[ix["# Samples"]])
    for h in stalls:
        by_reason[h[6:]] += int(r[ix[h]])
print(f"kernel: {kernel}\ntotal warp-stall samples: {tot}\n")
print("| stall reason | samples | share |\n|---|---|---|")
for k, v in by_reason.most_common(10):
    print(f"| {k} | {v} | {v / tot:.1%} |")
print("\n| opcode | samples | share |\n|---|---|---|")
for k, v in by_op.most_common(12):
    print(f"| {k} | {v} | {v / tot:.1%} |")
print(f"\ninstructions holding >= {min_share:.0%} of the samples:\n\n| SASS | samples | executed (warp level) | top stall reasons |\n|---|---|---|---|")
for r in data:
    s = int(r[ix["# Samples"]])
    if s >= tot * min_share:
        top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"| `{r[ix['Source']].strip()[:60]}` | {s} | {r[ix['Instructions Executed']]} | {top[0][1]} {top[0][0]}, {top[1][1]} {top[1][0]} |")
