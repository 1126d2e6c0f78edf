#!/usr/bin/env python
"""Summarise an ncu report's SASS source page: top instructions by stall samples, with the stall reason split.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_source_top.py [topN]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
total = sum(int(r[col["# Samples"]] or 0) for r in body)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"instructions {len(body)}  samples {total}")
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in reasons}
print("by reason:", ", ".join(f"{h[6:]} {100 * v / max(total, 1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
ops = {}
for r in body:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    ops.setdefault(op.split(".")[0], [0, 0])
    ops[op.split(".")[0]][0] += int(r[col["Instructions Executed"]] or 0)
    ops[op.split(".")[0]][1] += int(r[col["# Samples"]] or 0)
tot_inst = sum(v[0] for v in ops.values())
print("by opcode (executed %, samples %):", ", ".join(f"{k} {100 * v[0] / tot_inst:.1f}/{100 * v[1] / max(total, 1):.1f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]))
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{idx:5d} {100 * n / max(total, 1):5.1f}%  {r[col['Source']].strip()[:70]:70s} {why}")
