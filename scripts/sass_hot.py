"""Scratch: hot spots of one kernel from `ncu --page source --csv --print-source sass` output.
usage: sass_hot.py file.csv <kernel-index> [top]"""
import csv, sys
path, which = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
with open(path) as f:
    for row in csv.reader(f):
        if row and row[0] == "Kernel Name":
            cur = dict(name=row[1], rows=[], hdr=None); blocks.append(cur); continue
        if cur is None or not row: continue
        if cur["hdr"] is None: cur["hdr"] = row; continue
        cur["rows"].append(row)
b = blocks[which]
h = {k: i for i, k in enumerate(b["hdr"])}
rows = b["rows"]
ie = [int(r[h["Instructions Executed"]] or 0) for r in rows]
sm = [int(r[h["# Samples"]] or 0) for r in rows]
tot, tots = sum(ie), sum(sm)
print(b["name"], "instr", tot, "samples", tots, "n_sass", len(rows))
# classify
import re, collections
cls = collections.Counter(); clss = collections.Counter()
for r, n, s in zip(rows, ie, sm):
    op = r[h["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    cls[op] += n; clss[op] += s
print("by opcode (instr share / sample share):")
for op, n in cls.most_common(22):
    print(f"   {op:12s} {100*n/tot:5.1f}%  {100*clss[op]/max(tots,1):5.1f}%")
print("top sampled instructions:")
order = sorted(range(len(rows)), key=lambda i: -sm[i])[:top]
for i in sorted(order):
    r = rows[i]
    stalls = {k[6:]: int(r[h[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}
    st = ",".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:2] if v)
    print(f"   [{i:5d}] {100*sm[i]/max(tots,1):5.2f}% smp  {100*ie[i]/tot:5.2f}% ins  {r[h['Source']].strip()[:70]:70s} {st}")
