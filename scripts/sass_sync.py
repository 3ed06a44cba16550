"""Scratch: every barrier wait / MMA / bulk-copy / TMEM instruction of one kernel with its sample count and stalls."""
import csv, sys
path, which = sys.argv[1], int(sys.argv[2])
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
sm = [int(r[h["# Samples"]] or 0) for r in rows]
ie = [int(r[h["Instructions Executed"]] or 0) for r in rows]
tots = sum(sm)
keys = ("TRYWAIT", "UTCHMMA", "UTCBAR", "UBLKCP", "LDTM", "ARRIVE", "BAR.SYNC", "WARPSYNC", "NANOSLEEP")
for i, r in enumerate(rows):
    src = r[h["Source"]].strip()
    nxt = sm[i + 1] if i + 1 < len(rows) else 0
    nxt2 = sm[i + 2] if i + 2 < len(rows) else 0
    nxt3 = sm[i + 3] if i + 3 < len(rows) else 0
    if any(k in src for k in keys):
        stalls = {k[6:]: int(r[h[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}
        st = ",".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"[{i:5d}] smp {sm[i]:6d} (+next3 {nxt + nxt2 + nxt3:6d}) exec {ie[i]:9d}  {src[:90]:90s} {st}")
print("total samples", tots)
