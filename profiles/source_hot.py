#!/usr/bin/env python
"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass | python profiles/source_hot.py [N] [kernel-substring]"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(sys.stdin))
fpath, func, hdr = None, None, None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Kernel Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if want and want not in (func or ""): continue
    try: line = int(r[0])
    except ValueError: continue
    if r[2] != "-": continue   # keep the per-source-line aggregate rows only (Address == '-')
    i_s = hdr.index("# Samples")
    try: s = int(r[i_s])
    except ValueError: continue
    if s == 0: continue
    stalls = {}
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: v = int(r[j])
            except ValueError: v = 0
            if v: stalls[h[6:]] = v
    key = (func, fpath.split("/")[-1] if fpath else "?", line)
    a = agg.setdefault(key, [0, r[1].strip()[:110], {}, 0])
    a[0] += s
    try: a[3] += int(r[hdr.index("Instructions Executed")])
    except ValueError: pass
    for k, v in stalls.items(): a[2][k] = a[2].get(k, 0) + v
tot = sum(a[0] for a in agg.values())
print(f"total samples {tot}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{a[0]:7d} {100*a[0]/max(tot,1):5.1f}%  {key[1]}:{key[2]:<5d} inst={a[3]:<9d} {top}  | {a[1]}")
