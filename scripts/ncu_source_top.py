#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump: python scripts/ncu_source_top.py file.source.csv[.gz] [N]"""
import csv, gzip, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:90]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2: body.append(rows[j])
            j += 1
        si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si] or 0) for r in body)
        print(f"== {name}: {len(body)} instr, {tot} samples")
        agg = {hdr[k]: sum(int(r[k] or 0) for r in body) for k in stall_cols}
        print("   stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
        idx = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:topn]
        for k in sorted(idx):
            r = body[k]
            st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
            print(f"   #{k:5d} {int(r[si] or 0):6d} ({100.0*int(r[si] or 0)/max(tot,1):4.1f}%) ex={r[ie]:>8} {r[src].strip()[:70]:70s} " + " ".join(f"{n}={v}" for v, n in st if v))
        i = j
    else:
        i += 1
