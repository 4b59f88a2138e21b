"""Summarise an `ncu --page source --csv` export: the instructions with the most stall samples and their dominant stall reasons.
usage: python tools/src_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    s = int(r[isamp] or 0)
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    data.append((k, s, int(r[iexec] or 0), r[isrc].strip(), st))
tot = sum(d[1] for d in data)
print("total samples", tot, "instructions", len(data))
for k, s, ex, src, st in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{k:5d} {s:7d} {100*s/tot:5.1f}% exec={ex:9d} {src[:60]:60s} " + " ".join(f"{n}:{v}" for v, n in st if v))
