"""Hottest SASS instructions (warp-stall samples) of one launch in an ncu report's source page.
    ncu -i rep --page source --csv --kernel-id ::regex:NAME:N > src.csv ; python tools/ncu_hot.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Address", "Kernel Name") or len(r) != len(hdr):
        break
    data.append(r)
si, src = hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for c in stall_cols:
        if r[c]:
            agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c])
print("stall totals", sorted(agg.items(), key=lambda x: -x[1])[:8])
top = sorted(enumerate(data), key=lambda x: -int(x[1][si] or 0))[:top_n]
for i, r in sorted(top):
    st = {hdr[c][6:]: int(r[c]) for c in stall_cols if r[c] and int(r[c]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(i, r[si], r[src].strip()[:80], st)
