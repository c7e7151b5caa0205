"""Summarise the source page of an ncu report: stall reasons overall, and the hottest SASS instructions.
   ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_src_summary.py src.csv [top=40] [lo hi]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
num = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, len(data))
sel = data[lo:hi]
tot = sum(num(r[ix["# Samples"]]) for r in sel)
print(f"samples {tot:.0f} over SASS instructions {lo}..{hi} of {len(data)}")
agg = {c: sum(num(r[ix[c]]) for r in sel) for c in stall_cols}
print("stalls:", ", ".join(f"{c[6:]}={v:.0f}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
top = sorted(range(lo, hi), key=lambda i: -num(data[i][ix["# Samples"]]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = sorted(((c[6:], num(r[ix[c]])) for c in stall_cols), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {num(r[ix['# Samples']]):6.0f} x{num(r[ix['Instructions Executed']]):8.0f}  {st[0][0]}={st[0][1]:.0f} {st[1][0]}={st[1][1]:.0f}  {r[ix['Source']][:80]}")
