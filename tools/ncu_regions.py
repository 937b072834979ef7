"""Debug: aggregate an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) of the megakernel by code region and list the
instructions with the most stall samples. usage: python tools/ncu_regions.py SRC.csv [lo_hex hi_hex]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Address"); isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
names = [hdr[i].replace("stall_", "") for i in stall_cols]
recs, base = [], None
for r in data:
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if base is None:
        base = a
    recs.append((a - base, r[isrc], int(r[isamp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i in stall_cols]))
tot = sum(x[2] for x in recs)
print("total samples", tot, "instructions", len(recs))
# split into functions at RET / EXIT boundaries
bounds = [0]
for i, x in enumerate(recs):
    if x[1].strip().startswith("RET.REL.NODEC") or x[1].strip().startswith("EXIT"):
        bounds.append(recs[i + 1][0] if i + 1 < len(recs) else x[0] + 16)
bounds = sorted(set(bounds))
for lo, hi in zip(bounds, bounds[1:] + [1 << 40]):
    sel = [x for x in recs if lo <= x[0] < hi]
    s = sum(x[2] for x in sel)
    if s < tot * 0.005:
        continue
    st = [sum(x[4][k] for x in sel) for k in range(len(names))]
    top = sorted(zip(st, names), reverse=True)[:6]
    print(f"[{lo:7x},{hi if hi < (1 << 39) else 0:7x}) {len(sel) * 16 / 1024:6.1f} KB samples {s:7d} ({100 * s / tot:4.1f}%) warp-instr {sum(x[3] for x in sel):9d} ", [(n, v) for v, n in top])
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    sel = [x for x in recs if lo <= x[0] < hi]
    print("top instructions by samples:")
    for x in sorted(sel, key=lambda x: -x[2])[:int(sys.argv[4]) if len(sys.argv) > 4 else 50]:
        top = sorted(zip(x[4], names), reverse=True)[:2]
        print(f"  {x[0]:7x} {x[2]:6d} ex {x[3]:8d}  {x[1][:72]:72s} {[(n, v) for v, n in top]}")
