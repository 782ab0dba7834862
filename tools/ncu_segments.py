"""Per-address-range instruction counts of an `ncu --page source --csv` export, normalised per warp tile.
usage: ncu_segments.py <src.csv> <tiles> [min_instr_per_tile]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
T = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
ci, ai, si = hdr.index('Instructions Executed'), hdr.index('Address'), hdr.index('Source')
data = []
for r in rows[hi + 1:]:
    try:
        a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai]); v = int(r[ci])
    except Exception:
        continue
    data.append((a, v, r[si]))
base = data[0][0]
segs = []; prev = None
for a, v, s in data:
    off = a - base
    if prev is None or v != prev:
        if prev is not None: segs.append((s0, off, prev, sm))
        s0 = off; sm = 0; prev = v
    sm += v
segs.append((s0, off + 16, prev, sm))
tot = 0
for s0, s1, c, sm in segs:
    tot += sm
    if sm / T >= thr: print(f"{s0:#06x}-{s1:#06x} n={(s1-s0)//16:4d} exec/tile={c/T:8.4f} instr/tile={sm/T:7.2f}")
print("total instr/tile", tot / T)
