"""Shared-memory wavefronts per instruction of an `ncu --page source --csv` export, per 64 points. usage: ncu_wavefronts.py <src.csv> [points]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
pts = float(sys.argv[2]) if len(sys.argv) > 2 else 148 * 347200
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]; ci = hdr.index('Instructions Executed'); ai = hdr.index('Address'); si = hdr.index('Source')
wi = hdr.index('L1 Wavefronts Shared'); ii = hdr.index('L1 Wavefronts Shared Ideal'); gi = hdr.index('L1 Tag Requests Global')
base = None; tot = 0; out = []; byop = collections.Counter(); gl = 0
for r in rows[hi + 1:]:
    try: a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai]); v = int(r[ci])
    except Exception: continue
    if base is None: base = a
    try: g = float(r[gi]); gl += g
    except Exception: pass
    try: w = float(r[wi]); idl = float(r[ii])
    except Exception: continue
    if w > 0:
        tot += w
        t = r[si].split(); op = t[1] if t[0].startswith('@') else t[0]
        byop[op] += w
        out.append((w, idl, v, a - base, r[si]))
k = 64.0 / pts
print('shared wavefronts per 64 points: %.1f   global tag requests per 64 points: %.1f' % (tot * k, gl * k))
print('by opcode:', ' '.join('%s:%.1f' % (o, w * k) for o, w in byop.most_common(12)))
for w, idl, v, off, s in sorted(out, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f'{off:#06x} wf/64pt={w*k:6.2f} ideal={idl*k:6.2f} exec/64pt={v*k:5.2f} wf/exec={w/max(v,1):5.2f} {s[:60]}')
