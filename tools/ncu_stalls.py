"""Top stall sites of an `ncu --page source --csv` export. usage: ncu_stalls.py <src.csv> [top] [tiles]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
T = float(sys.argv[3]) if len(sys.argv) > 3 else 401450
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]; ci = hdr.index('Instructions Executed'); ai = hdr.index('Address'); si = hdr.index('Source'); ns = hdr.index('# Samples')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None; data = []; tot = 0; agg = {}
for r in rows[hi + 1:]:
    try: a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai]); v = int(r[ci]); s = int(r[ns])
    except Exception: continue
    if base is None: base = a
    st = {hdr[i][6:]: int(r[i]) for i in stall if r[i] not in ('', '0')}
    for k, n in st.items(): agg[k] = agg.get(k, 0) + n
    data.append((a - base, v, s, r[si], st)); tot += s
print('total samples', tot, ' by reason:', ' '.join('%s:%.1f%%' % (k, 100 * n / tot) for k, n in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for off, v, s, src, st in sorted(data, key=lambda d: -d[2])[:top]:
    t = ','.join('%s:%d' % (k, n) for k, n in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f'{off:#06x} {v/T:5.2f} {100*s/tot:5.2f}%  {src[:64]:64s} {t}')
