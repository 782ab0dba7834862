"""Aggregate an `ncu --page source --csv` (SASS) export per CUDA source line using nvdisasm -g line markers.
usage: ncu_lines.py <src.csv> <nvdisasm -g -c output> <kernel-substring> [top]"""
import csv, re, sys
src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> line map for the chosen function
amap = {}; cur = None; infn = False; fname = None
for ln in open(dis, errors='replace'):
    m = re.match(r'\s*\.text\.(\S+):', ln) or re.match(r'\s*//-+ \.text\.(\S+)', ln)
    if m:
        infn = kern in m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and infn:
        amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
ci, ai, ns = hdr.index('Instructions Executed'), hdr.index('Address'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
per = {}; tot = 0; tots = 0
base = None
for r in rows[hi + 1:]:
    try:
        a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai]); v = int(r[ci]); s = int(r[ns])
    except Exception:
        continue
    if base is None: base = a
    key, _ = amap.get(a - base, ((None, -1), ''))
    d = per.setdefault(key, [0, 0, {}])
    d[0] += v; d[1] += s; tot += v; tots += s
    for c in stall_cols:
        try: x = int(r[c])
        except Exception: x = 0
        if x: d[2][hdr[c]] = d[2].get(hdr[c], 0) + x
print('total warp-instructions', tot, 'samples', tots, 'mapped addrs', len(amap))
srcs = {}
def line_text(key):
    if not key or key[0] is None: return ''
    f, l = key
    import glob
    if f not in srcs:
        c = glob.glob('/root/repo/**/' + f, recursive=True)
        srcs[f] = open(c[0]).read().split('\n') if c else []
    t = srcs[f]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ''
for key, (v, s, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ','.join('%s:%d' % (k[6:], n) for k, n in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print('%5.1f%% inst %5.1f%% smp  %s:%s  %s   [%s]' % (100 * v / tot, 100 * s / max(tots, 1), key[0] if key else None, key[1] if key else -1, line_text(key), tops))
