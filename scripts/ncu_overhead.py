#!/usr/bin/env python
"""Non-FP, non-memory ("overhead") warp instructions per source line from an
`ncu --page source --csv --print-source cuda,sass` dump, scaled to instructions per frame with
the FFMA2 count of the FFT (pass 1 + inter-pass twiddles + pass 2 + epilogue).
usage: ncu_overhead.py src_page.csv [top_n] [frames]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 160256
rows = list(csv.reader(open(path)))
hdr = None; cur = None
agg = collections.defaultdict(collections.Counter)
txt = {None: '(no line info)'}
for r in rows:
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0]:
        cur = int(r[0]); txt[cur] = r[1].strip(); continue
    if r[2] in ('...', ''): continue
    d = dict(zip(hdr[4:], r[4:]))
    try: n = int(d['Instructions Executed'])
    except (KeyError, ValueError): continue
    toks = r[3].split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    agg[cur][op.split('.')[0]] += n
tot = sum(sum(a.values()) for a in agg.values())
ops = collections.Counter()
for a in agg.values(): ops.update(a)
sc = 1.0 / frames
print('source-page units per frame: %.0f' % (tot * sc))
print(' '.join('%s:%.0f' % (o, c * sc) for o, c in ops.most_common(40)))
skip = {'FFMA2', 'FADD2', 'FMUL2', 'LDS', 'STS', 'SHFL', 'STG', 'MUFU'}
res = []
for k, a in agg.items():
    s = sum(c for o, c in a.items() if o not in skip)
    res.append((s * sc, k, a))
res.sort(key=lambda x: -x[0])
print('overhead total %.0f (%.1f %% of all)' % (sum(r[0] for r in res), 100 * sum(r[0] for r in res) / (tot * sc)))
for s, k, a in res[:top]:
    print('%6.1f  L%-5s %-62s %s' % (s, k, txt[k][:62], ' '.join('%s:%.1f' % (o, c * sc) for o, c in a.most_common(8) if o not in skip)))
