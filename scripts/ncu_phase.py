#!/usr/bin/env python
"""Executed warp instructions per kernel phase (source line ranges of k_fused.cu / headers)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
files = {n: [l.strip() for l in open('challenge_b200/csrc/' + n)] for n in ('k_fused.cu', 'fftcore.cuh', 'iris_common.cuh')}
def which(line, text):
    for n, ls in files.items():
        if 0 < line <= len(ls) and ls[line - 1][:40] == text[:40]: return n
    return 'other'
# phases of k_fused.cu by marker comments
marks = []
for i, l in enumerate(files['k_fused.cu'], 1):
    m = re.search(r'// ---- (.*?) ----|^// (KB: number|min-max \+ log)', l)
    if m: marks.append((i, (m.group(1) or m.group(2))[:40]))
def phase(line):
    p = 'head'
    for i, n in marks:
        if line >= i: p = n
    return p
hdr = None; cur = None; agg = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0]:
        ln = int(r[0]); f = which(ln, r[1].strip())
        cur = (f + ':' + phase(ln)) if f == 'k_fused.cu' else f
        continue
    if r[2] in ('...', ''): continue
    d = dict(zip(hdr[4:], r[4:]))
    try: n = int(d['Instructions Executed'])
    except: continue
    toks = r[3].split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    cls = 'fp' if op in ('FADD', 'FFMA', 'FMUL') else ('lds/sts' if op in ('LDS', 'STS') else ('local' if op in ('LDL', 'STL') else 'other'))
    agg[cur][cls] += n; agg[cur]['_'] += n
tot = sum(a['_'] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['_']):
    print('%-60s %5.1f%%  fp %5.1f%%  lds/sts %4.1f%%  local %4.1f%%  other %5.1f%%' % (k, 100 * a['_'] / tot, 100 * a['fp'] / tot, 100 * a['lds/sts'] / tot, 100 * a['local'] / tot, 100 * a['other'] / tot))
