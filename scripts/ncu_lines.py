#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py src_page.csv [top_n]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])  # samples, instr, excess wavefronts, stalls
src_text = {}
line_key = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Name':
        cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0]:   # a CUDA source line
        line_key = (cur_file, int(r[0])); src_text[line_key] = r[1].strip(); continue
    if r[2] in ('...', ''): continue
    d = dict(zip(hdr[4:], r[4:]))
    def num(k):
        try: return int(d.get(k, '0').replace(',', ''))
        except ValueError: return 0
    a = agg[line_key or ('?', 0)]
    a[0] += num('# Samples'); a[1] += num('Instructions Executed'); a[2] += num('L1 Wavefronts Shared Excessive')
    for k in hdr:
        if k.startswith('stall_') and 'Not Issued' not in k:
            v = num(k)
            if v: a[3][k[6:]] += v
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print('total samples %d, total warp-instr %d' % (tot_s, tot_i))
print('--- top lines by stall samples')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join('%s:%d' % x for x in a[3].most_common(3))
    print('%5.1f%% smp %5.1f%% ins  xwf %8d  %s:%s  %-70s | %s' % (100*a[0]/tot_s, 100*a[1]/tot_i, a[2], k[0], k[1], (src_text.get(k) or '')[:70], st))
print('--- top lines by instructions')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5.1f%% ins %5.1f%% smp  %s:%s  %s' % (100*a[1]/tot_i, 100*a[0]/tot_s, k[0], k[1], (src_text.get(k) or '')[:90]))
tot = collections.Counter()
for a in agg.values():
    tot.update(a[3])
s = sum(tot.values()) or 1
print('--- stall reasons overall:', ' '.join('%s:%.1f%%' % (k, 100 * v / s) for k, v in tot.most_common(12)))
