#!/usr/bin/env python
"""Opcode histogram (executed warp instructions) from an ncu source+sass CSV dump."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; ops = collections.Counter(); tot = 0
for r in rows:
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0]: continue
    if r[2] in ('...', ''): continue
    d = dict(zip(hdr[4:], r[4:]))
    try: n = int(d['Instructions Executed'])
    except: continue
    toks = r[3].split()
    op = toks[0]
    if op.startswith('@'): op = toks[1]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL')) and '.' in op else '')
    ops[op] += n; tot += n
print('total', tot)
for k, v in ops.most_common(45): print('%-14s %12d %5.1f%%' % (k, v, 100 * v / tot))
