#!/usr/bin/env python
"""Writes tests/golden/mel_matrix_80.json: the mel weight matrix of transforms.py:55-56
(tf.signal.linear_to_mel_weight_matrix(80, 257, 16000): HTK mel scale 1127 ln(1 + f / 700), 125 Hz ..
3800 Hz, triangles in the mel domain, DC row zero) built INDEPENDENTLY of oracle/ and of the product,
from the published formula in float64 with exact rational bin frequencies.  It pins the support
(nnz = 231, rows 5..121), every column sum, every row sum and every entry of 12 columns."""
import json
import math
import os

N_MEL, N_BINS, SR, LO, HI = 80, 257, 16000, 125.0, 3800.0


def mel(f):
    return 1127.0 * math.log1p(f / 700.0)


edges = [mel(LO) + (mel(HI) - mel(LO)) * i / (N_MEL + 1) for i in range(N_MEL + 2)]
W = [[0.0] * N_MEL for _ in range(N_BINS)]
for f in range(1, N_BINS):                      # bin 0 (DC) is a zero row
    m = mel(f * (SR / 2.0) / (N_BINS - 1))
    for j in range(N_MEL):
        lo, ce, up = edges[j], edges[j + 1], edges[j + 2]
        W[f][j] = max(0.0, min((m - lo) / (ce - lo), (up - m) / (up - ce)))
nnz = sum(1 for r in W for v in r if v > 0)
rows = [f for f in range(N_BINS) if any(v > 0 for v in W[f])]
entries = [[f, j, W[f][j]] for j in (0, 1, 5, 10, 20, 30, 40, 50, 60, 70, 75, 79) for f in range(N_BINS) if W[f][j] > 0]
out = {
    'generator': 'scripts/make_mel_golden.py (float64, independent of oracle/ and challenge_b200/)',
    'reference': 'transforms.py:55-56 -> tf.signal.linear_to_mel_weight_matrix(80, 257, 16000) [TF 2.2 defaults 125 .. 3800 Hz]',
    'shape': [N_BINS, N_MEL], 'nnz': nnz, 'first_row': rows[0], 'last_row': rows[-1],
    'max_taps_per_row': max(sum(1 for v in r if v > 0) for r in W),
    'taps_per_column': [sum(1 for f in range(N_BINS) if W[f][j] > 0) for j in range(N_MEL)],
    'column_sums': [sum(W[f][j] for f in range(N_BINS)) for j in range(N_MEL)],
    'row_sums': [sum(W[f]) for f in range(N_BINS)],
    'entries': entries,
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'mel_matrix_80.json')
json.dump(out, open(path, 'w'), indent=1)
print(path, 'nnz', nnz, 'rows', rows[0], rows[-1])
