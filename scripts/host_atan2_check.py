#!/usr/bin/env python
"""Float32 emulation of fast_atan2f_mag_x2 (iris_common.cuh) against numpy arctan2 in float64: error over
random and edge inputs (signed zeros, axes, tiny magnitudes)."""
import numpy as np
f = np.float32
C = [0.0028662257, -0.0161657367, 0.0429096138, -0.0752896400, 0.1065626393, -0.1420889944, 0.1999355085, -0.3333314528, 1.0]


def atan2_mag(y, x):
    y, x = f(y), f(x)
    m = np.sqrt(f(x * x + y * y), dtype=f)
    d = np.maximum(np.maximum(f(m + np.abs(x)), np.abs(y)), f(1e-30))
    t = f(y / d)
    s = f(t * t)
    r = f(2 * C[0])
    for c in C[1:]:
        r = f(r * s + f(2 * c))
    r = f(r * t)
    neg = np.signbit(x)
    return np.where(neg, f(np.copysign(f(np.pi), y) - r), r).astype(f)


rng = np.random.default_rng(0)
x = rng.standard_normal(2_000_000).astype(f) * f(10) ** rng.integers(-6, 3, 2_000_000).astype(f)
y = rng.standard_normal(2_000_000).astype(f) * f(10) ** rng.integers(-6, 3, 2_000_000).astype(f)
got = atan2_mag(y, x).astype(np.float64)
ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
err = np.abs(np.angle(np.exp(1j * (got - ref))))
print('random: max circular error %.3e rad' % err.max())
edge = [(0.0, 1.0), (-0.0, 1.0), (0.0, -1.0), (-0.0, -1.0), (0.0, 0.0), (-0.0, 0.0), (0.0, -0.0), (-0.0, -0.0), (1.0, 0.0),
        (-1.0, 0.0), (1.0, -0.0), (-1.0, -0.0), (1e-25, -1e-25), (1e-22, 1e-30), (3.0, -3.0), (-3.0, -3.0)]
for yy, xx in edge:
    g = float(atan2_mag(np.array([yy], f), np.array([xx], f))[0])
    r = float(np.arctan2(f(yy), f(xx)))
    ok = (abs(g - r) < 2e-6) and (np.signbit(g) == np.signbit(r))
    print('atan2(%g, %g) = %.7f  numpy %.7f %s' % (yy, xx, g, r, 'ok' if ok else 'DIFFERS'))
