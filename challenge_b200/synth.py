"""Synthetic waveform banks of the benchmark workloads (SURVEY.md 8d / BASELINE.md 3)."""
import numpy as np


def synthetic_banks(seed, n_chan, n_bg=64, n_voice=256, n_noise=64, sr=16000, bg_seconds=10.0,
                    lo_s=0.5, hi_s=4.0, n_classes=3):
    """The synthetic banks of SURVEY.md 8(d): N(0,1)*0.1 waveforms; voices get a trailing
    0-25 % of exact zeros (exercises the ``> 0`` activity mask like pipeline_test.py:21-24)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f32 = np.float32
    bgs = [(rng.standard_normal((n_chan, int(bg_seconds * sr)), dtype=f32) * f32(0.1))
           for _ in range(n_bg)]
    voices, labels = [], []
    for _ in range(n_voice):
        n = int(rng.integers(int(lo_s * sr), int(hi_s * sr) + 1))
        w = rng.standard_normal((n_chan, n), dtype=f32) * f32(0.1)
        z = int(n * rng.uniform(0, 0.25))
        if z:
            w[:, n - z:] = 0
        voices.append(w)
        labels.append(int(rng.integers(n_classes)))
    labels = np.eye(n_classes, dtype=f32)[labels]
    noises = []
    for _ in range(n_noise):
        n = int(rng.integers(int(lo_s * sr), int(hi_s * sr) + 1))
        noises.append(rng.standard_normal((n_chan, n), dtype=f32) * f32(0.1))
    return bgs, voices, labels, noises
