import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def nmax_err(a, b):
    """Normalised max error max|a-b| / max|b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def phase_err(mag_ref, ph, ph_ref, gate=1e-3):
    """Circular phase error gated on magnitude > gate * max(mag)."""
    d = np.angle(np.exp(1j * (np.asarray(ph, np.float64) - np.asarray(ph_ref, np.float64))))
    sel = mag_ref > gate * mag_ref.max()
    return float(np.abs(d[sel]).max()) if sel.any() else 0.0


@pytest.fixture(scope='session')
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from challenge_b200.engine import Engine
    eng = Engine(0)
    eng.set_mel(80)
    yield eng
    eng.close()


class Workload:
    """Small synthetic banks registered on the GPU and mirrored on the oracle."""

    def __init__(self, engine, n_chan, seed, n_bg=4, n_voice=24, n_noise=6, bg_seconds=10.0):
        from challenge_b200 import _lib as L
        from challenge_b200.synth import synthetic_banks
        from oracle.chain import OracleBank
        self.bgs, self.voices, self.labels, self.noises = synthetic_banks(
            seed, n_chan, n_bg=n_bg, n_voice=n_voice, n_noise=n_noise, bg_seconds=bg_seconds)
        self.eng = engine
        self.bg_frames = engine.register_bank(L.BANK_BG, self.bgs)
        self.voice_frames = engine.register_bank(L.BANK_VOICE, self.voices, labels=self.labels)
        self.noise_frames = engine.register_bank(L.BANK_NOISE, self.noises)
        self.o_bg = OracleBank(self.bgs)
        self.o_voice = OracleBank(self.voices)
        self.o_noise = OracleBank(self.noises)


_workloads = {}


@pytest.fixture(scope='session')
def workload_factory(engine):
    def make(n_chan, seed=20202, **kw):
        key = (n_chan, seed, tuple(sorted(kw.items())))
        if key not in _workloads:
            _workloads[key] = Workload(engine, n_chan, seed, **kw)
        else:
            # re-register: the engine holds one bank set at a time
            w = _workloads[key]
            from challenge_b200 import _lib as L
            engine.register_bank(L.BANK_BG, w.bgs)
            engine.register_bank(L.BANK_VOICE, w.voices, labels=w.labels)
            engine.register_bank(L.BANK_NOISE, w.noises)
        return _workloads[key]
    return make
