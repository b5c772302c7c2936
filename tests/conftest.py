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


def rel_err_gated(a, b, gate=1e-3):
    """Element-wise relative error max|a-b| / |b| over the cells with |b| > gate * max|b| (below the
    gate an absolute fp32 error of ~3e-7 max|b| is more than gate^-1 * 3e-7 relative: ill-posed)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    sel = np.abs(b) > gate * np.abs(b).max()
    return float((np.abs(a - b)[sel] / np.abs(b)[sel]).max()) if sel.any() else 0.0


def phase_report(mag_ref, ph, ph_ref):
    """The phase figures SURVEY.md 8d asks for.  An absolute error e of re / im (fp32 FFT rounding,
    ~1e-6 of max|X| on both sides) turns into a phase error e / |X|, so the circular error is
    reported by magnitude gate, together with the magnitude-weighted error |dphi| |X| / max|X| over
    ALL cells (the complex-domain error the phase error stands for), the branch-cut sign mismatches
    (ref and got on opposite sides of +-pi: |raw difference| > pi) among the gated cells, and the
    cells of exactly zero magnitude (masked: atan2(+-0, +-0) is decided by sign bits alone) whose
    phase is not bit-identical."""
    mag_ref = np.asarray(mag_ref, np.float64)
    ph = np.asarray(ph, np.float64)
    ph_ref = np.asarray(ph_ref, np.float64)
    raw = ph - ph_ref
    d = np.abs(np.angle(np.exp(1j * raw)))
    mx = mag_ref.max()
    out = {}
    for name, gate in (('gate1e-3', 1e-3), ('gate1e-2', 1e-2), ('gate1e-1', 1e-1)):
        sel = mag_ref > gate * mx
        out[name] = float(d[sel].max()) if sel.any() else 0.0
    out['weighted'] = float((d * mag_ref).max() / mx)
    sel = mag_ref > 1e-3 * mx
    out['cut_flips_gated'] = int((np.abs(raw[sel]) > np.pi).sum())
    out['cut_flips_all'] = int((np.abs(raw) > np.pi).sum())
    zero = mag_ref == 0
    out['zero_cells'] = int(zero.sum())
    out['zero_cells_differ'] = int(((ph[zero] != ph_ref[zero]) | (np.signbit(ph[zero]) != np.signbit(ph_ref[zero]))).sum())
    return out


def logmag_report(lg, lg_ref):
    """log(|X| + 1e-8) features: absolute error of the log by magnitude gate (an absolute error e of
    |X| is e / |X| in the log) and the normalised max error back in the linear domain."""
    lg = np.asarray(lg, np.float64)
    lg_ref = np.asarray(lg_ref, np.float64)
    m, m_ref = np.exp(lg), np.exp(lg_ref)
    out = {}
    for name, gate in (('gate1e-3', 1e-3), ('gate1e-2', 1e-2), ('gate1e-1', 1e-1)):
        sel = m_ref > gate * m_ref.max()
        out[name] = float(np.abs(lg[sel] - lg_ref[sel]).max()) if sel.any() else 0.0
    out['linear_nmax'] = float(np.abs(m - m_ref).max() / m_ref.max())
    return out


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
