"""Drop-in mirror of the data-path functions of the reference's ``trainer.py`` (the legacy
trainer): ``minmax_log_on_mel`` (61-76), ``augment`` (80-83), ``preprocess_labels`` (86-94),
``to_density_labels`` (97-104) and ``make_dataset`` (107-141).  SURVEY.md 8f rank 4."""
import os

import numpy as np

from . import _lib as L
from . import _ops as O
from .data_utils import augment, log_on_mel, minmax   # noqa: F401  (augment: trainer.py:80-83 == data_utils.py:58-61)
from .pipeline import AUTOTUNE, make_pipeline
from .transforms import complex_to_magphase, magphase_to_mel
from .utils import load_data


def minmax_log_on_mel(mel, labels=None):
    """trainer.py:61-76 -- per-sample min-max (safe_div) then ``log(x + 1e-8)``; after
    ``magphase_to_mel`` in a dataset chain the pair is lowered into the fused launch."""
    out = log_on_mel(minmax(mel))
    if labels is not None:
        return out, labels
    return out


minmax_log_on_mel._iris_stage = ('minmax_log',)


def preprocess_labels(multiplier):
    """trainer.py:86-94 -- five ``avg_pool1d(2, 2, 'SAME') * 2`` stages (sum pooling by 32 with
    TF's valid-cell average: a lone last cell is doubled), then ``* multiplier``."""
    def _preprocess(x, y):
        t = O.dev(y)
        squeeze = t.dim() == 2
        if squeeze:
            t = t[None]
        for i in range(5):
            B, T, K = (int(s) for s in t.shape)
            out = O.empty((B, (T + 1) // 2, K))
            O.call('iris_op_sum_pool2', O.ptr(t), O.ptr(out), B, T, K, float(multiplier) if i == 4 else 1.0)
            t = out
        return x, (t[0] if squeeze else t)
    return _preprocess


def to_density_labels(x, y):
    """
    :param y: [..., n_voices, n_frames, n_classes]
    :return: [..., n_frames, n_classes]
    (trainer.py:97-104)
    """
    t = O.dev(y)
    shape = tuple(t.shape)
    outer = int(np.prod(shape[:-3], dtype=np.int64))
    V, inner = int(shape[-3]), int(shape[-2] * shape[-1])
    out = O.empty(shape[:-3] + shape[-2:])
    O.call('iris_op_density_labels', O.ptr(t), O.ptr(out), outer, V, inner)
    return x, out


to_density_labels._iris_stage = ('density_labels',)


def make_dataset(config, training=True, n_classes=3):
    """trainer.py:107-141 with the banks read by ``utils.load_data`` (spectrogram pickles)."""
    if not os.path.exists(config.datapath):
        config.datapath = ''
    if training:
        backgrounds = load_data(os.path.join(config.datapath, config.background_sounds))
        voices = load_data(os.path.join(config.datapath, config.voices))
        labels = load_data(os.path.join(config.datapath, config.labels))
    else:
        backgrounds = load_data(os.path.join(config.datapath, config.test_background_sounds))
        voices = load_data(os.path.join(config.datapath, config.test_voices))
        labels = load_data(os.path.join(config.datapath, config.test_labels))
    if labels.max() - 1 != config.n_classes:
        labels //= 10
    labels = np.eye(n_classes, dtype='float32')[labels]   # to one-hot vectors
    noises = load_data(os.path.join(config.datapath, config.noises))
    pipeline = make_pipeline(backgrounds, voices, labels, noises, n_frame=config.n_frame,
                             max_voices=config.max_voices, max_noises=config.max_noises,
                             n_classes=n_classes, snr=config.snr, min_ratio=1)
    pipeline = pipeline.map(to_density_labels)
    if training:
        pipeline = pipeline.map(augment)
    pipeline = pipeline.batch(config.batch_size, drop_remainder=False)
    pipeline = pipeline.map(complex_to_magphase)
    pipeline = pipeline.map(magphase_to_mel(config.n_mels))
    pipeline = pipeline.map(minmax_log_on_mel)
    pipeline = pipeline.map(preprocess_labels(config.multiplier))
    return pipeline.prefetch(AUTOTUNE)
