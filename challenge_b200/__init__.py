"""challenge_b200 -- B200-native (sm_100a) preprocessing hot path of IRIS-AUDIO/challenge.

Drop-in mirrors of the reference's ``pipeline.py`` / ``transforms.py`` / ``data_utils.py``
/ ``metrics.py`` function surface over the C ABI of ``libiris.so`` (include/iris.h):

    from challenge_b200 import pipeline, transforms, data_utils, metrics

There is no CPU fallback: the modules need ``libiris.so`` (``python -m challenge_b200.build``)
and a CUDA device, and fail loudly without them.
"""
from .errors import InvalidArgumentError, IrisError  # noqa: F401


def set_seed(seed):
    """Seed the host generator behind every random draw (``tf.random.set_seed`` stand-in)."""
    from ._ops import set_seed as _s
    _s(seed)


__all__ = ['InvalidArgumentError', 'IrisError', 'set_seed']
