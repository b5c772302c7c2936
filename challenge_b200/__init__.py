"""challenge_b200 -- B200-native (sm_100a) preprocessing hot path of IRIS-AUDIO/challenge.

Drop-in mirrors of the reference's ``pipeline.py`` / ``transforms.py`` / ``data_utils.py``
/ ``metrics.py`` function surface over the C ABI of ``libiris.so`` (include/iris.h).
"""
from .errors import InvalidArgumentError, IrisError  # noqa: F401

__all__ = ['InvalidArgumentError', 'IrisError']
