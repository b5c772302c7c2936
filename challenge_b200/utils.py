"""The slice of the reference's ``utils.py`` the data path touches: the bank file format
(``load_data``, utils.py:88-94), ``EPSILON`` (6), ``label_downsample_model`` (7) and
``safe_div`` (114-116).  The model / optimizer helpers of that file are out of scope."""
import pickle

import numpy as np

EPSILON = 1e-8
label_downsample_model = (3, 6, 7, 8, 9)


def load_data(path):
    """utils.py:88-94 -- banks are ``.pickle`` files holding a list of complex spectrograms
    ``[257, t, chan*2]`` (what ``data_utils.load_wav`` returns), labels ``.npy`` integer arrays.
    The lists go straight into :func:`challenge_b200.pipeline.make_pipeline`."""
    if path.endswith('.pickle'):
        return pickle.load(open(path, 'rb'))
    elif path.endswith('.npy'):
        return np.load(path)
    else:
        raise ValueError('invalid file format')


def save_bank(path, specs):
    """Write a bank in the reference's format (the inverse of :func:`load_data` for ``.pickle``):
    a list of float32 arrays ``[257, t, chan*2]``, e.g. ``load_wav`` of every file."""
    items = [np.ascontiguousarray(np.asarray(getattr(s, 'cpu', lambda: s)()), np.float32) for s in specs]
    with open(path, 'wb') as f:
        pickle.dump(items, f, protocol=pickle.HIGHEST_PROTOCOL)


def safe_div(x, y, eps=EPSILON):
    """utils.py:114-116 -- ``x / max(y, eps)`` on torch tensors."""
    import torch
    return x / torch.clamp(y, min=eps)
