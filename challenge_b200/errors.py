"""Error types of the drop-in layer, mirroring the reference's error behaviour."""


class IrisError(RuntimeError):
    """A libiris call failed (CUDA error, wrong call order, unsupported shape)."""


class InvalidArgumentError(ValueError):
    """Stand-in for ``tf.errors.InvalidArgumentError``: the reference raises it when an
    integer ``tf.random.uniform`` gets an empty range, e.g. a voice group whose padded
    length equals ``n_frame`` with ``min_ratio=1`` (pipeline.py:68-69)."""
