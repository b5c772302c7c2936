"""CPU oracle for the IRIS-AUDIO/challenge preprocessing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``challenge_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the product path.

What it is
----------
A line-by-line numpy (fp32) restatement of the reference's algorithm for

* ``pipeline.py``      (sample synthesis: tile/crop, voice/noise mix, labels),
* ``transforms.py``    (masks, mag/phase, mel, log / min-max scalings),
* ``data_utils.py``    (waveform ingest / STFT, min-max, log, label ops, remaps),
* ``metrics.py``       (event error-rate counts, micro-F1 counts, cos_sim),

each function citing the reference ``file:line`` it follows.  The reference
itself cannot be imported here: every hot-path module imports TensorFlow at top
level (pipeline.py:1, transforms.py:2, data_utils.py:3, metrics.py:5) and
TensorFlow / tensorflow_addons are absent from this image.  The one reference
call that *is* runnable -- ``torchaudio.transforms.Spectrogram(512, power=None)``
(data_utils.py:17) -- is called as-is by :func:`oracle.data_utils.stft`.

All randomness is explicit: every function that draws from ``tf.random`` in the
reference takes the draws as arguments here, in the reference's draw order, so
that the CUDA path and the oracle consume identical randomness.

Pinning status (see DESIGN.md "Oracle"):

* PINNED by the reference's own known-answer tests (tests/golden/):
  ``er_score`` (metrics_test.py:12-25), ``complex_to_magphase`` /
  ``magphase_to_complex`` (transforms_test.py:79-96), ``log_magphase`` (57-62),
  ``mask`` patterns (10-32), ``random_shift`` pattern (34-43),
  ``minmax_norm_magphase`` property (64-77), output shapes of
  ``merge_complex_specs`` / ``make_pipeline`` / ``magphase_to_mel``.
* PARITY UNPINNED (no reference test or fixture holds values; the restatement
  follows the cited lines and documented TF-2.2 / tfa semantics, cross-checked
  by an independent float64 computation): STFT values, the mel weight matrix
  and mel values, numeric values of ``merge_complex_specs``, ``minmax``,
  ``log_on_mel``, ``to_frame_labels``, ``label_downsample``, ``stereo_mono``,
  ``random_merge_aug``, ``stft_filter``, ``f1_score``, ``cos_sim``.
"""

EPSILON = 1e-8  # utils.py:6 / transforms.py:7
