"""Oracle restatement of the reference's ``pipeline.py`` (TEST INFRASTRUCTURE ONLY).

The mix is done in the SPECTROGRAM domain exactly as the reference does, so the
CUDA path's time-domain-sum shortcut (one FFT of the gained sum) is genuinely
cross-checked.  PARITY UNPINNED for values (pipeline_test.py checks shapes only).
"""
import numpy as np


class InvalidArgumentError(ValueError):
    """Stand-in for TF's error on an empty integer-uniform range (pipeline.py:68-69)."""


def merge_complex_specs(background, voices_and_labels, noises=None, n_frame=300,
                        n_classes=3, t_axis=1, min_ratio=2 / 3, min_noise_ratio=1 / 2,
                        snr=-20, seperate_noise_voice=False, *, draws, debug=None):
    """pipeline.py:6-110 with explicit randomness.

    ``draws`` (dict), in the reference's draw order:
      bg_offset      int, crop offset in [0, tiledT - n_frame]          (line 35)
      n_voices       int in [1, max_voices) (1 if max_voices == 1)       (41-46)
      voice_u[v]     fp32 in [0, -snr/10); gain = pow(10f, -u)          (50)
      voice_offset[v] int in [0, len - n_frame)                          (68-69)
      n_noises       int in [0, max_noises)                              (87-88)
      noise_u[n]     fp32 in [0, 2); gain = pow(10f, -u)                (94)
      noise_offset[n] int in [0, len - n_frame]  (random_crop)           (103)
      voice_gain / noise_gain (optional): the host's fp32 ``pow(10., -u)`` handed over as a draw
      (libm ``powf`` and numpy's ``power`` differ in the last bit for ~1 % of the inputs)
    """
    f32 = np.float32
    assert t_axis == 1
    voices, labels = voices_and_labels
    background = np.asarray(background, f32)
    voices = np.asarray(voices, f32)
    labels = np.asarray(labels, f32)

    # background: tile + random crop (29-35)
    bg_frame = background.shape[1]
    reps = (n_frame + bg_frame - 1) // bg_frame
    tiled = np.tile(background, [1, reps, 1])
    o_b = int(draws['bg_offset'])
    assert 0 <= o_b <= tiled.shape[1] - n_frame
    complex_spec = tiled[:, o_b:o_b + n_frame].copy()

    only_voice = np.zeros_like(complex_spec)
    only_noise = complex_spec.copy()

    # voices (41-84)
    max_voices = voices.shape[0]
    n_voices = int(draws['n_voices']) if max_voices > 1 else 1
    if max_voices > 1:
        assert 1 <= n_voices < max_voices
    label = np.zeros([max_voices, n_frame, n_classes], f32)
    for v in range(n_voices):
        voice = voices[v]
        if draws.get('voice_gain') is not None:   # the host's pow(10., -u), passed explicitly like every draw
            v_ratio = f32(draws['voice_gain'][v])
        else:
            u = f32(draws['voice_u'][v])
            v_ratio = np.power(f32(10.), -u, dtype=f32)
        v_frame = voice.shape[1]

        l = np.tile(labels[v:v + 1], [v_frame, 1])              # [v_frame, K]
        m = (voice.max(axis=(0, 2)) > 0).astype(f32)             # line 55
        l = l * m[:, None]

        pad_size = n_frame - int(np.int32(f32(min_ratio) * f32(v_frame)))   # 58-59
        if pad_size > 0:
            voice = np.pad(voice, [(0, 0), (pad_size, pad_size), (0, 0)])
            l = np.pad(l, [(pad_size, pad_size), (0, 0)])

        maxval = voice.shape[1] - n_frame
        if maxval <= 0:
            raise InvalidArgumentError('Need minval < maxval, got 0 >= %d' % maxval)
        offset = int(draws['voice_offset'][v])
        assert 0 <= offset < maxval
        voice = voice[:, offset:offset + n_frame]
        l = l[offset:offset + n_frame]
        onehot = np.zeros(max_voices, f32)
        onehot[v] = 1
        l = onehot.reshape(-1, 1, 1) * l[None]

        no_overlap = f32(np.max(np.sum(label + l, axis=0)) < 2)   # 78-79
        if debug is not None:
            debug.setdefault('no_overlap', [0] * max_voices)[v] = int(no_overlap)
        complex_spec = complex_spec + v_ratio * voice * no_overlap
        if seperate_noise_voice:
            only_voice = only_voice + v_ratio * voice * no_overlap
        label = label + l * no_overlap

    if noises is not None:
        noises = np.asarray(noises, f32)
        n_noises = int(draws['n_noises'])
        assert 0 <= n_noises < max(noises.shape[0], 1) or n_noises == 0
        for n in range(n_noises):
            noise = noises[n]
            if draws.get('noise_gain') is not None:
                n_ratio = f32(draws['noise_gain'][n])
            else:
                u = f32(draws['noise_u'][n])
                n_ratio = np.power(f32(10.), -u, dtype=f32)
            ns_frame = f32(noise.shape[1])
            pad_size = n_frame - int(np.int32(f32(min_noise_ratio) * ns_frame))
            if pad_size > 0:
                noise = np.pad(noise, [(0, 0), (pad_size, pad_size), (0, 0)])
            off = int(draws['noise_offset'][n])
            assert 0 <= off <= noise.shape[1] - n_frame
            noise = noise[:, off:off + n_frame]
            if seperate_noise_voice:
                only_noise = only_noise + n_ratio * noise
            complex_spec = complex_spec + n_ratio * noise
    if seperate_noise_voice:
        label = (label, only_voice, only_noise)
    return complex_spec, label


def padded_batch(items, t_axis=1):
    """``Dataset.padded_batch`` (pipeline.py:155-156, 165-166): zero-pad every
    member of the group to the group's longest along time."""
    longest = max(x.shape[t_axis] for x in items)
    out = []
    for x in items:
        pad = [(0, 0)] * x.ndim
        pad[t_axis] = (0, longest - x.shape[t_axis])
        out.append(np.pad(x, pad))
    return np.stack(out)


def make_pipeline_element(backgrounds, voices, labels, noises, element, *, bg_stream,
                          voice_stream, noise_stream=None, n_frame=300, max_voices=10,
                          max_noises=10, n_classes=3, draws, **kwargs):
    """One element of ``make_pipeline`` (pipeline.py:113-175): the three shuffled
    streams are given explicitly (``*_stream`` = item ids in stream order);
    element ``e`` zips bg ``e`` with voice group ``[e*V, (e+1)*V)`` and noise
    group ``[e*M, (e+1)*M)`` (padded_batch then zip, lines 150-167)."""
    assert len(backgrounds[0].shape) == 3, 'each spec must be a 3D-tensor'
    assert len(voices) == len(labels)
    assert len(labels[0].shape) == 1 and labels[0].shape[0] == n_classes
    e = element
    bg = backgrounds[bg_stream[e]]
    vids = voice_stream[e * max_voices:(e + 1) * max_voices]
    vgroup = padded_batch([np.asarray(voices[i], np.float32) for i in vids])
    lgroup = np.stack([np.asarray(labels[i], np.float32) for i in vids])
    ngroup = None
    if noises is not None:
        nids = noise_stream[e * max_noises:(e + 1) * max_noises]
        ngroup = padded_batch([np.asarray(noises[i], np.float32) for i in nids])
    return merge_complex_specs(bg, (vgroup, lgroup), ngroup, n_frame=n_frame,
                               n_classes=n_classes, draws=draws, **kwargs)
