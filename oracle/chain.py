"""The composed hot path on the CPU oracle (TEST INFRASTRUCTURE ONLY): waveform banks ->
``load_wav`` per source (offline in the reference) -> ``make_pipeline`` element ->
``sj_train.make_dataset`` stage chain (sj_train.py:74-130), with explicit draws.

Used by tests/ as the parity checker and by bench.py's ``cpu_baseline`` /
``--impl reference`` legs as the timed CPU port.  Never on the product path.
"""
import numpy as np

from . import data_utils as D
from . import pipeline as P
from . import transforms as TR


class OracleBank:
    """What the reference keeps in its ``*.pickle`` banks: ``load_wav`` of every source
    (normalize + STFT + ``[F,T,2C]`` layout, data_utils.py:9-29)."""

    def __init__(self, waveforms, normalize=True):
        self.specs = [D.load_wav_array(w, do_normalize=normalize) for w in waveforms]

    def __len__(self):
        return len(self.specs)

    def __getitem__(self, i):
        return self.specs[i]

    def activity(self, i):
        """pipeline.py:55 -- frame active iff any coefficient > 0."""
        return (self.specs[i].max(axis=(0, 2)) > 0).astype(np.uint8)


def clip_draws(d, b):
    """Per-clip draw dict for :func:`oracle.pipeline.merge_complex_specs` from batch arrays."""
    out = {'bg_offset': int(d.bg_offset[b])}
    if d.max_voices > 0:
        out['n_voices'] = int(d.n_voices[b])
        out['voice_u'] = d.voice_u[b]
        if d.voice_gain is not None:    # the host's pow(10., -u) (libm powf in iris_draw_batch): a draw like the others
            out['voice_gain'] = d.voice_gain[b]
        out['voice_offset'] = d.voice_offset[b]
    if d.max_noises > 0:
        out['n_noises'] = int(d.n_noises[b])
        out['noise_u'] = d.noise_u[b]
        if d.noise_gain is not None:
            out['noise_gain'] = d.noise_gain[b]
        out['noise_offset'] = d.noise_offset[b]
    return out


def synth_clip(bg, voices, voice_labels, noises, d, b, n_classes=3, seperate_noise_voice=False,
               debug=None):
    """One ``make_pipeline`` element (pipeline.py:142-174) -> ``(spec[F,T,2C], label[V,T,K])``."""
    T = d.n_frame
    background = bg[int(d.bg_id[b])]
    if d.max_voices == 0:
        # no voice stream (BASELINE config 1): tile + crop only (pipeline.py:29-35)
        bg_frame = background.shape[1]
        tiled = np.tile(background, [1, (T + bg_frame - 1) // bg_frame, 1])
        o = int(d.bg_offset[b])
        return tiled[:, o:o + T].copy(), None
    vids = [int(i) for i in d.voice_id[b]]
    vgroup = P.padded_batch([voices[i] for i in vids])
    lgroup = np.stack([np.asarray(voice_labels[i], np.float32) for i in vids])
    ngroup = None
    if d.max_noises > 0:
        ngroup = P.padded_batch([noises[int(i)] for i in d.noise_id[b]])
    return P.merge_complex_specs(background, (vgroup, lgroup), ngroup, n_frame=T,
                                 n_classes=n_classes, min_ratio=d.min_ratio,
                                 min_noise_ratio=d.min_noise_ratio,
                                 seperate_noise_voice=seperate_noise_voice,
                                 draws=clip_draws(d, b), debug=debug)


def dataset_batch(bg, voices, voice_labels, noises, d, n_classes=3, mode='logmel_minmax',
                  n_mels=80, remap=None, n_out_chan=0, stft_filter=0, clips=None,
                  mel_fn=None):
    """``sj_train.make_dataset`` (sj_train.py:92-130) for one batch of draws ``d``.

    mode: 'complex' | 'magphase' | 'log_magphase' | 'mel' | 'logmel' | 'logmel_minmax'.
    Returns ``(x[B,...], frame_labels[B,T,K] or None, labels_vtk list, keep list)``.
    """
    clips = range(d.batch) if clips is None else clips
    xs, ys, vtks, keep = [], [], [], []
    for b in clips:
        dbg = {}
        spec, label = synth_clip(bg, voices, voice_labels, noises, d, b, n_classes, debug=dbg)
        keep.append(np.asarray(dbg.get('no_overlap', []), np.uint8))
        y = None
        if label is not None:
            vtks.append(label)
            spec, y = D.to_frame_labels(spec, label)                      # sj_train.py:107
        if d.time_masks is not None or d.freq_masks is not None:           # augment (109)
            if d.time_masks is not None:
                spec = TR.mask(spec, axis=-2, max_mask_size=None, n_mask=d.time_masks.shape[1],
                               draws=d.time_masks[b])
            if d.freq_masks is not None:
                spec = TR.mask(spec, axis=-3, max_mask_size=None, n_mask=d.freq_masks.shape[1],
                               draws=d.freq_masks[b])
        if remap == 'stereo_mono':                                         # (110-115)
            spec = D.stereo_mono(spec)
        elif remap == 'merge_aug':
            spec = D.random_merge_aug(n_out_chan)(spec, factor=d.merge_factor[b])
        if stft_filter:
            spec = D.stft_filter(stft_filter)(spec)                        # (116-117)
        xs.append(spec)
        ys.append(y)
    x = np.stack(xs).astype(np.float32)                                    # batch (118)
    y = np.stack(ys) if ys and ys[0] is not None else None
    if mode != 'complex':
        x = TR.complex_to_magphase(x)                                      # (119)
        if mode == 'log_magphase':
            x = TR.log_magphase(x, n_chan=x.shape[-1] // 2)
        elif mode != 'magphase':
            mel = mel_fn or TR.magphase_to_mel(n_mels)
            x = mel(x)                                                     # (120)
            if mode == 'logmel_minmax':
                x = D.minmax(x)                                            # (121-122)
            if mode != 'mel':
                x = D.log_on_mel(x)                                        # (123)
    return x, y, vtks, keep
