// Per-bin epilogue shared by the fused waveform kernel (k_fused.cu) and the spectrogram-bank
// mix (k_spec.cu): masks, stft_filter, channel remaps, complex / mag-phase / log-mag-phase output.
#pragma once
#include "iris_common.cuh"

namespace iris {

// One (bin f, frame t, channel pair) of the output: (r0, i0) / (r1, i1) are the complex values of
// channels 2*pair and 2*pair + 1, m the product of the SpecAugment masks of this cell.
template <int MODE>
__device__ __forceinline__ void store_bin(const FusedParams& p, int b, int f, int t, int pair,
                                          bool has1, float r0, float i0, float r1, float i1,
                                          float m) {
    // masks are applied by multiplication (transforms.py:40) so zeros keep their sign
    r0 *= m; i0 *= m; r1 *= m; i1 *= m;
    const float filt = (f >= 1 && f <= p.filter_k) ? 0.f : 1.f;   // data_utils.py:126-136
    if (p.remap == REMAP_NONE) {
        const int C = p.C;
        float* o = p.out + ((size_t(b) * kBins + f) * p.T + t) * size_t(2 * C);
        if (p.filter_k > 0) { r0 *= filt; i0 *= filt; r1 *= filt; i1 *= filt; }
        float a0 = r0, a1 = r1, b0 = i0, b1 = i1;   // first half / second half of the last dim
        if (MODE != FM_COMPLEX) {
            a0 = sqrt_approx(fmaf(r0, r0, i0 * i0));   // transforms.py:116
            b0 = fast_atan2f(i0, r0);                  // transforms.py:117
            a1 = sqrt_approx(fmaf(r1, r1, i1 * i1));
            b1 = fast_atan2f(i1, r1);
            if (MODE == FM_LOGMAGPHASE) {            // transforms.py:80-86
                a0 = __logf(a0 + 1e-8f);   // MUFU.LG2 * ln 2, as in the log-mel epilogue (|error| ~1e-6)
                a1 = __logf(a1 + 1e-8f);
            }
        }
        if (C == 2) {
            *reinterpret_cast<float4*>(o) = make_float4(a0, a1, b0, b1);
        } else if (has1 && (C & 1) == 0) {
            *reinterpret_cast<float2*>(o + 2 * pair) = make_float2(a0, a1);
            *reinterpret_cast<float2*>(o + C + 2 * pair) = make_float2(b0, b1);
        } else {
            o[2 * pair] = a0;
            o[C + 2 * pair] = b0;
            if (has1) {
                o[2 * pair + 1] = a1;
                o[C + 2 * pair + 1] = b1;
            }
        }
    } else {
        // C == 2 input; c_out output channels (data_utils.py:79-82, 100-117)
        const int Co = p.c_out;
        float* o = p.out + ((size_t(b) * kBins + f) * p.T + t) * size_t(2 * Co);
        for (int c = 0; c < Co; ++c) {
            float re, im;
            if (c == 0) { re = r0; im = i0; }
            else if (c == 1) { re = r1; im = i1; }
            else if (p.remap == REMAP_STEREO_MONO) { re = r0 + r1; im = i0 + i1; }
            else {
                const float fa = p.merge_f[size_t(b) * (Co - 2) + (c - 2)];
                const float sf = p.merge_sf[size_t(b) * (Co - 2) + (c - 2)];
                re = fa * r0 + sf * r1;
                im = i0 + i1;
            }
            if (p.filter_k > 0) { re *= filt; im *= filt; }
            float a = re, ph = im;
            if (MODE != FM_COMPLEX) {
                a = sqrt_approx(fmaf(re, re, im * im));
                ph = fast_atan2f(im, re);
                if (MODE == FM_LOGMAGPHASE) a = __logf(a + 1e-8f);
            }
            o[c] = a;
            o[Co + c] = ph;
        }
    }
}

}  // namespace iris
