// Spectrogram-bank path: the reference's own data format (utils.load_data, utils.py:88-94 --
// a pickled list of pre-computed complex spectrograms [257, t, 2C]; pipeline.py:113-175 mixes
// those).  Banks registered in this format are mixed in the SPECTROGRAM domain, exactly the
// arithmetic of pipeline.merge_complex_specs (pipeline.py:29-106):
//     out[f, t, :] = bg[f, t + o_b, :] + sum_v (g_v * voice_v[f, t + s_v, :]) * keep_v
//                                      + sum_n  g_n * noise_n[f, t + s_n, :]
// with the products and sums rounded separately in the reference's order (no FMA contraction),
// so the complex output is bit-identical to the CPU restatement.  The per-cell epilogue
// (masks, remap, stft_filter, mag / phase / log) is the one of the fused waveform kernel
// (iris_epilogue.cuh).  Pure streaming: every source cell is read once, every output cell
// written once; a warp covers 32 consecutive frames of one bin row (512 contiguous bytes per
// source for 2 channels).
#include "iris_common.cuh"
#include "iris_epilogue.cuh"
#include "iris_launch.h"

namespace iris {

// frame "active" iff any coefficient of the frame (any bin, re or im, any channel) is > 0
// (pipeline.py:55).  One thread per frame; consecutive threads read consecutive cells of a
// bin row.  items [n_items] spectrograms [F, t_i, W] packed back to back.
__global__ void k_spec_activity(const float* __restrict__ specs, const int64_t* __restrict__ frame_off,
                                int n_items, int F, int W, int max_frames, uint8_t* __restrict__ activity) {
    const int item = blockIdx.y;
    const int64_t f0 = frame_off[item];
    const int tI = int(frame_off[item + 1] - f0);
    const float* base = specs + size_t(f0) * F * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < tI; t += gridDim.x * blockDim.x) {
        float mx = 0.f;
        for (int f = 0; f < F; ++f) {
            const float* cell = base + (size_t(f) * tI + t) * W;
            for (int w = 0; w < W; ++w) mx = fmaxf(mx, cell[w]);
        }
        activity[size_t(item) * max_frames + t] = mx > 0.f ? 1 : 0;
    }
}

// Seg of a spectrogram bank: base = the item's [F, tI, 2C] array, pair_stride = tI.
template <int MODE>
__global__ void __launch_bounds__(128) k_specmix(const __grid_constant__ FusedParams p) {
    const int b = blockIdx.z;
    const int f = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.T) return;
    const int C = p.C, W = 2 * C;
    const int s0 = p.seg_ptr[b], s1 = p.seg_ptr[b + 1];
    // SpecAugment masks of this cell (transforms.py:12-40): product of 0/1 factors
    float m = 1.f;
    if (p.tmask != nullptr) {
        const int32_t* tm = p.tmask + size_t(b) * p.n_tmask * 2;
        for (int i = 0; i < p.n_tmask; ++i)
            if (unsigned(t - tm[2 * i + 1]) < unsigned(tm[2 * i])) m = 0.f;
    }
    if (p.fmask != nullptr) {
        const int32_t* fm = p.fmask + size_t(b) * p.n_fmask * 2;
        for (int i = 0; i < p.n_fmask; ++i)
            if (unsigned(f - fm[2 * i + 1]) < unsigned(fm[2 * i])) m = 0.f;
    }
    for (int pair = 0; pair < p.n_pairs; ++pair) {
        const bool has1 = 2 * pair + 1 < C;
        float r0 = 0.f, r1 = 0.f, i0 = 0.f, i1 = 0.f;
        bool first = true;
        for (int s = s0; s < s1; ++s) {
            const Seg sg = p.segs[s];
            if (t < sg.t_lo || t >= sg.t_hi) continue;
            if (sg.keep_idx >= 0 && p.keep[sg.keep_idx] == 0) continue;
            const float* cell = sg.base + (size_t(f) * sg.pair_stride + size_t(t + sg.shift)) * W;
            float x0, x1 = 0.f, y0, y1 = 0.f;
            if (C == 2) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(cell));
                x0 = v.x; x1 = v.y; y0 = v.z; y1 = v.w;
            } else {
                x0 = __ldg(cell + 2 * pair);
                y0 = __ldg(cell + C + 2 * pair);
                if (has1) {
                    x1 = __ldg(cell + 2 * pair + 1);
                    y1 = __ldg(cell + C + 2 * pair + 1);
                }
            }
            if (first) {   // the background enters with gain 1 (pipeline.py:35)
                r0 = __fmul_rn(sg.gain, x0); r1 = __fmul_rn(sg.gain, x1);
                i0 = __fmul_rn(sg.gain, y0); i1 = __fmul_rn(sg.gain, y1);
                first = false;
            } else {       // spec += gain * source (pipeline.py:81, 106): product, then sum
                r0 = __fadd_rn(r0, __fmul_rn(sg.gain, x0)); r1 = __fadd_rn(r1, __fmul_rn(sg.gain, x1));
                i0 = __fadd_rn(i0, __fmul_rn(sg.gain, y0)); i1 = __fadd_rn(i1, __fmul_rn(sg.gain, y1));
            }
        }
        store_bin<MODE>(p, b, f, t, pair, has1, r0, i0, r1, i1, m);
    }
}

cudaError_t launch_spec_activity(const float* specs, const int64_t* frame_off, int n_items, int F, int W,
                                 int max_frames, uint8_t* activity, cudaStream_t st) {
    if (n_items <= 0 || max_frames <= 0) return cudaSuccess;
    dim3 grid(unsigned((max_frames + 127) / 128), unsigned(n_items));
    k_spec_activity<<<grid, 128, 0, st>>>(specs, frame_off, n_items, F, W, max_frames, activity);
    return cudaGetLastError();
}

cudaError_t launch_specmix(const FusedParams& p, int mode, cudaStream_t st) {
    if (p.B <= 0 || p.T <= 0) return cudaSuccess;
    if (p.B > 65535) return cudaErrorInvalidValue;
    dim3 grid(unsigned((p.T + 127) / 128), unsigned(kBins), unsigned(p.B));
    switch (mode) {
        case FM_COMPLEX: k_specmix<FM_COMPLEX><<<grid, 128, 0, st>>>(p); break;
        case FM_MAGPHASE: k_specmix<FM_MAGPHASE><<<grid, 128, 0, st>>>(p); break;
        case FM_LOGMAGPHASE: k_specmix<FM_LOGMAGPHASE><<<grid, 128, 0, st>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace iris
