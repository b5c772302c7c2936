// Frame-label construction and same-class overlap rejection of
// pipeline.merge_complex_specs (pipeline.py:41-84) + data_utils.to_frame_labels
// (data_utils.py:64-70), as an integer-exact kernel over precomputed per-frame
// activity flags (activity = reduce_max(voice, (f, 2C)) > 0, pipeline.py:55).
// One CTA per clip; voices are visited in order because acceptance of voice v depends
// on the labels of the voices accepted before it (pipeline.py:78-84).
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

__global__ void __launch_bounds__(256) k_labels(const LabelParams p) {
    const int b = blockIdx.x;
    const int TK = p.T * p.K;
    float* L = p.frame_labels + size_t(b) * TK;
    __shared__ float s_max[8];
    __shared__ int s_keep;
    for (int i = threadIdx.x; i < TK; i += blockDim.x) L[i] = 0.f;
    const int nv = p.n_voices[b];
    for (int v = 0; v < p.V; ++v) {
        float* lv = p.labels_vtk ? p.labels_vtk + (size_t(b) * p.V + v) * TK : nullptr;
        if (v >= nv) {
            if (lv) for (int i = threadIdx.x; i < TK; i += blockDim.x) lv[i] = 0.f;
            if (threadIdx.x == 0) p.keep[size_t(b) * p.V + v] = 0;
            continue;
        }
        const int id = p.voice_id[size_t(b) * p.V + v];
        const int shift = p.voice_shift[size_t(b) * p.V + v];
        const int kT = p.n_frames[id];
        const uint8_t* act = p.activity + size_t(id) * p.act_stride;
        const float* lab = p.bank_labels + size_t(id) * p.K;
        // max over (t, c) of (sum of accepted labels + candidate)   (pipeline.py:78)
        float mx = 0.f;
        for (int i = threadIdx.x; i < TK; i += blockDim.x) {
            const int t = i / p.K, c = i - t * p.K;
            const int k = t + shift;
            const float a = (k >= 0 && k < kT && act[k]) ? 1.f : 0.f;
            mx = fmaxf(mx, L[i] + lab[c] * a);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = s_max[0];
            for (int w = 1; w < int(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]);
            s_keep = (m < 2.f) ? 1 : 0;                       // no_overlap (pipeline.py:78-79)
            p.keep[size_t(b) * p.V + v] = uint8_t(s_keep);
        }
        __syncthreads();
        const float keep = s_keep ? 1.f : 0.f;
        for (int i = threadIdx.x; i < TK; i += blockDim.x) {
            const int t = i / p.K, c = i - t * p.K;
            const int k = t + shift;
            const float a = (k >= 0 && k < kT && act[k]) ? 1.f : 0.f;
            const float cand = lab[c] * a * keep;              // l * no_overlap (pipeline.py:84)
            L[i] += cand;
            if (lv) lv[i] = cand;
        }
        __syncthreads();
    }
}

cudaError_t launch_labels(const LabelParams& p, cudaStream_t stream) {
    if (p.B <= 0) return cudaSuccess;
    k_labels<<<p.B, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace iris
