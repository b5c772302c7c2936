// Second pass of the log-mel features: per-clip min-max normalisation
// (data_utils.minmax, data_utils.py:37-47 with utils.safe_div, utils.py:114-116)
// followed by log(x + 1e-8) (data_utils.log_on_mel, data_utils.py:50-55), in place.
// The per-clip min/max were reduced by k_fused (atomicMax on (~bits, bits)).
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

__device__ __forceinline__ float post_one(float x, float mn, float den, int do_minmax, int do_log) {
    if (do_minmax) x = (x - mn) / den;
    if (do_log) x = logf(x + 1e-8f);
    return x;
}

__global__ void __launch_bounds__(256) k_logmel_post(float* __restrict__ x,
                                                     const uint32_t* __restrict__ minmax,
                                                     size_t per_clip, int do_minmax, int do_log) {
    const int b = blockIdx.y;
    float mn = 0.f, den = 1.f;
    if (do_minmax) {
        mn = __uint_as_float(~minmax[2 * b]);
        const float mx = __uint_as_float(minmax[2 * b + 1]);
        den = fmaxf(mx - mn, 1e-8f);
    }
    float* base = x + size_t(b) * per_clip;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const size_t i0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if ((per_clip & 3) == 0) {
        float4* v = reinterpret_cast<float4*>(base);
        for (size_t i = i0; i < per_clip / 4; i += stride) {
            float4 a = v[i];
            a.x = post_one(a.x, mn, den, do_minmax, do_log);
            a.y = post_one(a.y, mn, den, do_minmax, do_log);
            a.z = post_one(a.z, mn, den, do_minmax, do_log);
            a.w = post_one(a.w, mn, den, do_minmax, do_log);
            v[i] = a;
        }
    } else {
        for (size_t i = i0; i < per_clip; i += stride)
            base[i] = post_one(base[i], mn, den, do_minmax, do_log);
    }
}

cudaError_t launch_logmel_post(float* x, const uint32_t* minmax, int B, size_t per_clip,
                               int do_minmax, int do_log, cudaStream_t stream) {
    if (B <= 0 || per_clip == 0 || (!do_minmax && !do_log)) return cudaSuccess;
    size_t work = (per_clip & 3) == 0 ? per_clip / 4 : per_clip;
    int gx = int((work + 256 * 4 - 1) / (256 * 4));
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    for (int b0 = 0; b0 < B; b0 += 65535) {
        int nb = B - b0 < 65535 ? B - b0 : 65535;
        dim3 grid(gx, nb);
        k_logmel_post<<<grid, 256, 0, stream>>>(x + size_t(b0) * per_clip,
                                                minmax ? minmax + 2 * size_t(b0) : nullptr,
                                                per_clip, do_minmax, do_log);
    }
    return cudaGetLastError();
}

}  // namespace iris
