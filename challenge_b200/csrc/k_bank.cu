// Bank registration = data_utils.load_wav minus decode and STFT (data_utils.py:9-29):
// RMS-normalise each source over all channels and samples (normalize, 32-34) and lay it
// out as the reflect-padded, CHANNEL-PAIR-INTERLEAVED waveform
//     P[pair][i][c] = x~[2*pair + c][i - 256],   i in [0, 256*(kT+1)),  c in {0,1}
// (a missing odd channel is zero), exactly the samples torch.stft(center=True,
// pad_mode='reflect') frames at hop 256.  Row h of a pair plane (256 float2 = 2 KB) is both
// the 2nd half of frame h-1 and the 1st half of frame h, so the hot kernel fetches any frame
// range of a source with ONE aligned bulk copy per channel pair, and a lane reads the packed
// complex FFT input (re = even channel, im = odd channel) with one 64-bit shared load.
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

__global__ void __launch_bounds__(1024) k_bank_prepare(const float* __restrict__ wav,
                                                       const int64_t* __restrict__ offsets,
                                                       const int64_t* __restrict__ pad_offsets,
                                                       int n_chan, int normalize,
                                                       float* __restrict__ padded) {
    const int item = blockIdx.x;
    const int64_t n = offsets[item + 1] - offsets[item];          // samples per channel
    const float* x = wav + offsets[item] * n_chan;                 // [C, n]
    const int64_t total = n * n_chan;
    __shared__ double s_part[32];
    __shared__ float s_scale;
    float scale_div = 1.f;
    if (normalize) {
        double acc = 0.0;
        for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
            const float v = x[i];
            acc += double(v) * double(v);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < int(blockDim.x >> 5); ++w) t += s_part[w];
            // rms = sqrt(mean(wav^2)) * 10  (data_utils.py:33), fp32 like torch
            const float mean = float(t / double(total));
            s_scale = sqrtf(mean) * 10.f;
        }
        __syncthreads();
        scale_div = s_scale;
    }
    const int n_pairs = (n_chan + 1) >> 1;
    const int64_t kT = 1 + n / kHop;
    const int64_t plen = 256 * (kT + 1);
    float2* P = reinterpret_cast<float2*>(padded) + pad_offsets[item] * n_pairs;   // [pairs, plen]
    for (int64_t i = threadIdx.x; i < plen * n_pairs; i += blockDim.x) {
        const int64_t pr = i / plen, j = i - pr * plen;
        int64_t src = j - 256;
        if (src < 0) src = -src;
        if (src >= n) src = 2 * (n - 1) - src;
        const int c0 = int(2 * pr), c1 = c0 + 1;
        float a = x[c0 * n + src];
        float b = c1 < n_chan ? x[c1 * n + src] : 0.f;
        if (normalize) { a = a / scale_div; b = b / scale_div; }
        P[i] = make_float2(a, b);
    }
}

cudaError_t launch_bank_prepare(const float* wav, const int64_t* d_offsets,
                                const int64_t* d_pad_offsets, int n_items, int n_chan,
                                int normalize, float* padded, cudaStream_t stream) {
    if (n_items <= 0) return cudaSuccess;
    k_bank_prepare<<<n_items, 1024, 0, stream>>>(wav, d_offsets, d_pad_offsets, n_chan, normalize,
                                                 padded);
    return cudaGetLastError();
}

}  // namespace iris
