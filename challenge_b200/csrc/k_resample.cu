// torchaudio.compliance.kaldi.resample_waveform (data_utils.py:20-21; Kaldi's LinearResample): a
// polyphase windowed-sinc FIR.  With U_in = orig / gcd and U_out = new / gcd,
//   out[c, u * U_out + i] = sum_j w[i][j] * wav[c, first[i] + u * U_in + j]      (zero outside)
// One thread per output sample; the U_out x W weight table sits in shared memory when it fits,
// consecutive threads read overlapping, nearly contiguous input windows (L1 / L2 hits).
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

__global__ void __launch_bounds__(256) k_resample(const float* __restrict__ wav, float* __restrict__ out,
                                                  int n_chan, long long n_in, long long n_out, int u_in,
                                                  int u_out, int W, const int32_t* __restrict__ first,
                                                  const float* __restrict__ weights, int w_in_smem) {
    extern __shared__ float s_w[];
    const float* wt = weights;
    if (w_in_smem) {
        for (int i = threadIdx.x; i < u_out * W; i += blockDim.x) s_w[i] = weights[i];
        __syncthreads();
        wt = s_w;
    }
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_out) return;
    const long long u = n / u_out;
    const int i = int(n - u * u_out);
    const long long k0 = (long long)first[i] + u * u_in;
    const float* w = wt + size_t(i) * W;
    for (int c = blockIdx.y; c < n_chan; c += gridDim.y) {
        const float* x = wav + size_t(c) * n_in;
        float acc = 0.f;
        for (int j = 0; j < W; ++j) {
            const long long k = k0 + j;
            if (k >= 0 && k < n_in) acc = fmaf(w[j], x[k], acc);
        }
        out[size_t(c) * n_out + n] = acc;
    }
}

cudaError_t launch_resample(const float* wav, float* out, int n_chan, long long n_in, long long n_out,
                            int u_in, int u_out, int W, const int32_t* first, const float* weights,
                            cudaStream_t st) {
    if (n_out <= 0) return cudaSuccess;
    const size_t tab = size_t(u_out) * W * sizeof(float);
    const int in_smem = tab <= 96 * 1024;
    if (in_smem && tab > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(unsigned((n_out + 255) / 256), unsigned(n_chan < 8 ? n_chan : 8));
    k_resample<<<grid, 256, in_smem ? tab : 0, st>>>(wav, out, n_chan, n_in, n_out, u_in, u_out, W, first,
                                                      weights, in_smem);
    return cudaGetLastError();
}

}  // namespace iris
