// metrics.py counting as an integer reduction:
//  * er_score(smoothing=False) core (metrics.py:217-266): per sample the number of true
//    events, predicted events and true events hit by the midpoint of a predicted event of
//    the same class -> int32 triple (n_true, n_pred, correct);
//  * tfa F1Score(threshold=0.5, 'micro') counts (metrics.py:290-298): TP / FP / FN over the
//    whole [B,T,K] tensor with pred = y_pred > 0.5 (strict), accumulated into 3 x uint64.
// One CTA per sample.  Events are runs of ones in per-class bitmaps held in shared memory.
// er_score(smoothing=True) average-pools y_pred with stride 31 first (metrics.py:222-224), so its
// events live on a time base of Tp = ceil(T / 31) frames while y_true keeps T: the reference
// compares the frame indices of the two bases as they are (metrics.py:256-266), and so does the
// kernel when Tp != T (the F1 counts need equal shapes and are skipped then).
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

// start of the run of ones that contains bit position e (bit e must be set)
__device__ __forceinline__ int run_start(const uint32_t* bits, int e) {
    int w = e >> 5;
    const int bpos = e & 31;
    uint32_t zeros_below = ~bits[w] & ((bpos == 0) ? 0u : (0xffffffffu >> (32 - bpos)));
    while (true) {
        if (zeros_below) return (w << 5) + (32 - __clz(zeros_below));
        if (w == 0) return 0;
        --w;
        zeros_below = ~bits[w];
    }
}

__global__ void __launch_bounds__(128) k_metric_counts(const float* __restrict__ y_true,
                                                       const float* __restrict__ y_pred, int T,
                                                       int Tp, int K, float thr, int32_t* triples,
                                                       unsigned long long* tpfpfn,
                                                       unsigned long long* sums) {
    extern __shared__ uint32_t s_bits[];
    const int Wt = (T + 31) >> 5, Wp = (Tp + 31) >> 5;
    const int W = max(Wt, Wp);           // one row pitch for the three bitmaps
    uint32_t* tb = s_bits;               // [K][W] y_true >= thr (T frames)
    uint32_t* pb = tb + K * W;           // [K][W] y_pred >= thr (Tp frames)
    uint32_t* mb = pb + K * W;           // [K][W] midpoints of predicted events (pred time base)
    __shared__ int s_cnt[6];
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const float* yt = y_true + size_t(b) * T * K;
    const float* yp = y_pred + size_t(b) * Tp * K;
    for (int i = threadIdx.x; i < K * W; i += blockDim.x) mb[i] = 0u;
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    int tp = 0, fp = 0, fn = 0;
    const int T32 = W << 5;
    for (int t = threadIdx.x; t < T32; t += blockDim.x) {   // warp-uniform trip count
        for (int c = 0; c < K; ++c) {
            bool a = false, q = false;
            float vp = 0.f;
            if (t < T) a = yt[size_t(t) * K + c] >= thr;   // metrics.py:221
            if (t < Tp) {
                vp = yp[size_t(t) * K + c];
                q = vp >= thr;                  // metrics.py:225
            }
            if (t < T && T == Tp) {
                const bool f1p = vp > thr;      // tfa F1Score: strict
                tp += (f1p && a);
                fp += (f1p && !a);
                fn += (!f1p && a);
            }
            const uint32_t wa = __ballot_sync(0xffffffffu, a);
            const uint32_t wq = __ballot_sync(0xffffffffu, q);
            if (lane == 0) {
                tb[c * W + (t >> 5)] = wa;
                pb[c * W + (t >> 5)] = wq;
            }
        }
    }
    __syncthreads();
    // predicted events: count + midpoint bitmap  (metrics.py:243-256)
    int n_true = 0, n_pred = 0, correct = 0;
    for (int i = threadIdx.x; i < K * W; i += blockDim.x) {
        const int c = i / W, w = i - c * W;
        const uint32_t* bits = pb + c * W;
        const uint32_t cur = bits[w];
        const uint32_t prev_msb = w > 0 ? bits[w - 1] >> 31 : 0u;
        const uint32_t next_lsb = (w + 1 < W) ? (bits[w + 1] & 1u) : 0u;
        const uint32_t starts = cur & ~((cur << 1) | prev_msb);
        uint32_t ends = cur & ~((cur >> 1) | (next_lsb << 31));
        n_pred += __popc(starts);
        while (ends) {
            const int e = (w << 5) + __ffs(ends) - 1;
            ends &= ends - 1;
            const int s = run_start(bits, e);
            const int m = (s + e) >> 1;             // int64((start+end)/2)
            atomicOr(&mb[c * W + (m >> 5)], 1u << (m & 31));
        }
    }
    __syncthreads();
    // true events: count + hit test against midpoints of the same class (metrics.py:259-266)
    for (int i = threadIdx.x; i < K * W; i += blockDim.x) {
        const int c = i / W, w = i - c * W;
        const uint32_t* bits = tb + c * W;
        const uint32_t* mid = mb + c * W;
        const uint32_t cur = bits[w];
        const uint32_t prev_msb = w > 0 ? bits[w - 1] >> 31 : 0u;
        const uint32_t next_lsb = (w + 1 < W) ? (bits[w + 1] & 1u) : 0u;
        const uint32_t starts = cur & ~((cur << 1) | prev_msb);
        uint32_t ends = cur & ~((cur >> 1) | (next_lsb << 31));
        n_true += __popc(starts);
        while (ends) {
            const int e = (w << 5) + __ffs(ends) - 1;
            ends &= ends - 1;
            const int s = run_start(bits, e);
            bool hit = false;
            for (int ww = s >> 5; ww <= (e >> 5); ++ww) {
                uint32_t msk = 0xffffffffu;
                if (ww == (s >> 5)) msk &= 0xffffffffu << (s & 31);
                if (ww == (e >> 5)) msk &= 0xffffffffu >> (31 - (e & 31));
                if (mid[ww] & msk) hit = true;
            }
            correct += hit;
        }
    }
    // block reduce the six counters
    int vals[6] = {n_true, n_pred, correct, tp, fp, fn};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&s_cnt[k], v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        triples[3 * b + 0] = s_cnt[0];
        triples[3 * b + 1] = s_cnt[1];
        triples[3 * b + 2] = s_cnt[2];
        if (tpfpfn) {
            if (s_cnt[3]) atomicAdd(&tpfpfn[0], (unsigned long long)s_cnt[3]);
            if (s_cnt[4]) atomicAdd(&tpfpfn[1], (unsigned long long)s_cnt[4]);
            if (s_cnt[5]) atomicAdd(&tpfpfn[2], (unsigned long long)s_cnt[5]);
        }
        if (sums) {   // batch totals of (n_true, n_pred, correct): the all-reduce payload
            if (s_cnt[0]) atomicAdd(&sums[0], (unsigned long long)s_cnt[0]);
            if (s_cnt[1]) atomicAdd(&sums[1], (unsigned long long)s_cnt[1]);
            if (s_cnt[2]) atomicAdd(&sums[2], (unsigned long long)s_cnt[2]);
        }
    }
}

// score = (n_true + n_pred - 2*correct) / clip(n_true, 1, max_b n_true)   (metrics.py:268-273)
__global__ void __launch_bounds__(256) k_er_finalize(const int32_t* __restrict__ triples, int B,
                                                     float* __restrict__ er) {
    __shared__ int s_m[8];
    int m = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) m = max(m, triples[3 * i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_m[0];
    for (int w = 1; w < 8; ++w) m = max(m, s_m[w]);
    const float hi = float(m);
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float nt = float(triples[3 * i]);
        const float score = nt + float(triples[3 * i + 1]) - 2.f * float(triples[3 * i + 2]);
        er[i] = score / fminf(fmaxf(nt, 1.f), hi);   // tf.clip_by_value = min(max(x, lo), hi)
    }
}

cudaError_t launch_metric_counts(const float* y_true, const float* y_pred, int B, int T, int Tp, int K,
                                 float threshold, int32_t* triples, unsigned long long* tpfpfn,
                                 unsigned long long* sums, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (Tp != T && tpfpfn != nullptr) return cudaErrorInvalidValue;   // F1 counts need equal shapes
    const int W = (max(T, Tp) + 31) >> 5;
    const size_t smem = size_t(3) * K * W * sizeof(uint32_t);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    k_metric_counts<<<B, 128, smem, stream>>>(y_true, y_pred, T, Tp, K, threshold, triples, tpfpfn, sums);
    return cudaGetLastError();
}

cudaError_t launch_er_finalize(const int32_t* triples, int B, float* er, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    k_er_finalize<<<1, 256, 0, stream>>>(triples, B, er);
    return cudaGetLastError();
}

}  // namespace iris
