// Evaluation-side chain of metrics.evaluate (metrics.py:59-87): sliding windows over the
// features of one file, overlap-and-add averaging of the model's per-window predictions,
// smoothing pools + threshold, event extraction and the greedy error-rate matching.
// The tensors are small (one file: a few thousand frames x 3 classes); what matters here is
// that the stage stays on the device between the model and the score (no host round trip
// per stage) and that the integer parts are exact.
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

// metrics.py:60-61 -- tf.signal.frame(x, frame_len, step, pad_end=True, axis=-2) then
// transpose (1, 0, 2, 3): x [outer, T, inner] -> out [n_win, outer, frame_len, inner], zero
// beyond the end.
__global__ void k_eval_windows(const float* __restrict__ x, float* __restrict__ out, long long outer,
                               long long T, long long inner, int frame_len, int step, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long in = i % inner;
        long long r = i / inner;
        const long long j = r % frame_len;
        r /= frame_len;
        const long long o = r % outer;
        const long long w = r / outer;
        const long long t = w * step + j;
        out[i] = t < T ? x[(o * T + t) * inner + in] : 0.f;
    }
}

// metrics.py:67-75 -- UpSampling1D(up) + overlap_and_add(preds) / overlap_and_add(ones),
// [..., :L]: preds [n_win, n_p, K] -> out [L, K].  Windows are accumulated in ascending order;
// a position no window covers is 0 / 0 = NaN, as in the reference.
__global__ void k_eval_merge(const float* __restrict__ preds, float* __restrict__ out, int n_win, int n_p,
                             int K, int up, int step, int L) {
    const int F = n_p * up;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * K; i += gridDim.x * blockDim.x) {
        const int t = i / K, k = i - t * K;
        int w_lo = t - F + 1;
        w_lo = w_lo <= 0 ? 0 : (w_lo + step - 1) / step;
        int w_hi = t / step;
        if (w_hi > n_win - 1) w_hi = n_win - 1;
        float s = 0.f, c = 0.f;
        for (int w = w_lo; w <= w_hi; ++w) {
            const int j = (t - w * step) / up;
            s += preds[(size_t(w) * n_p + j) * K + k];
            c += 1.f;
        }
        out[i] = __fdiv_rn(s, c);
    }
}

// AveragePooling1D(k, 1, 'same') on [L, K] (metrics.py:79): mean over the valid cells of
// [t - (k-1)/2, t - (k-1)/2 + k), accumulated in ascending order.
__global__ void k_avg_pool_same1(const float* __restrict__ x, float* __restrict__ out, int L, int K, int k) {
    const int before = (k - 1) / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * K; i += gridDim.x * blockDim.x) {
        const int t = i / K, c = i - t * K;
        const int lo = max(t - before, 0), hi = min(t - before + k, L);
        float s = 0.f;
        for (int u = lo; u < hi; ++u) s = __fadd_rn(s, x[u * K + c]);
        out[i] = __fdiv_rn(s, float(hi - lo));
    }
}

// MaxPooling1D(k, 1, 'same') then `>= thr` -> 0/1 floats (metrics.py:80-81).
__global__ void k_max_pool_same1_thr(const float* __restrict__ x, float* __restrict__ out, int L, int K,
                                     int k, float thr) {
    const int before = (k - 1) / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * K; i += gridDim.x * blockDim.x) {
        const int t = i / K, c = i - t * K;
        const int lo = max(t - before, 0), hi = min(t - before + k, L);
        float m = x[lo * K + c];
        for (int u = lo + 1; u < hi; ++u) m = fmaxf(m, x[u * K + c]);
        out[i] = m >= thr ? 1.f : 0.f;
    }
}

// Challenge_Metric.get_start_end_frame + output_to_metric (metrics.py:109-133, 196-214).
// One CTA.  Per class the change points (y[t] != y[t-1], zero row before t = 0) are ranked
// with a block scan; change point r is the start of event r/2 (r even) or one past its end
// (r odd); an odd count is closed with L.  rows[e] = (class, start, end, int32(((start +
// end) / 2) * hop / sr)) in float64, ordered by class then time; n_rows[0] = events,
// n_rows[1 + c] = events of class c.
__global__ void __launch_bounds__(1024) k_eval_events(const float* __restrict__ y, int L, int K, int hop,
                                                       int sr, int4* __restrict__ rows, int max_rows,
                                                       int32_t* __restrict__ n_rows) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int base = 0;   // rows of the classes before c
    for (int c = 0; c < K; ++c) {
        if (tid == 0) carry = 0;
        __syncthreads();
        for (int t0 = 0; t0 < L; t0 += 1024) {
            const int t = t0 + tid;
            bool ch = false;
            if (t < L) {
                const float cur = y[t * K + c];
                const float prev = t ? y[(t - 1) * K + c] : 0.f;
                ch = cur != prev;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ch);
            if (lane == 0) warp_sum[wid] = __popc(bal);
            __syncthreads();
            int before = carry;
            for (int w = 0; w < wid; ++w) before += warp_sum[w];
            const int r = before + __popc(bal & ((1u << lane) - 1u));
            if (ch) {
                const int e = base + (r >> 1);
                if (e < max_rows) {
                    if (r & 1) rows[e].z = t - 1;
                    else { rows[e].x = c; rows[e].y = t; }
                }
            }
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < 32; ++w) tot += warp_sum[w];
                carry += tot;
            }
            __syncthreads();
        }
        const int n_ch = carry;
        const int n_ev = (n_ch + 1) >> 1;
        if (tid == 0) {
            if ((n_ch & 1) && base + n_ev - 1 < max_rows) rows[base + n_ev - 1].z = L - 1;   // closed with len(data)
            n_rows[1 + c] = n_ev;
        }
        base += n_ev;
        __syncthreads();
    }
    const int n = min(base, max_rows);
    for (int e = tid; e < n; e += 1024) {
        const int4 r = rows[e];
        const double v = ((double(r.y) + double(r.z)) / 2.0) * double(hop) / double(sr);
        rows[e].w = int(v);   // tf.cast(float64 -> int32) truncates
    }
    if (tid == 0) n_rows[0] = base;
}

// metrics.get_er (metrics.py:176-193).  One CTA.  gt [m, 3] (class, start, end), pred rows
// (class, time) with a row stride (2 for a plain [n, 2] tensor, 4 + offset for the events
// kernel's rows).  Both lists are rank-sorted by time (stable); every ground-truth row, in
// order, takes the first remaining prediction of its class with start <= time <= end.
// out = (N = n + m, answer = 2 * matches, m).
__global__ void __launch_bounds__(1024) k_get_er(const int32_t* __restrict__ gt, int m,
                                                  const int32_t* __restrict__ pred, int pred_stride,
                                                  int pred_time_col, const int32_t* __restrict__ n_pred_ptr,
                                                  int n_pred_max, int32_t* __restrict__ order_p,
                                                  int32_t* __restrict__ order_g, int32_t* __restrict__ out) {
    __shared__ int best;
    __shared__ int answer;
    const int tid = threadIdx.x;
    int n = n_pred_ptr ? n_pred_ptr[0] : n_pred_max;
    if (n > n_pred_max) n = n_pred_max;
    for (int i = tid; i < n; i += 1024) {
        const int ti = pred[size_t(i) * pred_stride + pred_time_col];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const int tj = pred[size_t(j) * pred_stride + pred_time_col];
            r += (tj < ti) || (tj == ti && j < i);
        }
        order_p[r] = i;
    }
    for (int i = tid; i < m; i += 1024) {
        const int ti = gt[3 * i + 1];
        int r = 0;
        for (int j = 0; j < m; ++j) {
            const int tj = gt[3 * j + 1];
            r += (tj < ti) || (tj == ti && j < i);
        }
        order_g[r] = i;
    }
    if (tid == 0) answer = 0;
    __syncthreads();
    for (int g = 0; g < m; ++g) {
        if (tid == 0) best = 0x7fffffff;
        __syncthreads();
        const int gi = order_g[g];
        const int gc = gt[3 * gi], gs = gt[3 * gi + 1], ge = gt[3 * gi + 2];
        for (int p = tid; p < n; p += 1024) {
            const int pi = order_p[p];
            if (pi < 0) continue;   // taken
            const int pc = pred[size_t(pi) * pred_stride];
            const int pt = pred[size_t(pi) * pred_stride + pred_time_col];
            if (pc == gc && gs <= pt && pt <= ge) {
                atomicMin(&best, p);
                break;   // later positions of this thread are larger
            }
        }
        __syncthreads();
        if (tid == 0 && best != 0x7fffffff) {
            order_p[best] = -1;
            answer += 2;
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[0] = n + m;
        out[1] = answer;
        out[2] = m;
    }
}

static inline int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return int(g);
}

cudaError_t launch_eval_windows(const float* x, float* out, long long outer, long long T, long long inner,
                                int frame_len, int step, int n_win, cudaStream_t st) {
    const long long total = (long long)n_win * outer * frame_len * inner;
    if (total <= 0) return cudaSuccess;
    k_eval_windows<<<grid_for(total, 256), 256, 0, st>>>(x, out, outer, T, inner, frame_len, step, total);
    return cudaGetLastError();
}
cudaError_t launch_eval_merge(const float* preds, float* out, int n_win, int n_p, int K, int up, int step,
                              int L, cudaStream_t st) {
    if (L * K <= 0) return cudaSuccess;
    k_eval_merge<<<grid_for((long long)L * K, 256), 256, 0, st>>>(preds, out, n_win, n_p, K, up, step, L);
    return cudaGetLastError();
}
cudaError_t launch_eval_smooth(const float* x, float* tmp, float* out, int L, int K, int k_avg, int k_max,
                               float thr, cudaStream_t st) {
    if (L * K <= 0) return cudaSuccess;
    const int g = grid_for((long long)L * K, 256);
    k_avg_pool_same1<<<g, 256, 0, st>>>(x, tmp, L, K, k_avg);
    k_max_pool_same1_thr<<<g, 256, 0, st>>>(tmp, out, L, K, k_max, thr);
    return cudaGetLastError();
}
cudaError_t launch_eval_events(const float* y, int L, int K, int hop, int sr, int32_t* rows, int max_rows,
                               int32_t* n_rows, cudaStream_t st) {
    k_eval_events<<<1, 1024, 0, st>>>(y, L, K, hop, sr, reinterpret_cast<int4*>(rows), max_rows, n_rows);
    return cudaGetLastError();
}
cudaError_t launch_get_er(const int32_t* gt, int m, const int32_t* pred, int pred_stride, int pred_time_col,
                          const int32_t* n_pred_ptr, int n_pred_max, int32_t* order_p, int32_t* order_g,
                          int32_t* out, cudaStream_t st) {
    k_get_er<<<1, 1024, 0, st>>>(gt, m, pred, pred_stride, pred_time_col, n_pred_ptr, n_pred_max, order_p,
                                 order_g, out);
    return cudaGetLastError();
}

}  // namespace iris
