// Fused hot path: gather + time-domain mix of the gained source frames (TMA bulk
// copies into shared memory), window, 512-point FFT per frame (two real channels
// packed into one complex transform, one half-warp per FFT), then the epilogue
// (SpecAugment masks, channel remap, stft_filter, complex / mag-phase / log-mag-phase
// output, or magnitude -> sparse mel -> per-clip min/max), writing each feature once.
//
// Replaces, for one output clip, the chain
//   data_utils.load_wav (STFT, data_utils.py:9-29)  ->  pipeline.merge_complex_specs
//   (pipeline.py:6-110)  ->  data_utils.augment (58-61)  ->  stereo_mono /
//   random_merge_aug / stft_filter (79-136)  ->  transforms.complex_to_magphase
//   (transforms.py:111-123)  ->  magphase_to_mel (51-77)  [-> minmax/log in k_post.cu]
// using linearity of the STFT: sum_k g_k STFT(src_k)[frame] = FFT(w * sum_k g_k frame_k).
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

constexpr int kSlots = 16;                 // half-warp FFT slots per CTA
constexpr int kTF = 16;                    // output frames per tile (one per slot)
constexpr int kRows = kTF + 1;             // 256-sample rows staged per channel
constexpr int kRowFloats = 256 + 16;       // 16-float skew => rows j, j+1 hit disjoint banks
constexpr int kChanFloats = kRows * kRowFloats;
constexpr int kStageFloats = 2 * kChanFloats;
constexpr int kStages = 2;
constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;  // + one TMA producer warp
constexpr int kMaxMelBinsWindow = kXchSlotFloats / 2;  // mags of 2 channels alias the slot buffer

static size_t fused_smem_bytes(int mode, int n_mel) {
    size_t floats = size_t(kStages) * kStageFloats + size_t(kSlots) * kXchSlotFloats + 1024 + 512 +
                    16 /*barriers*/ + 16;
    if (mode == FM_MEL) floats += size_t(n_mel) * 32;
    return floats * 4;
}

struct TileCoord {
    int b, pair, t0;
};
__device__ __forceinline__ TileCoord decode_tile(int tile, int tpc, int n_pairs) {
    TileCoord tc;
    int per_clip = tpc * n_pairs;
    tc.b = tile / per_clip;
    int r = tile - tc.b * per_clip;
    tc.pair = r / tpc;
    tc.t0 = (r - tc.pair * tpc) * kTF;
    return tc;
}

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
            a0 = sqrtf(r0 * r0 + i0 * i0);           // transforms.py:116
            b0 = atan2f(i0, r0);                     // transforms.py:117
            a1 = sqrtf(r1 * r1 + i1 * i1);
            b1 = atan2f(i1, r1);
            if (MODE == FM_LOGMAGPHASE) {            // transforms.py:80-86
                a0 = logf(a0 + 1e-8f);
                a1 = logf(a1 + 1e-8f);
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
                a = sqrtf(re * re + im * im);
                ph = atan2f(im, re);
                if (MODE == FM_LOGMAGPHASE) a = logf(a + 1e-8f);
            }
            o[c] = a;
            o[Co + c] = ph;
        }
    }
}

template <int MODE>
__global__ void __maxnreg__(112) k_fused(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    float* xch = stage + kStages * kStageFloats;
    float2* s_tw = reinterpret_cast<float2*>(xch + kSlots * kXchSlotFloats);
    float* s_wh = reinterpret_cast<float*>(s_tw + 512);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wh + 512);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    float* meltile = reinterpret_cast<float*>(bars + 8) + 16;

    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += kThreads) {
        s_tw[i] = p.tw[i];
        s_wh[i] = p.whalf[i];
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeThreads / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int tpc = (p.T + kTF - 1) / kTF;
    const int n_tiles = p.B * p.n_pairs * tpc;

    if (tid >= kComputeThreads) {
        // ===== TMA producer: one elected lane streams 1 KB rows of every contributing
        // segment of every tile into the stage ring =====
        if (tid == kComputeThreads) {
            int si = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(tile, tpc, p.n_pairs);
                const int s_end = p.seg_ptr[tc.b + 1];
                const int nch = (2 * tc.pair + 1 < p.C) ? 2 : 1;
                const int t_end = min(tc.t0 + kTF, p.T);
                for (int s = p.seg_ptr[tc.b]; s < s_end; ++s) {
                    const Seg sg = p.segs[s];
                    if (sg.keep_idx >= 0 && p.keep[sg.keep_idx] == 0) continue;
                    const int f_lo = max(sg.t_lo, tc.t0), f_hi = min(sg.t_hi, t_end);
                    if (f_lo >= f_hi) continue;
                    const int j_lo = f_lo - tc.t0;
                    const int n_rows = f_hi - f_lo + 1;   // frames j..j' need rows j..j'+1
                    mbar_wait(&empty[si], phase ^ 1);
                    mbar_arrive_expect_tx(&full[si], uint32_t(n_rows) * 1024u * uint32_t(nch));
                    float* dst = stage + si * kStageFloats + j_lo * kRowFloats;
                    for (int c = 0; c < nch; ++c) {
                        const float* src = sg.base + size_t(2 * tc.pair + c) * size_t(sg.chan_stride) +
                                           size_t(f_lo + sg.shift) * 256;
                        for (int r = 0; r < n_rows; ++r)
                            bulk_g2s(dst + c * kChanFloats + r * kRowFloats, src + r * 256, 1024,
                                     &full[si]);
                    }
                    if (++si == kStages) { si = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===== compute warps: one 512-point FFT per half-warp slot =====
    const int warp = tid >> 5, lane = tid & 31, hw = lane >> 4, n2 = lane & 15;
    const int slot = warp * 2 + hw;
    const unsigned hmask = hw ? 0xFFFF0000u : 0x0000FFFFu;
    float* xs = xch + slot * kXchSlotFloats;
    const int ka = n2, kb = (n2 == 0) ? 16 : 32 - n2;
    const bool l0 = (n2 == 0);

    int si = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, tpc, p.n_pairs);
        const int b = tc.b;
        const int t = tc.t0 + slot;
        const bool in_range = t < p.T;
        const bool has1 = (2 * tc.pair + 1 < p.C);
        const int t_end = min(tc.t0 + kTF, p.T);

        float re[32], im[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { re[i] = 0.f; im[i] = 0.f; }

        // ---- gather + mix: acc += gain * frame_k of every contributing segment ----
        const int s_end = p.seg_ptr[b + 1];
        for (int s = p.seg_ptr[b]; s < s_end; ++s) {
            const Seg sg = p.segs[s];
            if (sg.keep_idx >= 0 && p.keep[sg.keep_idx] == 0) continue;
            if (max(sg.t_lo, tc.t0) >= min(sg.t_hi, t_end)) continue;
            mbar_wait(&full[si], phase);
            if (in_range && t >= sg.t_lo && t < sg.t_hi) {
                const float g = sg.gain;
                const float* r0 = stage + si * kStageFloats + slot * kRowFloats + n2;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    re[i] = fmaf(g, r0[16 * i], re[i]);
                    re[16 + i] = fmaf(g, r0[kRowFloats + 16 * i], re[16 + i]);
                }
                if (has1) {
                    const float* r1 = r0 + kChanFloats;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        im[i] = fmaf(g, r1[16 * i], im[i]);
                        im[16 + i] = fmaf(g, r1[kRowFloats + 16 * i], im[16 + i]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[si]);
            if (++si == kStages) { si = 0; phase ^= 1; }
        }

        // ---- SpecAugment time mask of this frame (transforms.py:12-40) ----
        float mt = 1.f;
        if (p.tmask != nullptr) {
            const int32_t* tm = p.tmask + size_t(b) * p.n_tmask * 2;
            for (int i = 0; i < p.n_tmask; ++i) {
                const int size = tm[2 * i], off = tm[2 * i + 1];
                if (t >= off && t < off + size) mt = 0.f;
            }
        }
        int fm_size[4] = {0, 0, 0, 0}, fm_off[4] = {0, 0, 0, 0};
        if (p.fmask != nullptr) {
            const int32_t* fmk = p.fmask + size_t(b) * p.n_fmask * 2;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < p.n_fmask) { fm_size[i] = fmk[2 * i]; fm_off[i] = fmk[2 * i + 1]; }
        }
        auto freq_mult = [&](int f) -> float {
            float m = 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (f >= fm_off[i] && f < fm_off[i] + fm_size[i]) m = 0.f;
            return m;
        };

        // a fully time-masked frame has zero magnitude everywhere: no FFT needed for mel
        const bool do_fft = in_range && !(MODE == FM_MEL && mt == 0.f);

        cpx Za[16], Zb[16];
        if (do_fft) {
            cpx v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float w = s_wh[16 * i + n2];
                v[i] = cpx{re[i] * w, im[i] * w};
            }
            Fft<32>::run(v);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                const float2 w = s_tw[k1 * 16 + n2];
                v[k1] = cmul(v[k1], cpx{w.x, w.y});
            }
            // ---- exchange through shared memory, 4 rounds of 8 k1 ----
#pragma unroll
            for (int rho = 0; rho < 4; ++rho) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    *reinterpret_cast<float4*>(xs + xch_write_off(a, n2)) =
                        make_float4(v[8 * rho + 2 * a].x, v[8 * rho + 2 * a].y,
                                    v[8 * rho + 2 * a + 1].x, v[8 * rho + 2 * a + 1].y);
                __syncwarp(hmask);
                if ((ka >> 3) == rho) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(ka & 7, j));
                        Za[j] = cpx{z.x, z.y};
                    }
                }
                if ((kb >> 3) == rho) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(kb & 7, j));
                        Zb[j] = cpx{z.x, z.y};
                    }
                }
                __syncwarp(hmask);
            }
            Fft<16>::run(Za);
            Fft<16>::run(Zb);
        }

        // ---- epilogue ----
        // bin f = ka+32*k2 pairs with its mirror 512-f held in Zb[15-k2] (lane 0: Za[(16-k2)&15]);
        // bin f = kb+32*k2 pairs with Za[15-k2] (lane 0: Zb[15-k2]).  The 0.5 of the
        // two-channel split is folded into the window table.
        if (MODE == FM_MEL) {
            float* mg = xs;   // [2][mel_f_n] magnitudes, aliases the exchange slot
            const int f_lo = p.mel_f_lo, f_n = p.mel_f_n;
            if (do_fft) {
                auto emit = [&](int f, cpx zf, cpx zm) {
                    const int fi = f - f_lo;
                    if (fi >= 0 && fi < f_n) {
                        const float r0 = zf.x + zm.x, i0 = zf.y - zm.y;
                        const float r1 = zf.y + zm.y, i1 = zm.x - zf.x;
                        const float m = freq_mult(f);
                        mg[fi] = sqrtf(r0 * r0 + i0 * i0) * m;
                        mg[f_n + fi] = sqrtf(r1 * r1 + i1 * i1) * m;
                    }
                };
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const cpx pa = l0 ? Za[(16 - k2) & 15] : Zb[15 - k2];
                    const cpx pb = l0 ? Zb[15 - k2] : Za[15 - k2];
                    emit(ka + 32 * k2, Za[k2], pa);
                    emit(kb + 32 * k2, Zb[k2], pb);
                }
                if (l0) emit(256, Za[8], Za[8]);
            }
            __syncwarp();                         // reconverge both half-warps
            named_bar_sync(1, kComputeThreads);   // previous tile's store phase is over
            if (in_range) {
                for (int m = n2; m < p.n_mel; m += 16) {
                    float a0 = 0.f, a1 = 0.f;
                    if (do_fft) {
                        const int e = p.mel_ptr[m + 1];
                        for (int i = p.mel_ptr[m]; i < e; ++i) {
                            const int fi = int(p.mel_f[i]) - f_lo;
                            const float w = p.mel_w[i];
                            a0 = fmaf(w, mg[fi], a0);
                            a1 = fmaf(w, mg[f_n + fi], a1);
                        }
                    }
                    *reinterpret_cast<float2*>(meltile + m * 32 + slot * 2) = make_float2(a0, a1);
                }
            }
            __syncwarp();
            named_bar_sync(2, kComputeThreads);   // tile complete in shared memory
            // coalesced store [B, n_mel, T, C] + per-clip min/max (data_utils.py:37-47)
            float mn = __int_as_float(0x7f800000), mx = 0.f;
            const int C = p.C;
            for (int e = tid; e < p.n_mel * 32; e += kComputeThreads) {
                const int m = e >> 5, r = e & 31, j = r >> 1, c = r & 1;
                const int tt = tc.t0 + j, ch = 2 * tc.pair + c;
                if (tt < p.T && ch < C) {
                    const float v = meltile[e];
                    p.out[((size_t(b) * p.n_mel + m) * p.T + tt) * C + ch] = v;
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                }
            }
            if (p.minmax != nullptr) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0 && mn <= mx) {
                    atomicMax(&p.minmax[2 * b], ~__float_as_uint(mn));
                    atomicMax(&p.minmax[2 * b + 1], __float_as_uint(mx));
                }
            }
        } else if (MODE == FM_ACTIVITY) {
            // frame "active" iff any STFT coefficient (any bin, re or im, any channel) > 0
            // (pipeline.py:55)
            float mxv = 0.f;
            if (do_fft) {
                auto emit = [&](cpx zf, cpx zm) {
                    mxv = fmaxf(mxv, fmaxf(fmaxf(zf.x + zm.x, zf.y - zm.y),
                                           fmaxf(zf.y + zm.y, zm.x - zf.x)));
                };
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const cpx pa = l0 ? Za[(16 - k2) & 15] : Zb[15 - k2];
                    const cpx pb = l0 ? Zb[15 - k2] : Za[15 - k2];
                    emit(Za[k2], pa);
                    emit(Zb[k2], pb);
                }
                if (l0) emit(Za[8], Za[8]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mxv = fmaxf(mxv, __shfl_xor_sync(hmask, mxv, o));
            if (in_range && l0 && mxv > 0.f) p.activity[size_t(b) * p.T + t] = 1;
        } else {
            if (do_fft) {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const cpx pa = l0 ? Za[(16 - k2) & 15] : Zb[15 - k2];
                    const cpx pb = l0 ? Zb[15 - k2] : Za[15 - k2];
                    {
                        const int f = ka + 32 * k2;
                        const cpx zf = Za[k2];
                        store_bin<MODE>(p, b, f, t, tc.pair, has1, zf.x + pa.x, zf.y - pa.y,
                                        zf.y + pa.y, pa.x - zf.x, mt * freq_mult(f));
                    }
                    {
                        const int f = kb + 32 * k2;
                        const cpx zf = Zb[k2];
                        store_bin<MODE>(p, b, f, t, tc.pair, has1, zf.x + pb.x, zf.y - pb.y,
                                        zf.y + pb.y, pb.x - zf.x, mt * freq_mult(f));
                    }
                }
                if (l0) {
                    const cpx zf = Za[8];
                    store_bin<MODE>(p, b, 256, t, tc.pair, has1, zf.x + zf.x, zf.y - zf.y,
                                    zf.y + zf.y, zf.x - zf.x, mt * freq_mult(256));
                }
            }
        }
    }
}

int fused_max_mel_window() { return kMaxMelBinsWindow; }

cudaError_t launch_fused(const FusedParams& p, int mode, int num_sms, cudaStream_t stream) {
    const int tpc = (p.T + kTF - 1) / kTF;
    const long long n_tiles = (long long)p.B * p.n_pairs * tpc;
    if (n_tiles <= 0) return cudaSuccess;
    const size_t smem = fused_smem_bytes(mode, p.n_mel);
    const int grid = int(n_tiles < 2LL * num_sms ? n_tiles : 2LL * num_sms);
#define IRIS_LAUNCH(M)                                                                          \
    {                                                                                           \
        static bool attr_set = false;                                                           \
        if (!attr_set) {                                                                        \
            cudaError_t e = cudaFuncSetAttribute(k_fused<M>,                                    \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 int(fused_smem_bytes(M, 256)));                \
            if (e != cudaSuccess) return e;                                                     \
            cudaFuncSetAttribute(k_fused<M>, cudaFuncAttributePreferredSharedMemoryCarveout,    \
                                 cudaSharedmemCarveoutMaxShared);                               \
            attr_set = true;                                                                    \
        }                                                                                       \
        k_fused<M><<<grid, kThreads, smem, stream>>>(p);                                        \
    }
    switch (mode) {
        case FM_COMPLEX: IRIS_LAUNCH(FM_COMPLEX) break;
        case FM_MAGPHASE: IRIS_LAUNCH(FM_MAGPHASE) break;
        case FM_LOGMAGPHASE: IRIS_LAUNCH(FM_LOGMAGPHASE) break;
        case FM_MEL: IRIS_LAUNCH(FM_MEL) break;
        case FM_ACTIVITY: IRIS_LAUNCH(FM_ACTIVITY) break;
        default: return cudaErrorInvalidValue;
    }
#undef IRIS_LAUNCH
    return cudaGetLastError();
}

}  // namespace iris
