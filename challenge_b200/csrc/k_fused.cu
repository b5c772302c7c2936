// Fused hot path (v4): gather + time-domain mix of the gained source frames (TMA bulk
// copies into a 2-slot shared-memory ring), window, 512-point FFT per frame (two real
// channels packed into one complex transform, one half-warp per FFT), then the epilogue
// (SpecAugment masks, channel remap, stft_filter, complex / mag-phase / log-mag-phase
// output, or magnitude -> sparse mel -> per-clip min-max -> log), writing each feature once.
//
// Replaces, for one output clip, the chain
//   data_utils.load_wav (STFT, data_utils.py:9-29)  ->  pipeline.merge_complex_specs
//   (pipeline.py:6-110)  ->  data_utils.augment (58-61)  ->  stereo_mono /
//   random_merge_aug / stft_filter (79-136)  ->  transforms.complex_to_magphase
//   (transforms.py:111-123)  ->  magphase_to_mel (51-77)  ->  data_utils.minmax (37-47)
//   ->  data_utils.log_on_mel (50-55)
// using linearity of the STFT: sum_k g_k STFT(src_k)[frame] = FFT(w * sum_k g_k frame_k).
//
// Structure.  A tile is FR = 16/NP output frames x NP channel pairs of one clip (NP = 1 for
// <= 2 channels, 2 for <= 4, ...): 16 half-warp FFT slots, 8 warps, 2 CTAs per SM.
//  * k_tiles (one thread per tile) compacts, for every tile, the mixing segments that are
//    kept and overlap it into a TileBlock (stage descriptors + the tile's mask bits).
//  * k_fused CTAs claim tiles from a global counter (dynamic scheduling) and stream the
//    TileBlocks of the next tiles into an 8-deep shared-memory ring with cp.async, three
//    tiles ahead; every ring entry has an mbarrier that completes when its block has landed.
//  * Every stage of a tile is (FR+1) 2 KB rows per pair, fetched with one cp.async.bulk per
//    pair into slot (stage & 1) and consumed by all 8 warps.  There is no producer warp: the
//    LAST warp to finish reading a slot (shared-memory counter) issues the copy of stage+2
//    into it, so loads run two stages ahead while the FFTs execute.
//  * There is NO CTA-wide barrier in the tile loop: warps are coupled only through the slot
//    protocol (a warp runs at most two stages ahead of the slowest one), so mix, FFT and
//    epilogue phases of different warps overlap.  Every warp stores its own frames.
//  * LOGMEL_MINMAX: per-warp min/max go to global atomics; k_logmel_post (k_post.cu)
//    normalises and logs the batch in place right after, while it is still in L2.  (An
//    in-kernel tail run by the CTA that completes a clip was measured 35 % slower: the CTA
//    that falls behind becomes the last finisher of every clip it touches.)
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

constexpr int kSlots = 16;       // half-warp FFT slots per CTA
constexpr int kWarps = 8;
constexpr int kThreads = 256;
constexpr int kRing = 8;         // TileBlock ring entries per CTA (>= 6: see the skew bound below)
constexpr int kMaxStages = 16;   // mixing segments of one clip (upper bound on stages per tile)
constexpr int kMaxMel = 128;
constexpr int kMaxTaps = 64;     // sum over the 16-filter groups of the longest filter in the group

struct StageDesc {
    const float* src;      // first row of pair plane (group * NP) of the source
    uint32_t pair_stride;  // floats between pair planes
    uint16_t j_lo, j_cnt;  // tile-relative frames [j_lo, j_lo + j_cnt); j_cnt == 0: empty tile
    float gain;
    uint32_t pad_;
};
static_assert(sizeof(StageDesc) == 24, "StageDesc layout");

struct TileBlock {
    int32_t n;             // stages (>= 1); 0 = end of work
    int32_t b;             // clip
    int32_t t0_group;      // first frame | group << 24
    uint32_t tmask_bits;   // bit j: frame t0 + j is time-masked (transforms.py:12-40)
    int16_t fm[8];         // (size, offset) x 4 frequency masks of the clip
    StageDesc d[kMaxStages];
};
static_assert(sizeof(TileBlock) == 32 + 24 * kMaxStages, "TileBlock layout");
constexpr int kTileBlockBytes = int(sizeof(TileBlock));

// ---- shared memory map (bytes) ----
constexpr int OFF_FULL = 0;      // uint64 full[2]      slot s holds a complete stage
constexpr int OFF_CNT = 16;      // int cnt[2]          warps done reading slot s
constexpr int OFF_CLAIM = 32;    // int claimed[3]      first tiles of the CTA
constexpr int OFF_RFULL = 64;    // uint64 rfull[8]     ring entry k & 7 has landed
constexpr int OFF_RING = 128;
constexpr int OFF_TW = OFF_RING + kRing * kTileBlockBytes;
constexpr int OFF_WH = OFF_TW + 256 * 16;
constexpr int OFF_MINFO = OFF_WH + 512 * 4;
constexpr int OFF_MW = OFF_MINFO + kMaxMel * 4;
constexpr int OFF_XCH = OFF_MW + kMaxTaps * 16 * 4;
constexpr int OFF_SLOTS = OFF_XCH + kSlots * kXchSlotFloats * 4;
static_assert(OFF_SLOTS % 128 == 0, "slot alignment");

__host__ __device__ inline uint32_t slot_bytes(int np_shift) {
    return (1u << np_shift) * ((16u >> np_shift) + 1u) * 2048u;
}
__host__ __device__ inline uint32_t smem_total(int np_shift) {
    return OFF_SLOTS + 2 * slot_bytes(np_shift);
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// arrive on an mbarrier once all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- pre-kernel: one thread per tile builds its TileBlock ----
__global__ void __launch_bounds__(128) k_tiles(const FusedParams p) {
    const int FR = 16 >> p.np_shift;
    const int tpc = (p.T + FR - 1) / FR;
    const int per_clip = tpc * p.n_groups;
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= p.B * per_clip) return;
    const int b = tile / per_clip;
    const int r = tile - b * per_clip;
    const int group = r / tpc;
    const int t0 = (r - group * tpc) * FR;
    const int t_end = min(t0 + FR, p.T);
    unsigned char* blk = p.tile_blocks + size_t(tile) * p.tile_stride;
    StageDesc* d = reinterpret_cast<StageDesc*>(blk + 32);
    int n = 0;
    const int s1 = p.seg_ptr[b + 1];
    for (int s = p.seg_ptr[b]; s < s1; ++s) {
        const Seg sg = p.segs[s];
        const int lo = max(sg.t_lo, t0), hi = min(sg.t_hi, t_end);
        if (lo >= hi || (sg.keep_idx >= 0 && p.keep[sg.keep_idx] == 0)) continue;
        if (n < p.max_segs) {
            StageDesc e;
            e.src = sg.base + size_t(group << p.np_shift) * size_t(sg.pair_stride) +
                    size_t(lo + sg.shift) * 512;
            e.pair_stride = uint32_t(sg.pair_stride);
            e.j_lo = uint16_t(lo - t0);
            e.j_cnt = uint16_t(hi - lo);
            e.gain = sg.gain;
            e.pad_ = 0;
            d[n++] = e;
        }
    }
    if (n == 0) {   // nothing overlaps: one empty stage keeps the ring protocol uniform
        StageDesc e;
        e.src = nullptr; e.pair_stride = 0; e.j_lo = 0; e.j_cnt = 0; e.gain = 0.f; e.pad_ = 0;
        d[n++] = e;
    }
    uint32_t tbits = 0;
    if (p.tmask != nullptr) {
        const int32_t* tm = p.tmask + size_t(b) * p.n_tmask * 2;
        for (int i = 0; i < p.n_tmask; ++i) {
            const int size = tm[2 * i], off = tm[2 * i + 1];
            const int lo = max(off, t0), hi = min(off + size, t_end);
            if (lo < hi) tbits |= ((1u << (hi - lo)) - 1u) << (lo - t0);
        }
    }
    int4 hdr;
    hdr.x = n; hdr.y = b; hdr.z = t0 | (group << 24); hdr.w = int(tbits);
    *reinterpret_cast<int4*>(blk) = hdr;
    int16_t fm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) fm[i] = 0;
    if (p.fmask != nullptr) {
        const int32_t* fk = p.fmask + size_t(b) * p.n_fmask * 2;
        for (int i = 0; i < p.n_fmask && i < 4; ++i) {
            fm[2 * i] = int16_t(fk[2 * i]);
            fm[2 * i + 1] = int16_t(fk[2 * i + 1]);
        }
    }
    *reinterpret_cast<int4*>(blk + 16) = *reinterpret_cast<const int4*>(fm);
}

__device__ __forceinline__ void issue_stage(const StageDesc* dp, uint32_t slot_addr, uint32_t full_addr,
                                            int np_here, uint32_t plane_bytes) {
    const StageDesc d = *dp;
    if (d.j_cnt == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_addr) : "memory");
        return;
    }
    const uint32_t bytes = (uint32_t(d.j_cnt) + 1u) * 2048u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_addr),
                 "r"(bytes * uint32_t(np_here))
                 : "memory");
    for (int pr = 0; pr < np_here; ++pr)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(slot_addr + pr * plane_bytes + uint32_t(d.j_lo) * 2048u),
            "l"(d.src + size_t(pr) * d.pair_stride), "r"(bytes), "r"(full_addr)
            : "memory");
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

// KB: number of 32-bin groups the epilogue needs (mel support below bin 32*KB); 8 = all bins.
template <int MODE, int KB>
__global__ void __launch_bounds__(kThreads, 2) k_fused(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr bool kMel = (MODE == FM_MEL);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, n2 = lane & 15;
    const uint32_t sm_base = smem_u32(sm);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + OFF_FULL);
    uint64_t* rfull = reinterpret_cast<uint64_t*>(sm + OFF_RFULL);
    int* cnt = reinterpret_cast<int*>(sm + OFF_CNT);
    int* claimed = reinterpret_cast<int*>(sm + OFF_CLAIM);
    float4* s_tw4 = reinterpret_cast<float4*>(sm + OFF_TW);
    float* s_wh = reinterpret_cast<float*>(sm + OFF_WH);
    uint32_t* s_minfo = reinterpret_cast<uint32_t*>(sm + OFF_MINFO);
    float* s_mw = reinterpret_cast<float*>(sm + OFF_MW);
    auto ring = [&](int k) -> TileBlock* {
        return reinterpret_cast<TileBlock*>(sm + OFF_RING + (k & (kRing - 1)) * kTileBlockBytes);
    };
    const int NP = 1 << p.np_shift;
    const int FR = 16 >> p.np_shift;
    const int n_tiles = p.B * ((p.T + FR - 1) / FR) * p.n_groups;
    const uint32_t slotB = slot_bytes(p.np_shift);
    const uint32_t planeB = uint32_t(FR + 1) * 2048u;

    // one warp streams the TileBlock of `tile` into ring entry k (or writes the end marker);
    // completion is signalled later by fetch_done(k) (32 arrivals on rfull[k & 7])
    auto fetch_block = [&](int k, int tile) {
        TileBlock* dst = ring(k);
        if (tile < n_tiles) {
            const unsigned char* src = p.tile_blocks + size_t(tile) * p.tile_stride;
            const uint32_t d0 = smem_u32(dst);
            for (int c = lane * 16; c < p.tile_stride; c += 32 * 16) cp_async16(d0 + c, src + c);
        } else {
            if (lane == 0) dst->n = 0;
            __threadfence_block();
        }
    };
    auto fetch_done = [&](int k) { cp_async_arrive(&rfull[k & (kRing - 1)]); };

    for (int i = tid; i < 256; i += kThreads) s_tw4[i] = p.tw4[i];
    for (int i = tid; i < 512; i += kThreads) s_wh[i] = p.whalf[i];
    if (kMel) {
        for (int i = tid; i < kMaxMel; i += kThreads) s_minfo[i] = i < p.n_mel ? p.mel_info[i] : 0u;
        for (int i = tid; i < p.mel_taps * 16; i += kThreads) s_mw[i] = p.mel_w[i];
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        for (int i = 0; i < kRing; ++i) mbar_init(&rfull[i], 32);
        cnt[0] = 0;
        cnt[1] = 0;
        fence_mbar_init();
        // the first three tiles of this CTA, in order (an end marker must never precede work)
        const int c = int(atomicAdd(&p.sched[0], 3u));
        for (int i = 0; i < 3; ++i) claimed[i] = c + i;
    }
    __syncthreads();
    if (warp < 3) {
        fetch_block(warp, claimed[warp]);
        fetch_done(warp);
    }

    // ring entries 0 .. ready are known to have landed (per warp; entries land in order)
    int ready = -1;
    auto ensure = [&](int k) {
        while (ready < k) {
            ++ready;
            mbar_wait(&rfull[ready & (kRing - 1)], uint32_t(ready >> 3) & 1u);
        }
    };
    // issue cursor = position (tile iteration ki, entry ei) of stage q + 2
    int ki = 0, ei = 0;
    auto advance = [&]() {
        ensure(ki);
        const int n = ring(ki)->n;
        if (n != 0 && ++ei >= n) { ++ki; ei = 0; }
    };
    auto np_of = [&](int k) { return min(NP, p.n_pairs - ((ring(k)->t0_group >> 24) << p.np_shift)); };
    ensure(1);
    if (tid == 0) {
        int a = 0, e = 0;
        for (int s = 0; s < 2; ++s) {
            const TileBlock* tb = ring(a);
            if (tb->n == 0) break;
            issue_stage(&tb->d[e], sm_base + OFF_SLOTS + s * slotB, sm_base + OFF_FULL + 8 * s, np_of(a), planeB);
            if (++e >= tb->n) { ++a; e = 0; }
        }
    }
    advance();
    advance();

    const int slot = warp * 2 + (lane >> 4);
    const unsigned hmask = (lane >> 4) ? 0xFFFF0000u : 0x0000FFFFu;
    float* xs = reinterpret_cast<float*>(sm + OFF_XCH) + slot * kXchSlotFloats;
    const int ka = n2, kb = (n2 == 0) ? 16 : 32 - n2;
    const bool l0 = (n2 == 0);
    const int j = slot >> p.np_shift;          // tile-relative frame of this slot
    const int pr = slot & (NP - 1);            // pair within the tile's group
    const float2* my_rows = reinterpret_cast<const float2*>(sm + OFF_SLOTS + pr * planeB) + j * 256 + n2;
    // running per-clip extrema of this lane (flushed when the clip changes)
    float mn = __int_as_float(0x7f800000), mx = 0.f;
    int mm_clip = -1;
    auto flush_minmax = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0 && mn <= mx) {
            atomicMax(&p.minmax[2 * mm_clip], ~__float_as_uint(mn));
            atomicMax(&p.minmax[2 * mm_clip + 1], __float_as_uint(mx));
        }
        mn = __int_as_float(0x7f800000);
        mx = 0.f;
    };

    int q = 0;          // stages consumed so far
    for (int k = 0;; ++k) {
        ensure(k);
        const TileBlock* tb = ring(k);
        const int n_st = tb->n;
        if (n_st == 0) break;
        // Skew bound: a slot is refilled only after all 8 warps have read it and every tile
        // has at least one stage, so no warp is more than two tiles ahead of the slowest one;
        // entry k+3 therefore never overwrites an entry (>= k-2) that is still being read.
        int new_claim = 0;
        // one fixed builder warp: its claims are then ordered like the ring entries (an end
        // marker must never precede a valid tile)
        const bool builder = warp == 0;
        // the builder warp of this iteration claims the tile of ring entry k+3 now and streams
        // its block at the end of the iteration, when the atomic has long returned
        if (builder && lane == 0) new_claim = int(atomicAdd(&p.sched[0], 1u));

        // ---- gather + mix: acc = sum over the stages of this tile of gain * frame ----
        float re[32], im[32];
        const int group = tb->t0_group >> 24;
        const bool pair_ok = (group << p.np_shift) + pr < p.n_pairs;
        // wait for stage e of this tile; returns the gain, or sets active = false
        auto stage_begin = [&](int e, bool& active) -> float {
            const uint32_t jj = *reinterpret_cast<const uint32_t*>(&tb->d[e].j_lo);   // j_lo | j_cnt << 16
            const float gn = tb->d[e].gain;
            mbar_wait(&full[q & 1], uint32_t(q >> 1) & 1u);
            active = pair_ok && unsigned(j - int(jj & 0xffffu)) < (jj >> 16);
            return gn;
        };
        // this warp is done reading the slot; the last warp to say so refills it
        auto stage_end = [&]() {
            const int s = q & 1;
            __syncwarp();
            if (lane == 0) {
                const int old = atomicAdd(&cnt[s], 1);
                if (old == kWarps - 1) {
                    cnt[s] = 0;
                    if (ready < ki) mbar_wait(&rfull[ki & (kRing - 1)], uint32_t(ki >> 3) & 1u);
                    const TileBlock* nb = ring(ki);
                    if (nb->n != 0) {
                        fence_proxy_async();
                        issue_stage(&nb->d[ei], sm_base + OFF_SLOTS + s * slotB, sm_base + OFF_FULL + 8 * s,
                                    np_of(ki), planeB);
                    }
                }
            }
            ++q;
            advance();
        };
        {   // first stage initialises the accumulators
            bool active;
            const float gn = stage_begin(0, active);
            const float2* src = reinterpret_cast<const float2*>(
                reinterpret_cast<const unsigned char*>(my_rows) + (q & 1) * slotB);
            const float g0 = active ? gn : 0.f;
            if (!active) src = reinterpret_cast<const float2*>(sm + OFF_TW);   // 4 KB of finite data
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float2 x = src[16 * i];
                re[i] = g0 * x.x;
                im[i] = g0 * x.y;
            }
            stage_end();
        }
        for (int e = 1; e < n_st; ++e) {
            bool active;
            const float gn = stage_begin(e, active);
            if (active) {
                const float2* src = reinterpret_cast<const float2*>(
                    reinterpret_cast<const unsigned char*>(my_rows) + (q & 1) * slotB);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float2 x = src[16 * i];
                    re[i] = fmaf(gn, x.x, re[i]);
                    im[i] = fmaf(gn, x.y, im[i]);
                }
            }
            stage_end();
        }

        const int b = tb->b;
        const int t0 = tb->t0_group & 0xffffff;
        const int t = t0 + j;
        const int pair = (group << p.np_shift) + pr;
        const bool in_range = t < p.T && pair_ok;
        const bool has1 = (2 * pair + 1 < p.C);
        // SpecAugment time mask of this frame (transforms.py:12-40), precomputed per tile
        const float mt = ((tb->tmask_bits >> j) & 1u) ? 0.f : 1.f;
        // a fully time-masked frame has zero magnitude everywhere: no FFT needed for mel
        const bool do_fft = in_range && !(kMel && mt == 0.f);

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
            for (int m = 0; m < 16; ++m) {
                const float4 w = s_tw4[m * 16 + n2];
                if (m > 0) v[2 * m] = cmul(v[2 * m], cpx{w.x, w.y});
                v[2 * m + 1] = cmul(v[2 * m + 1], cpx{w.z, w.w});
            }
            // ---- exchange through shared memory, 4 rounds of 8 k1: Za (k1 = ka < 16) is
            // served by rounds 0-1, Zb (k1 = kb >= 16) by rounds 2-3 ----
#pragma unroll
            for (int rho = 0; rho < 4; ++rho) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    *reinterpret_cast<float4*>(xs + xch_write_off(a, n2)) =
                        make_float4(v[8 * rho + 2 * a].x, v[8 * rho + 2 * a].y,
                                    v[8 * rho + 2 * a + 1].x, v[8 * rho + 2 * a + 1].y);
                __syncwarp(hmask);
                if (rho < 2) {
                    if ((ka >> 3) == rho) {
#pragma unroll
                        for (int jx = 0; jx < 16; ++jx) {
                            const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(ka & 7, jx));
                            Za[jx] = cpx{z.x, z.y};
                        }
                    }
                } else {
                    if ((kb >> 3) == rho) {
#pragma unroll
                        for (int jx = 0; jx < 16; ++jx) {
                            const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(kb & 7, jx));
                            Zb[jx] = cpx{z.x, z.y};
                        }
                    }
                }
                __syncwarp(hmask);
                // Za is complete after round 1: transform it now, while only v[16..31] is
                // still live, instead of holding v, Za and Zb together
                if (rho == 1) Fft<16>::run(Za);
            }
            Fft<16>::run(Zb);
        }

        // ---- epilogue ----
        // bin f = ka+32*k2 pairs with its mirror 512-f held in Zb[15-k2] (lane 0: Za[(16-k2)&15]);
        // bin f = kb+32*k2 pairs with Za[15-k2] (lane 0: Zb[15-k2]).  The 0.5 of the
        // two-channel split is folded into the window table.
        if (kMel) {
            if (p.do_minmax && b != mm_clip) {
                if (mm_clip >= 0) flush_minmax();
                mm_clip = b;
            }
            float2* mg = reinterpret_cast<float2*>(xs);   // [mel_f_n] (|ch0|, |ch1|), aliases the exchange slot
            const int f_lo = p.mel_f_lo, f_n = p.mel_f_n;
            float acc0[8], acc1[8];   // mel bins m = n2 + 16 r
#pragma unroll
            for (int r = 0; r < 8; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
            if (do_fft) {
                auto emit = [&](int f, cpx zf, cpx zm) {
                    const unsigned fi = unsigned(f - f_lo);
                    if (fi < unsigned(f_n)) {
                        const float r0 = zf.x + zm.x, i0 = zf.y - zm.y;
                        const float r1 = zf.y + zm.y, i1 = zm.x - zf.x;
                        mg[fi] = make_float2(sqrt_approx(fmaf(r0, r0, i0 * i0)),
                                             sqrt_approx(fmaf(r1, r1, i1 * i1)));
                    }
                };
#pragma unroll
                for (int k2 = 0; k2 < KB; ++k2) {
                    const cpx pa = l0 ? Za[(16 - k2) & 15] : Zb[15 - k2];
                    const cpx pb = l0 ? Zb[15 - k2] : Za[15 - k2];
                    emit(ka + 32 * k2, Za[k2], pa);
                    emit(kb + 32 * k2, Zb[k2], pb);
                }
                if (KB == 8 && l0) emit(256, Za[8], Za[8]);
                __syncwarp(hmask);
                // frequency masks (transforms.py:12-40) and stft_filter (data_utils.py:126-136)
                // zero whole bins: |x| * 0 == +0
                if (p.n_fmask > 0 || p.filter_k > 0) {
                    for (int i = 0; i <= p.n_fmask; ++i) {
                        int off, size;
                        if (i < p.n_fmask) { size = tb->fm[2 * i]; off = tb->fm[2 * i + 1]; }
                        else { off = 1; size = p.filter_k; }
                        for (int f = off + n2; f < off + size; f += 16) {
                            const unsigned fi = unsigned(f - f_lo);
                            if (fi < unsigned(f_n)) mg[fi] = make_float2(0.f, 0.f);
                        }
                    }
                    __syncwarp(hmask);
                }
                // sparse mel projection (transforms.py:51-77): filter m = n2 + 16 r reads
                // mel_L[r] taps starting at its first bin; shorter filters are zero-padded, so
                // the trip count is uniform across the half-warp
                const float* wr = s_mw + n2;
#define IRIS_TAP(i)                                            \
    {                                                          \
        const float2 x = a[i];                                 \
        const float w = wr[16 * (i)];                          \
        acc0[r] = fmaf(w, x.x, acc0[r]);                       \
        acc1[r] = fmaf(w, x.y, acc1[r]);                       \
    }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int L = p.mel_L[r];   // 0 for r >= ceil(n_mel / 16); uniform
                    const float2* a = mg + s_minfo[n2 + 16 * r];
                    switch (L) {   // one uniform jump, then straight-line taps
                        case 12: IRIS_TAP(11)
                        case 11: IRIS_TAP(10)
                        case 10: IRIS_TAP(9)
                        case 9: IRIS_TAP(8)
                        case 8: IRIS_TAP(7)
                        case 7: IRIS_TAP(6)
                        case 6: IRIS_TAP(5)
                        case 5: IRIS_TAP(4)
                        case 4: IRIS_TAP(3)
                        case 3: IRIS_TAP(2)
                        case 2: IRIS_TAP(1)
                        case 1: IRIS_TAP(0)
                        default: break;
                    }
                    wr += 16 * L;
                }
#undef IRIS_TAP
            }
            // ---- every lane stores its own mel values: out[b, m, t, 2*pair .. +1] ----
            if (in_range) {
                const int C = p.C;
                float* o = p.out + (size_t(b) * p.n_mel * p.T + t) * C + 2 * pair + size_t(n2) * p.T * C;
                const int rs16 = 16 * p.T * C;
                const bool lg = p.do_log && !p.do_minmax;
                const int n_r = (p.n_mel + 15) >> 4;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r >= n_r) break;                                  // uniform
                    if (r == n_r - 1 && n2 + 16 * r >= p.n_mel) break;    // ragged last group
                    float a0 = acc0[r], a1 = acc1[r];
                    if (p.do_minmax) {
                        mn = fminf(mn, has1 ? fminf(a0, a1) : a0);
                        mx = fmaxf(mx, has1 ? fmaxf(a0, a1) : a0);
                    }
                    if (lg) {
                        a0 = __logf(a0 + 1e-8f);
                        a1 = __logf(a1 + 1e-8f);
                    }
                    if ((C & 1) == 0) {
                        *reinterpret_cast<float2*>(o) = make_float2(a0, a1);
                    } else {
                        o[0] = a0;
                        if (has1) o[1] = a1;
                    }
                    o += rs16;
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
        } else if (do_fft) {
            // per-lane bitmap of frequency-masked bins: bit k2 -> ka + 32*k2,
            // bit 8 + k2 -> kb + 32*k2, bit 16 -> bin 256
            uint32_t zbits = 0;
            for (int i = 0; i < p.n_fmask; ++i) {
                const int size = tb->fm[2 * i], off = tb->fm[2 * i + 1];
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    if (unsigned(ka + 32 * k2 - off) < unsigned(size)) zbits |= 1u << k2;
                    if (unsigned(kb + 32 * k2 - off) < unsigned(size)) zbits |= 1u << (8 + k2);
                }
                if (unsigned(256 - off) < unsigned(size)) zbits |= 1u << 16;
            }
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const cpx pa = l0 ? Za[(16 - k2) & 15] : Zb[15 - k2];
                const cpx pb = l0 ? Zb[15 - k2] : Za[15 - k2];
                {
                    const int f = ka + 32 * k2;
                    const cpx zf = Za[k2];
                    store_bin<MODE>(p, b, f, t, pair, has1, zf.x + pa.x, zf.y - pa.y,
                                    zf.y + pa.y, pa.x - zf.x, ((zbits >> k2) & 1u) ? 0.f : mt);
                }
                {
                    const int f = kb + 32 * k2;
                    const cpx zf = Zb[k2];
                    store_bin<MODE>(p, b, f, t, pair, has1, zf.x + pb.x, zf.y - pb.y,
                                    zf.y + pb.y, pb.x - zf.x, ((zbits >> (8 + k2)) & 1u) ? 0.f : mt);
                }
            }
            if (l0) {
                const cpx zf = Za[8];
                store_bin<MODE>(p, b, 256, t, pair, has1, zf.x + zf.x, zf.y - zf.y,
                                zf.y + zf.y, zf.x - zf.x, ((zbits >> 16) & 1u) ? 0.f : mt);
            }
        }
        if (builder) {
            fetch_block(k + 3, __shfl_sync(0xffffffffu, new_claim, 0));
            fetch_done(k + 3);
        }
#ifndef IRIS_NO_TILE_BARRIER
        // Not needed for correctness: keeps the 8 warps in the same phase, so that the slots are
        // drained (and refilled) at the start of a tile and the loads fly during the FFTs.
        // Measured on cfg2: 357 us with, 376 us without (COMPLEX: 457 vs 639 us).
        __syncthreads();
#endif
    }
    if (kMel && p.do_minmax && mm_clip >= 0) flush_minmax();
    // the last CTA to leave resets the tile scheduler for the next launch
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(&p.sched[1], 1u) == gridDim.x - 1) {
            p.sched[0] = 0u;
            p.sched[1] = 0u;
        }
    }
}

size_t fused_smem_bytes(const FusedParams& p, int) { return smem_total(p.np_shift); }
int fused_max_segments() { return kMaxStages; }
int fused_max_mel_window() { return kXchSlotFloats / 2; }
int fused_max_mel_taps() { return kMaxTaps; }
size_t fused_tile_bytes(const FusedParams& p, int* stride_out) {
    const int FR = 16 >> p.np_shift;
    const long long n_tiles = (long long)p.B * p.n_groups * ((p.T + FR - 1) / FR);
    int ms = p.max_segs < 1 ? 1 : p.max_segs;
    const int stride = (32 + 24 * ms + 15) & ~15;
    if (stride_out) *stride_out = stride;
    return size_t(n_tiles) * size_t(stride);
}

// p.tile_blocks (fused_tile_bytes) and p.sched (2 x uint32, zero) are provided by the caller.
cudaError_t launch_fused(const FusedParams& p, int mode, int num_sms, cudaStream_t stream) {
    const int FR = 16 >> p.np_shift;
    const int tpc = (p.T + FR - 1) / FR;
    const long long n_tiles = (long long)p.B * p.n_groups * tpc;
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tiles > 0x7fffffffLL || p.max_segs > kMaxStages) return cudaErrorInvalidValue;
    const size_t smem = fused_smem_bytes(p, mode);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    k_tiles<<<unsigned((n_tiles + 127) / 128), 128, 0, stream>>>(p);
    const int grid = int(n_tiles < 2LL * num_sms ? n_tiles : 2LL * num_sms);
#define IRIS_LAUNCH(M, KBV)                                                                     \
    {                                                                                           \
        static bool attr_set = false;                                                           \
        if (!attr_set) {                                                                        \
            cudaError_t e = cudaFuncSetAttribute(k_fused<M, KBV>,                               \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 227 * 1024);                                   \
            if (e != cudaSuccess) return e;                                                     \
            cudaFuncSetAttribute(k_fused<M, KBV>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                 cudaSharedmemCarveoutMaxShared);                               \
            attr_set = true;                                                                    \
        }                                                                                       \
        k_fused<M, KBV><<<grid, kThreads, smem, stream>>>(p);                                   \
    }
    switch (mode) {
        case FM_COMPLEX: IRIS_LAUNCH(FM_COMPLEX, 8) break;
        case FM_MAGPHASE: IRIS_LAUNCH(FM_MAGPHASE, 8) break;
        case FM_LOGMAGPHASE: IRIS_LAUNCH(FM_LOGMAGPHASE, 8) break;
        case FM_MEL:
            if (p.mel_f_lo + p.mel_f_n <= 128) IRIS_LAUNCH(FM_MEL, 4)
            else IRIS_LAUNCH(FM_MEL, 8)
            break;
        case FM_ACTIVITY: IRIS_LAUNCH(FM_ACTIVITY, 8) break;
        default: return cudaErrorInvalidValue;
    }
#undef IRIS_LAUNCH
    return cudaGetLastError();
}

}  // namespace iris
