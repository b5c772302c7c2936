// Fused hot path (v2): gather + time-domain mix of the gained source frames (TMA bulk
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
// <= 2 channels, 2 for <= 4, ...): 16 half-warp FFT slots, 8 warps, 2 CTAs per SM.  Every
// mixing segment that overlaps a tile is one STAGE: (FR+1) 2 KB rows per pair, fetched with
// one cp.async.bulk per pair into slot (stage & 1) and consumed by all 8 warps.  There is no
// producer warp: the LAST warp to finish reading a slot (shared-memory counter) issues the
// copy of stage+2 into it, so loads run two stages ahead while the FFTs execute.  Stage lists
// of the next tiles are compacted by one warp three tiles ahead (keep flags, overlap tests)
// into a 4-deep ring in shared memory.
#include "iris_common.cuh"
#include "iris_launch.h"

namespace iris {

constexpr int kSlots = 16;       // half-warp FFT slots per CTA
constexpr int kWarps = 8;
constexpr int kThreads = 256;
constexpr int kRing = 4;         // tile stage lists in flight
constexpr int kMelPad = 34;      // mel tile row stride in floats (conflict-free float2 columns)
constexpr int kMaxStages = 32;   // mixing segments of one clip (upper bound on stages per tile)

struct StageDesc {
    const float* src;      // first row of pair plane (group * NP) of the source
    uint32_t pair_stride;  // floats between pair planes
    uint16_t j_lo, j_cnt;  // tile-relative frames [j_lo, j_lo + j_cnt); j_cnt == 0: empty tile
    float gain;
    uint32_t pad_;
};
static_assert(sizeof(StageDesc) == 24, "StageDesc layout");

struct TileHdr {
    int32_t n, b, group, t0;
};

struct Layout {
    uint32_t slot_floats, plane_floats, xch_floats;
    uint32_t off_slots, off_xch, off_mel, off_tw, off_wh, off_minfo, off_mw, off_lists, off_hdr,
        off_bars, total;
};

__host__ __device__ inline Layout make_layout(int np_shift, bool mel, int n_mel, int mel_f_n,
                                              int mel_nw) {
    Layout L;
    const uint32_t NP = 1u << np_shift, FR = 16u >> np_shift;
    L.plane_floats = (FR + 1) * 512;
    L.slot_floats = NP * L.plane_floats;
    L.xch_floats = kXchSlotFloats;
    if (mel && uint32_t(2 * mel_f_n) > L.xch_floats) L.xch_floats = (2 * mel_f_n + 3) & ~3u;
    uint32_t o = 0;
    auto take = [&o](uint32_t bytes) { const uint32_t at = o; o += (bytes + 15u) & ~15u; return at; };
    L.off_slots = take(2 * L.slot_floats * 4);
    L.off_xch = take(kSlots * L.xch_floats * 4);
    L.off_mel = take(mel ? uint32_t(n_mel) * kMelPad * 4 : 0);
    L.off_tw = take(256 * 16);
    L.off_wh = take(512 * 4);
    L.off_minfo = take(mel ? uint32_t(n_mel) * 4 : 0);
    L.off_mw = take(mel ? uint32_t(mel_nw) * 4 : 0);
    L.off_lists = take(kRing * kMaxStages * uint32_t(sizeof(StageDesc)));
    L.off_hdr = take(kRing * uint32_t(sizeof(TileHdr)));
    L.off_bars = take(64);
    L.total = o;
    return L;
}

struct TileGeom {
    int FR, tpc, per_clip, n_tiles;
};
__device__ __forceinline__ TileGeom tile_geom(const FusedParams& p) {
    TileGeom g;
    g.FR = 16 >> p.np_shift;
    g.tpc = (p.T + g.FR - 1) / g.FR;
    g.per_clip = g.tpc * p.n_groups;
    g.n_tiles = p.B * g.per_clip;
    return g;
}

// One warp compacts the stages of `tile` (mixing segments that are kept and overlap it).
__device__ __forceinline__ void build_list(const FusedParams& p, const TileGeom& g, int tile,
                                           TileHdr* hdr, StageDesc* list, int lane) {
    const int b = tile / g.per_clip;
    const int r = tile - b * g.per_clip;
    const int group = r / g.tpc;
    const int t0 = (r - group * g.tpc) * g.FR;
    const int t_end = min(t0 + g.FR, p.T);
    const int s1 = p.seg_ptr[b + 1];
    int n = 0;
    for (int base = p.seg_ptr[b]; base < s1; base += 32) {
        const int s = base + lane;
        bool valid = false;
        Seg sg;
        int lo = 0, hi = 0;
        if (s < s1) {
            sg = p.segs[s];
            lo = max(sg.t_lo, t0);
            hi = min(sg.t_hi, t_end);
            valid = lo < hi && !(sg.keep_idx >= 0 && p.keep[sg.keep_idx] == 0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int pos = n + __popc(bal & ((1u << lane) - 1u));
            if (pos < kMaxStages) {
                StageDesc d;
                d.src = sg.base + size_t(group << p.np_shift) * size_t(sg.pair_stride) +
                        size_t(lo + sg.shift) * 512;
                d.pair_stride = uint32_t(sg.pair_stride);
                d.j_lo = uint16_t(lo - t0);
                d.j_cnt = uint16_t(hi - lo);
                d.gain = sg.gain;
                d.pad_ = 0;
                list[pos] = d;
            }
        }
        n += __popc(bal);
    }
    if (lane == 0) {
        if (n == 0) {   // nothing overlaps: one empty stage keeps the ring protocol uniform
            StageDesc d;
            d.src = nullptr; d.pair_stride = 0; d.j_lo = 0; d.j_cnt = 0; d.gain = 0.f; d.pad_ = 0;
            list[0] = d;
            n = 1;
        }
        TileHdr h;
        h.n = min(n, kMaxStages); h.b = b; h.group = group; h.t0 = t0;
        *hdr = h;
    }
    __syncwarp();
}

__device__ __forceinline__ void issue_stage(const StageDesc& d, float* slot, uint64_t* full,
                                            int np_here, uint32_t plane_floats) {
    if (d.j_cnt == 0) {
        mbar_arrive(full);
        return;
    }
    const uint32_t bytes = (uint32_t(d.j_cnt) + 1u) * 2048u;
    mbar_arrive_expect_tx(full, bytes * uint32_t(np_here));
    for (int pr = 0; pr < np_here; ++pr)
        bulk_g2s(slot + pr * plane_floats + uint32_t(d.j_lo) * 512u,
                 d.src + size_t(pr) * d.pair_stride, bytes, full);
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
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool kMel = (MODE == FM_MEL);
    const Layout L = make_layout(p.np_shift, kMel, p.n_mel, p.mel_f_n, p.mel_nw);
    float* slots = reinterpret_cast<float*>(smem_raw + L.off_slots);
    float* xch = reinterpret_cast<float*>(smem_raw + L.off_xch);
    float* meltile = reinterpret_cast<float*>(smem_raw + L.off_mel);
    float4* s_tw4 = reinterpret_cast<float4*>(smem_raw + L.off_tw);
    float* s_wh = reinterpret_cast<float*>(smem_raw + L.off_wh);
    uint32_t* s_minfo = reinterpret_cast<uint32_t*>(smem_raw + L.off_minfo);
    float* s_mw = reinterpret_cast<float*>(smem_raw + L.off_mw);
    StageDesc* lists = reinterpret_cast<StageDesc*>(smem_raw + L.off_lists);
    TileHdr* hdrs = reinterpret_cast<TileHdr*>(smem_raw + L.off_hdr);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L.off_bars);   // [2]
    int* cnt = reinterpret_cast<int*>(full + 2);                           // [2]
    int* s_flag = cnt + 2;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, hw = lane >> 4, n2 = lane & 15;
    const TileGeom g = tile_geom(p);
    const int n_my = (g.n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    const int NP = 1 << p.np_shift;

    for (int i = tid; i < 256; i += kThreads) s_tw4[i] = p.tw4[i];
    for (int i = tid; i < 512; i += kThreads) s_wh[i] = p.whalf[i];
    if (kMel) {
        for (int i = tid; i < p.n_mel; i += kThreads) s_minfo[i] = p.mel_info[i];
        for (int i = tid; i < p.mel_nw; i += kThreads) s_mw[i] = p.mel_w[i];
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        cnt[0] = 0;
        cnt[1] = 0;
        *s_flag = 0;
        fence_mbar_init();
    }
    if (warp < 3 && warp < n_my)
        build_list(p, g, int(blockIdx.x) + warp * int(gridDim.x), &hdrs[warp],
                   lists + warp * kMaxStages, lane);
    __syncthreads();

    // issue cursor = position (tile iteration ki, entry ei) of stage q + 2
    int ki = 0, ei = 0;
    auto advance = [&]() {
        if (ki < n_my) {
            if (++ei >= hdrs[ki & (kRing - 1)].n) { ++ki; ei = 0; }
        }
    };
    auto np_of = [&](int k) { return min(NP, p.n_pairs - (hdrs[k & (kRing - 1)].group << p.np_shift)); };
    if (tid == 0) {
        int a = 0, e = 0;
        for (int s = 0; s < 2 && a < n_my; ++s) {
            issue_stage(lists[(a & (kRing - 1)) * kMaxStages + e], slots + s * L.slot_floats,
                        &full[s], np_of(a), L.plane_floats);
            if (++e >= hdrs[a & (kRing - 1)].n) { ++a; e = 0; }
        }
    }
    advance();
    advance();

    const int slot = warp * 2 + hw;
    const unsigned hmask = hw ? 0xFFFF0000u : 0x0000FFFFu;
    float* xs = xch + slot * L.xch_floats;
    const int ka = n2, kb = (n2 == 0) ? 16 : 32 - n2;
    const bool l0 = (n2 == 0);
    const int j = slot >> p.np_shift;          // tile-relative frame of this slot
    const int pr = slot & (NP - 1);            // pair within the tile's group
    const float2* my_rows = reinterpret_cast<const float2*>(slots + pr * L.plane_floats) + j * 256 + n2;
    const int slot_f2 = int(L.slot_floats >> 1);

    int q = 0;   // stages consumed so far
    for (int k = 0; k < n_my; ++k) {
        if (warp == (k & (kWarps - 1)) && k + 3 < n_my)
            build_list(p, g, int(blockIdx.x) + (k + 3) * int(gridDim.x), &hdrs[(k + 3) & (kRing - 1)],
                       lists + ((k + 3) & (kRing - 1)) * kMaxStages, lane);
        const TileHdr h = hdrs[k & (kRing - 1)];
        const StageDesc* list = lists + (k & (kRing - 1)) * kMaxStages;
        const int b = h.b;
        const int t = h.t0 + j;
        const int pair = (h.group << p.np_shift) + pr;
        const bool pair_ok = pair < p.n_pairs;
        const bool in_range = t < p.T && pair_ok;
        const bool has1 = (2 * pair + 1 < p.C);

        float re[32], im[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { re[i] = 0.f; im[i] = 0.f; }

        // ---- gather + mix: acc += gain * frame_k of every stage of this tile ----
        for (int e = 0; e < h.n; ++e) {
            const StageDesc d = list[e];
            const int s = q & 1;
            mbar_wait(&full[s], uint32_t(q >> 1) & 1u);
            if (pair_ok && j >= int(d.j_lo) && j < int(d.j_lo) + int(d.j_cnt)) {
                const float gn = d.gain;
                const float2* src = my_rows + s * slot_f2;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float2 x = src[16 * i];
                    re[i] = fmaf(gn, x.x, re[i]);
                    im[i] = fmaf(gn, x.y, im[i]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                const int old = atomicAdd(&cnt[s], 1);
                if (old == kWarps - 1) {          // last reader of the slot: refill it
                    cnt[s] = 0;
                    if (ki < n_my) {
                        fence_proxy_async();
                        issue_stage(lists[(ki & (kRing - 1)) * kMaxStages + ei],
                                    slots + s * L.slot_floats, &full[s], np_of(ki), L.plane_floats);
                    }
                }
            }
            ++q;
            advance();
        }

        // ---- SpecAugment time mask of this frame (transforms.py:12-40) ----
        float mt = 1.f;
        if (p.tmask != nullptr) {
            const int32_t* tm = p.tmask + size_t(b) * p.n_tmask * 2;
            bool hit = false;
            for (int i0 = 0; i0 < p.n_tmask; i0 += 16) {
                const int i = i0 + n2;
                if (i < p.n_tmask) {
                    const int2 so = *reinterpret_cast<const int2*>(tm + 2 * i);
                    hit = hit || (t >= so.y && t < so.y + so.x);
                }
            }
            if (__ballot_sync(hmask, hit) & hmask) mt = 0.f;
        }

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
                        for (int jj = 0; jj < 16; ++jj) {
                            const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(ka & 7, jj));
                            Za[jj] = cpx{z.x, z.y};
                        }
                    }
                } else {
                    if ((kb >> 3) == rho) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const float2 z = *reinterpret_cast<const float2*>(xs + xch_read_off(kb & 7, jj));
                            Zb[jj] = cpx{z.x, z.y};
                        }
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
        if (kMel) {
            float2* mg = reinterpret_cast<float2*>(xs);   // [mel_f_n] (|ch0|, |ch1|), aliases the exchange slot
            const int f_lo = p.mel_f_lo, f_n = p.mel_f_n;
            float acc0[8], acc1[8];   // mel bins m = n2 + 16 r
            const int n_r = (p.n_mel + 15) >> 4;
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
                if (p.fmask != nullptr || p.filter_k > 0) {
                    const int32_t* fmk = p.fmask + size_t(b) * p.n_fmask * 2;
                    const int n_z = (p.fmask != nullptr ? p.n_fmask : 0) + (p.filter_k > 0 ? 1 : 0);
                    for (int i = 0; i < n_z; ++i) {
                        int off, size;
                        if (p.fmask != nullptr && i < p.n_fmask) { size = fmk[2 * i]; off = fmk[2 * i + 1]; }
                        else { off = 1; size = p.filter_k; }
                        for (int f = off + n2; f < off + size; f += 16) {
                            const unsigned fi = unsigned(f - f_lo);
                            if (fi < unsigned(f_n)) mg[fi] = make_float2(0.f, 0.f);
                        }
                    }
                    __syncwarp(hmask);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    acc0[r] = 0.f; acc1[r] = 0.f;
                    const int m = n2 + 16 * r;
                    if (r < n_r && m < p.n_mel) {
                        const uint32_t info = s_minfo[m];
                        const int start = int(info & 511u) - f_lo, len = int((info >> 9) & 511u);
                        const float* w = s_mw + (info >> 18);
                        for (int i = 0; i < len; ++i) {
                            const float2 a = mg[start + i];
                            acc0[r] = fmaf(w[i], a.x, acc0[r]);
                            acc1[r] = fmaf(w[i], a.y, acc1[r]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
            }
            __syncwarp();
            // ---- tile of mel values [n_mel][FR frames x C channels] in shared memory ----
            const int C = p.C;
            if (in_range) {
                const int col = j * C + 2 * pr;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int m = n2 + 16 * r;
                    if (r < n_r && m < p.n_mel) {
                        if ((C & 1) == 0) {
                            *reinterpret_cast<float2*>(meltile + m * kMelPad + col) = make_float2(acc0[r], acc1[r]);
                        } else {
                            meltile[m * kMelPad + col] = acc0[r];
                            if (has1) meltile[m * kMelPad + col + 1] = acc1[r];
                        }
                    }
                }
            }
            __syncthreads();   // tile complete (also publishes the stage list built this iteration)
            // coalesced store [B, n_mel, T, C] (+ log) and per-clip min/max (data_utils.py:37-55)
            const int fr_valid = min(g.FR, p.T - h.t0);
            const int ncols = fr_valid * C;
            float mn = __int_as_float(0x7f800000), mx = 0.f;
            float* orow = p.out + (size_t(b) * p.n_mel * p.T + h.t0) * C;
            if ((C & 1) == 0) {
                const int c2 = tid & 15;
                if (2 * c2 < ncols) {
                    for (int m = tid >> 4; m < p.n_mel; m += kThreads / 16) {
                        float2 vv = *reinterpret_cast<const float2*>(meltile + m * kMelPad + 2 * c2);
                        mn = fminf(mn, fminf(vv.x, vv.y));
                        mx = fmaxf(mx, fmaxf(vv.x, vv.y));
                        if (p.do_log && !p.do_minmax) {
                            vv.x = __logf(vv.x + 1e-8f);
                            vv.y = __logf(vv.y + 1e-8f);
                        }
                        *reinterpret_cast<float2*>(orow + size_t(m) * p.T * C + 2 * c2) = vv;
                    }
                }
            } else {
                const int c1 = tid & 31;
                if (c1 < ncols) {
                    for (int m = tid >> 5; m < p.n_mel; m += kThreads / 32) {
                        float vv = meltile[m * kMelPad + c1];
                        mn = fminf(mn, vv);
                        mx = fmaxf(mx, vv);
                        if (p.do_log && !p.do_minmax) vv = __logf(vv + 1e-8f);
                        orow[size_t(m) * p.T * C + c1] = vv;
                    }
                }
            }
            if (p.do_minmax) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (lane == 0 && mn <= mx) {
                    atomicMax(&p.minmax[2 * b], ~__float_as_uint(mn));
                    atomicMax(&p.minmax[2 * b + 1], __float_as_uint(mx));
                }
                __threadfence();
            }
            __syncthreads();   // store phase over: the mel tile may be overwritten
            if (p.do_minmax) {
                // the CTA that completes a clip normalises it in place while it is still in L2:
                // (x - min) / max(max - min, 1e-8), then log(x + 1e-8)   (data_utils.py:37-55)
                if (tid == 0) {
                    const unsigned old = atomicAdd(&p.clip_done[b], 1u);
                    *s_flag = (old == unsigned(g.per_clip) - 1u) ? 1 : 0;
                    __threadfence();
                }
                __syncthreads();
                if (*s_flag) {
                    const float lo = __uint_as_float(~__ldcg(&p.minmax[2 * b]));
                    const float hi = __uint_as_float(__ldcg(&p.minmax[2 * b + 1]));
                    const float den = fmaxf(hi - lo, 1e-8f);
                    const size_t per = size_t(p.n_mel) * p.T * C;
                    float* base = p.out + size_t(b) * per;
                    if ((per & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
                        float4* v4 = reinterpret_cast<float4*>(base);
                        const int n4 = int(per >> 2);
                        for (int i = tid; i < n4; i += 4 * kThreads) {
                            float4 a[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (i + u * kThreads < n4) a[u] = __ldcg(v4 + i + u * kThreads);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (i + u * kThreads < n4) {
                                    a[u].x = __fdividef(a[u].x - lo, den);
                                    a[u].y = __fdividef(a[u].y - lo, den);
                                    a[u].z = __fdividef(a[u].z - lo, den);
                                    a[u].w = __fdividef(a[u].w - lo, den);
                                    if (p.do_log) {
                                        a[u].x = __logf(a[u].x + 1e-8f);
                                        a[u].y = __logf(a[u].y + 1e-8f);
                                        a[u].z = __logf(a[u].z + 1e-8f);
                                        a[u].w = __logf(a[u].w + 1e-8f);
                                    }
                                    v4[i + u * kThreads] = a[u];
                                }
                            }
                        }
                    } else {
                        for (size_t i = tid; i < per; i += kThreads) {
                            float a = __ldcg(base + i);
                            a = __fdividef(a - lo, den);
                            base[i] = p.do_log ? __logf(a + 1e-8f) : a;
                        }
                    }
                }
            }
        } else {
            __syncthreads();   // publishes the stage list built this iteration
            if (MODE == FM_ACTIVITY) {
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
                // per-lane bitmap of frequency-masked / filtered bins: bit k2 -> ka + 32*k2,
                // bit 8 + k2 -> kb + 32*k2, bit 16 -> bin 256
                uint32_t zbits = 0;
                {
                    const int n_z = p.fmask != nullptr ? p.n_fmask : 0;
                    const int32_t* fmk = p.fmask + size_t(b) * p.n_fmask * 2;
                    for (int i = 0; i < n_z; ++i) {
                        const int size = fmk[2 * i], off = fmk[2 * i + 1];
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) {
                            if (unsigned(ka + 32 * k2 - off) < unsigned(size)) zbits |= 1u << k2;
                            if (unsigned(kb + 32 * k2 - off) < unsigned(size)) zbits |= 1u << (8 + k2);
                        }
                        if (unsigned(256 - off) < unsigned(size)) zbits |= 1u << 16;
                    }
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
        }
    }
}

size_t fused_smem_bytes(const FusedParams& p, int mode) {
    return make_layout(p.np_shift, mode == FM_MEL, p.n_mel, p.mel_f_n, p.mel_nw).total;
}

int fused_max_segments() { return kMaxStages; }

cudaError_t launch_fused(const FusedParams& p, int mode, int num_sms, cudaStream_t stream) {
    const int FR = 16 >> p.np_shift;
    const int tpc = (p.T + FR - 1) / FR;
    const long long n_tiles = (long long)p.B * p.n_groups * tpc;
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    const size_t smem = fused_smem_bytes(p, mode);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
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
