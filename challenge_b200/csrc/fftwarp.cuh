// Warp-level 512-point complex FFT: 32 lanes x 16 points, ONE shared-memory exchange.
//
//   n = n2 + 32*i        (lane = n2, register = i)          k = k1 + 16*k2
//   pass 1 (registers):  A[n2][k1] = sum_i x[n2 + 32 i] W16^(i k1)          (16-point FFT)
//   twiddle:             B[n2][k1] = A[n2][k1] * W512^(n2 k1)
//   exchange (shared):   row k1 holds B[.][k1] for all 32 n2
//   pass 2 (registers):  lane L = k1 + 16*p computes the outputs k2 = 2j + p of the 32-point
//                        FFT of its row by one radix-2 DIF step + a 16-point FFT:
//                          u[i] = (B[i] + s*B[i+16]) * (p ? W32^i : 1),  s = p ? -1 : +1
//                          Y[j] = sum_i u[i] W16^(i j) = X[k1 + 16*(2j + p)]
// Mirror bins: X[512-k] of lane (k1, p), register j sits in lane (16-k1, 1-p), register 15-j
// (k1 = 0: same lane; p = 0: register (16-j) & 15), so the split of a packed two-channel
// transform needs one register-indexed-at-compile-time shuffle per bin.
// For the log-mel features only bins < 128 are needed: j < 4 and the mirrors j >= 12, i.e.
// half of the outputs of pass 2 (the dead butterflies are eliminated at compile time).
//
// Everything here compiles for the host too; tests/host/fftwarp_host.cpp emulates the 32
// lanes and checks the index maps and the math against a float64 DFT.
#pragma once
#include "fftcore.cuh"

namespace iris {

// exchange buffer: 16 rows (k1) of 32 complex, row stride padded by 16 B so that the 16
// rows read by a warp's 128-bit loads fall into distinct 16-byte bank groups (2 wavefronts
// for 256 distinct bytes = conflict-free)
constexpr int kXwRowBytes = 32 * 8 + 16;          // 272
constexpr int kXwBytes = 16 * kXwRowBytes;        // 4352 per warp

IRIS_HD int xw_write_off(int k1, int n2) { return k1 * kXwRowBytes + n2 * 8; }        // float2
IRIS_HD int xw_read_off(int k1, int m) { return k1 * kXwRowBytes + m * 16; }          // float4 {B[2m], B[2m+1]}

// lane L = k1 + 16 * par: the 16 lanes of a half-warp hold 16 consecutive bins, so that the
// shared-memory rows a quarter-warp touches are distinct (conflict-free 128-bit accesses)
IRIS_HD int warp_k1(int lane) { return lane & 15; }
IRIS_HD int warp_par(int lane) { return lane >> 4; }
// bin held by register j of a lane after pass 2
IRIS_HD int warp_bin(int lane, int j) { return (lane & 15) + 16 * (2 * j + (lane >> 4)); }
// lane holding the mirror bins of this lane
IRIS_HD int warp_partner(int lane) {
    return (lane & 15) == 0 ? lane : (16 - (lane & 15)) + 16 * (1 - (lane >> 4));
}
// register of the partner that holds the mirror of own register j
IRIS_HD int warp_mirror_reg(int lane, int j) { return lane == 0 ? ((16 - j) & 15) : 15 - j; }

// pass 1 on the 16 registers of a lane, then the inter-pass twiddle.
// tw[q] = {W512^(n2*2q), W512^(n2*(2q+1))} as (cos, sin) pairs for this lane.
template <class TwLoad>
IRIS_HD void warp_pass1(cpx (&v)[16], TwLoad tw) {
    Fft<16>::run(v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float wx0, wy0, wx1, wy1;
        tw(q, wx0, wy0, wx1, wy1);
        if (q > 0) v[2 * q] = cmul(v[2 * q], cpx{wx0, wy0});
        v[2 * q + 1] = cmul(v[2 * q + 1], cpx{wx1, wy1});
    }
}

// one DIF element of pass 2: u = (a + s*b) * (odd ? t : 1).  The odd half-warp multiplies by
// the compile-time twiddle t under a predicate (no selects); the even half keeps the sum.
IRIS_HD cpx warp_dif(cpx a, cpx b, float s, bool odd, float tx, float ty) {
    cpx d = caxpy(s, b, a);
    if (odd) d = cmul(d, cpx{tx, ty});
    return d;
}

}  // namespace iris
