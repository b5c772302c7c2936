// In-register radix FFT building blocks for the 512-point frame transform.
//
// One 512-point complex FFT is done by a HALF-WARP (16 lanes x 32 points):
//   n = 16*n1 + n2   (lane = n2, register = n1)       k = k1 + 32*k2
//   pass 1 (in registers): 32-pt FFT over n1            -> A[n2][k1]
//   twiddle:               A[n2][k1] *= W512^(n2*k1)
//   exchange (shared memory): lane L gets k1 in {L, 32-L} (lane 0: {0,16}), all n2
//   pass 2 (in registers): two 16-pt FFTs over n2       -> Z[k1 + 32*k2]
// With that ownership Z[k] and Z[512-k] live in the same lane, so splitting the
// packed FFT of two real channels (z = ch0 + i*ch1) into the two one-sided
// spectra is lane-local.
//
// Everything here compiles for host too (g++ -x c++), which is how the math is
// unit-tested without a GPU (tests/test_fftcore_host.py).
#pragma once
#include <utility>

#if defined(__CUDACC__)
#define IRIS_HD __host__ __device__ __forceinline__
#else
#define IRIS_HD inline
#endif

namespace iris {

struct alignas(8) cpx {
    float x, y;
};

// ---- complex arithmetic on packed pairs ----
// On sm_100a a cpx lives in an aligned 64-bit register pair and every operation below is ONE
// packed FP32 instruction (PTX add/mul/fma.f32x2 -> SASS FADD2 / FMUL2 / FFMA2: two IEEE fp32
// results per issue slot).  The operand swaps, sign patterns and scalar broadcasts written here
// as re-packed pairs ({a.y, -a.x}, {s, s}, ...) cost nothing: ptxas folds them into the
// .LO_HI / .NP / .F32 operand modifiers of the packed instructions (checked with cuobjdump).
// The host build (unit tests of the index maps and the math) uses the scalar forms.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long cx_pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ cpx cx_up(unsigned long long r) {
    cpx a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
// (a.x + b.x, a.y + b.y)
__device__ __forceinline__ cpx cadd(cpx a, cpx b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(cx_pk(a.x, a.y)), "l"(cx_pk(b.x, b.y)));
    return cx_up(d);
}
// (a.x * b.x, a.y * b.y)
__device__ __forceinline__ cpx cmul2(cpx a, cpx b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(cx_pk(a.x, a.y)), "l"(cx_pk(b.x, b.y)));
    return cx_up(d);
}
// (a.x * b.x + c.x, a.y * b.y + c.y)
__device__ __forceinline__ cpx cfma2(cpx a, cpx b, cpx c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(cx_pk(a.x, a.y)), "l"(cx_pk(b.x, b.y)), "l"(cx_pk(c.x, c.y)));
    return cx_up(d);
}
#else
inline cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
inline cpx cmul2(cpx a, cpx b) { return cpx{a.x * b.x, a.y * b.y}; }
inline cpx cfma2(cpx a, cpx b, cpx c) { return cpx{a.x * b.x + c.x, a.y * b.y + c.y}; }
#endif
IRIS_HD cpx csub(cpx a, cpx b) { return cadd(a, cpx{-b.x, -b.y}); }
IRIS_HD cpx cscale(cpx a, float s) { return cmul2(a, cpx{s, s}); }
// a + s * b
IRIS_HD cpx caxpy(float s, cpx b, cpx a) { return cfma2(cpx{s, s}, b, a); }
// -i * a
IRIS_HD cpx cmuli_neg(cpx a) { return cpx{a.y, -a.x}; }
// a * w: two packed instructions.  The swap / sign pattern sits on the ACCUMULATOR of the
// second one so that a compile-time w becomes two 32-bit immediates (with the swap on a
// multiplicand ptxas has to build the (w.y, w.y) pair in registers first: 2 extra MOVs).
IRIS_HD cpx cmul(cpx a, cpx w) {
    const cpx t = cmul2(a, cpx{w.y, w.y});              // (a.x w.y, a.y w.y)
    return cfma2(a, cpx{w.x, w.x}, cpx{-t.y, t.x});      // (a.x w.x - a.y w.y, a.y w.x + a.x w.y)
}

// ---- compile-time trigonometry (double precision Taylor on a reduced octant) ----
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double cx_sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
constexpr double cx_cos_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2*pi*num/den), exact octant reduction on the rational num/den
constexpr double cx_cos2pi(long long num, long long den) {
    num %= den;
    if (num < 0) num += den;
    if (2 * num > den) num = den - num;                           // cos(2pi(1-a)) = cos(2pi a)
    if (4 * num > den) return -cx_cos2pi(den - 2 * num, 2 * den);  // cos(2pi a) = -cos(2pi(1/2-a))
    if (8 * num > den) return cx_sin_small(2.0 * kPi * double(den - 4 * num) / double(4 * den));
    return cx_cos_small(2.0 * kPi * double(num) / double(den));
}
constexpr double cx_sin2pi(long long num, long long den) {
    return cx_cos2pi(4 * num - den, 4 * den);  // sin(x) = cos(x - pi/2)
}

// a * exp(-2*pi*i*NUM/DEN), trivial rotations folded at compile time
template <int NUM, int DEN>
IRIS_HD cpx mul_w(cpx a) {
    constexpr int n = ((NUM % DEN) + DEN) % DEN;
    constexpr float h = 0.70710678118654752440f;
    if constexpr (n == 0) {
        return a;
    } else if constexpr (4 * n == DEN) {
        return cpx{a.y, -a.x};
    } else if constexpr (2 * n == DEN) {
        return cpx{-a.x, -a.y};
    } else if constexpr (4 * n == 3 * DEN) {
        return cpx{-a.y, a.x};
    } else if constexpr (8 * n == DEN) {
        return cscale(cadd(a, cpx{a.y, -a.x}), h);        // ((x + y) h, (y - x) h)
    } else if constexpr (8 * n == 3 * DEN) {
        return cscale(cadd(a, cpx{-a.y, a.x}), -h);       // ((y - x) h, -(x + y) h)
    } else if constexpr (8 * n == 5 * DEN) {
        return cscale(cadd(a, cpx{a.y, -a.x}), -h);       // (-(x + y) h, (x - y) h)
    } else if constexpr (8 * n == 7 * DEN) {
        return cscale(cadd(a, cpx{-a.y, a.x}), h);        // ((x - y) h, (x + y) h)
    } else {
        constexpr float c = float(cx_cos2pi(n, DEN));
        constexpr float s = float(cx_sin2pi(n, DEN));
        return cmul(a, cpx{c, -s});                       // (x c + y s, y c - x s)
    }
}

template <int N>
struct Fft;

template <>
struct Fft<2> {
    static IRIS_HD void run(cpx (&v)[2]) {
        cpx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <>
struct Fft<4> {
    static IRIS_HD void run(cpx (&v)[4]) {
        const cpx t0 = cadd(v[0], v[2]);
        const cpx t1 = csub(v[0], v[2]);
        const cpx t2 = cadd(v[1], v[3]);
        const cpx t3 = cmuli_neg(csub(v[1], v[3]));  // -i * (v1 - v3)
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <int N, int Q, int M, int... Rs>
IRIS_HD void twiddle_row(cpx (&row)[M], std::integer_sequence<int, Rs...>) {
    ((row[Rs] = mul_w<Q * Rs, N>(row[Rs])), ...);
}
template <int N, int R, int M, int... Qs>
IRIS_HD void twiddle_all(cpx (&sub)[R][M], std::integer_sequence<int, Qs...>) {
    (twiddle_row<N, Qs, M>(sub[Qs], std::make_integer_sequence<int, M>{}), ...);
}

// Cooley-Tukey N = R*M, natural order in and out; everything unrolls to registers.
template <int N, int R>
IRIS_HD void fft_ct(cpx (&v)[N]) {
    constexpr int M = N / R;
    cpx sub[R][M];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
        for (int p = 0; p < M; ++p) sub[q][p] = v[R * p + q];
#pragma unroll
    for (int q = 0; q < R; ++q) Fft<M>::run(sub[q]);
    twiddle_all<N, R, M>(sub, std::make_integer_sequence<int, R>{});
#pragma unroll
    for (int r = 0; r < M; ++r) {
        cpx col[R];
#pragma unroll
        for (int q = 0; q < R; ++q) col[q] = sub[q][r];
        Fft<R>::run(col);
#pragma unroll
        for (int s = 0; s < R; ++s) v[r + M * s] = col[s];
    }
}

template <>
struct Fft<8> {
    static IRIS_HD void run(cpx (&v)[8]) { fft_ct<8, 2>(v); }
};
template <>
struct Fft<16> {
    static IRIS_HD void run(cpx (&v)[16]) { fft_ct<16, 4>(v); }
};
template <>
struct Fft<32> {
    static IRIS_HD void run(cpx (&v)[32]) { fft_ct<32, 4>(v); }
};

// ---- exchange-buffer geometry (per half-warp slot, 4 rounds of 8 k1 values) ----
// Round rho holds k1 in [8*rho, 8*rho+8) as 4 rows a (k1 pair 8*rho+2a, +1) of
// 16 lanes x float4 {A[k1].x, A[k1].y, A[k1+1].x, A[k1+1].y}; rows are padded by
// 16 B so that the per-lane column reads are bank-conflict free.
constexpr int kXchRowFloats = 16 * 4 + 4;                 // 272 B
constexpr int kXchSlotFloats = 4 * kXchRowFloats;         // 1088 B
IRIS_HD int xch_write_off(int a, int n2) { return a * kXchRowFloats + n2 * 4; }             // float4
IRIS_HD int xch_read_off(int k1_in_round, int n2) {                                         // float2
    return (k1_in_round >> 1) * kXchRowFloats + n2 * 4 + (k1_in_round & 1) * 2;
}

// k1 values owned by lane L after the exchange
IRIS_HD int own_k1a(int L) { return L; }                      // lane 0: 0
IRIS_HD int own_k1b(int L) { return L == 0 ? 16 : 32 - L; }

}  // namespace iris
