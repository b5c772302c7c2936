// Host-side check of challenge_b200/csrc/fftcore.cuh: simulates the 16 lanes of one
// half-warp slot (pass 1, twiddle, 4-round exchange, pass 2, lane-local split of the
// two packed real channels) and compares with a float64 DFT.  Prints max errors.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fftcore.cuh"
using namespace iris;

int main() {
    const int N = 512;
    std::vector<double> x0(N), x1(N);
    srand(7);
    for (int i = 0; i < N; ++i) {
        x0[i] = rand() / double(RAND_MAX) - 0.5;
        x1[i] = rand() / double(RAND_MAX) - 0.5;
    }
    // small FFT self-checks
    double err_small = 0;
    {
        cpx v[32];
        for (int i = 0; i < 32; ++i) v[i] = cpx{float(x0[i]), float(x1[i])};
        Fft<32>::run(v);
        for (int k = 0; k < 32; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 32; ++n) {
                double a = -2 * M_PI * n * k / 32.0;
                re += x0[n] * cos(a) - x1[n] * sin(a);
                im += x0[n] * sin(a) + x1[n] * cos(a);
            }
            err_small = fmax(err_small, fmax(fabs(re - v[k].x), fabs(im - v[k].y)));
        }
        cpx u[16];
        for (int i = 0; i < 16; ++i) u[i] = cpx{float(x0[i]), float(x1[i])};
        Fft<16>::run(u);
        for (int k = 0; k < 16; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 16; ++n) {
                double a = -2 * M_PI * n * k / 16.0;
                re += x0[n] * cos(a) - x1[n] * sin(a);
                im += x0[n] * sin(a) + x1[n] * cos(a);
            }
            err_small = fmax(err_small, fmax(fabs(re - u[k].x), fabs(im - u[k].y)));
        }
    }
    // tables
    std::vector<float> whalf(N);
    for (int n = 0; n < N; ++n) whalf[n] = float(0.5 * (0.5 - 0.5 * cos(2 * M_PI * n / N)));
    std::vector<cpx> tw(32 * 16);
    for (int k1 = 0; k1 < 32; ++k1)
        for (int n2 = 0; n2 < 16; ++n2) {
            double a = -2 * M_PI * (k1 * n2) / 512.0;
            tw[k1 * 16 + n2] = cpx{float(cos(a)), float(sin(a))};
        }
    // pass 1 per lane
    static cpx A[16][32];
    for (int n2 = 0; n2 < 16; ++n2) {
        cpx v[32];
        for (int n1 = 0; n1 < 32; ++n1) {
            int n = 16 * n1 + n2;
            v[n1] = cpx{float(x0[n]) * whalf[n], float(x1[n]) * whalf[n]};
        }
        Fft<32>::run(v);
        for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmul(v[k1], tw[k1 * 16 + n2]);
        for (int k1 = 0; k1 < 32; ++k1) A[n2][k1] = v[k1];
    }
    // exchange in 4 rounds through a slot buffer
    static cpx Za[16][16], Zb[16][16];
    std::vector<float> slot(kXchSlotFloats);
    for (int rho = 0; rho < 4; ++rho) {
        for (int n2 = 0; n2 < 16; ++n2)
            for (int a = 0; a < 4; ++a) {
                float* p = &slot[xch_write_off(a, n2)];
                p[0] = A[n2][8 * rho + 2 * a].x;
                p[1] = A[n2][8 * rho + 2 * a].y;
                p[2] = A[n2][8 * rho + 2 * a + 1].x;
                p[3] = A[n2][8 * rho + 2 * a + 1].y;
            }
        for (int L = 0; L < 16; ++L) {
            int ka = own_k1a(L), kb = own_k1b(L);
            if ((ka >> 3) == rho)
                for (int n2 = 0; n2 < 16; ++n2) {
                    const float* p = &slot[xch_read_off(ka & 7, n2)];
                    Za[L][n2] = cpx{p[0], p[1]};
                }
            if ((kb >> 3) == rho)
                for (int n2 = 0; n2 < 16; ++n2) {
                    const float* p = &slot[xch_read_off(kb & 7, n2)];
                    Zb[L][n2] = cpx{p[0], p[1]};
                }
        }
    }
    // pass 2 + split
    std::vector<double> R0(257), I0(257), R1(257), I1(257);
    std::vector<int> seen(257, 0);
    for (int L = 0; L < 16; ++L) {
        Fft<16>::run(Za[L]);
        Fft<16>::run(Zb[L]);
        int ka = own_k1a(L), kb = own_k1b(L);
        cpx PA[9], PB[8];
        for (int k2 = 0; k2 < 9; ++k2) PA[k2] = (L == 0) ? Za[L][(16 - k2) & 15] : Zb[L][15 - (k2 & 7) - (k2 >> 3) * 0];
        for (int k2 = 0; k2 < 8; ++k2) PB[k2] = (L == 0) ? Zb[L][15 - k2] : Za[L][15 - k2];
        auto emit = [&](int f, cpx zf, cpx zm) {
            R0[f] = zf.x + zm.x; I0[f] = zf.y - zm.y;
            R1[f] = zf.y + zm.y; I1[f] = zm.x - zf.x;
            seen[f]++;
        };
        for (int k2 = 0; k2 < 8; ++k2) {
            emit(ka + 32 * k2, Za[L][k2], PA[k2]);
            emit(kb + 32 * k2, Zb[L][k2], PB[k2]);
        }
        if (L == 0) emit(256, Za[L][8], PA[8]);
    }
    double err = 0, mx = 0;
    int bad_seen = 0;
    for (int f = 0; f <= 256; ++f) {
        if (seen[f] != 1) bad_seen++;
        double r0 = 0, i0 = 0, r1 = 0, i1 = 0;
        for (int n = 0; n < N; ++n) {
            double w = 0.5 - 0.5 * cos(2 * M_PI * n / N);
            double a = -2 * M_PI * n * f / double(N);
            r0 += w * x0[n] * cos(a); i0 += w * x0[n] * sin(a);
            r1 += w * x1[n] * cos(a); i1 += w * x1[n] * sin(a);
        }
        mx = fmax(mx, fmax(fabs(r0), fabs(i0)));
        err = fmax(err, fmax(fmax(fabs(r0 - R0[f]), fabs(i0 - I0[f])),
                             fmax(fabs(r1 - R1[f]), fabs(i1 - I1[f]))));
    }
    printf("{\"err_small\": %.3e, \"err512\": %.3e, \"max_abs\": %.3e, \"bad_seen\": %d, "
           "\"im_dc\": %.1f, \"im_nyq\": %.1f}\n",
           err_small, err, mx, bad_seen, I0[0] + I1[0], I0[256] + I1[256]);
    return 0;
}
