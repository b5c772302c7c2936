// Host-side check of challenge_b200/csrc/fftwarp.cuh: emulates the 32 lanes of one warp
// (window incl. the w[n+256] = 1 - w[n] form, pass 1, twiddle, exchange through a byte
// buffer with the kernel's offsets, DIF + pass 2, partner/mirror maps, split of the two packed
// real channels) and compares with a float64 DFT.  Prints max errors as JSON.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fftwarp.cuh"
using namespace iris;

int main() {
    const int N = 512;
    std::vector<double> x0(N), x1(N);
    srand(11);
    for (int i = 0; i < N; ++i) {
        x0[i] = rand() / double(RAND_MAX) - 0.5;
        x1[i] = rand() / double(RAND_MAX) - 0.5;
    }
    // tables exactly as iris_abi.cu builds them
    std::vector<float> tw1(8 * 32 * 4), ts(8 * 2 * 4), hann(N);
    for (int q = 0; q < 8; ++q)
        for (int n2 = 0; n2 < 32; ++n2)
            for (int h = 0; h < 2; ++h) {
                const double a = -2.0 * M_PI * double((2 * q + h) * n2) / 512.0;
                tw1[4 * (q * 32 + n2) + 2 * h] = float(cos(a));
                tw1[4 * (q * 32 + n2) + 2 * h + 1] = float(sin(a));
            }
    for (int m = 0; m < 8; ++m)
        for (int p = 0; p < 2; ++p)
            for (int h = 0; h < 2; ++h) {
                const double a = -2.0 * M_PI * double(2 * m + h) / 32.0;
                ts[4 * (m * 2 + p) + 2 * h] = p ? float(cos(a)) : 1.f;
                ts[4 * (m * 2 + p) + 2 * h + 1] = p ? float(sin(a)) : 0.f;
            }
    for (int n = 0; n < N; ++n) hann[n] = float(0.5 - 0.5 * cos(2.0 * M_PI * n / N));

    std::vector<unsigned char> xch(kXwBytes, 0);
    int bad_map = 0;
    // pass 1 per lane; gain 0.5 folded in like the kernel does (k_tiles halves the gains)
    for (int lane = 0; lane < 32; ++lane) {
        cpx v[16];
        float w8[8];
        for (int i = 0; i < 8; ++i) w8[i] = hann[lane + 32 * i];
        for (int i = 0; i < 16; ++i) {
            const int n = lane + 32 * i;
            v[i] = cpx{0.5f * float(x0[n]), 0.5f * float(x1[n])};
        }
        for (int i = 0; i < 8; ++i) {
            v[i].x *= w8[i];
            v[i].y *= w8[i];
            v[i + 8].x = v[i + 8].x - w8[i] * v[i + 8].x;   // w[n + 256] = 1 - w[n]
            v[i + 8].y = v[i + 8].y - w8[i] * v[i + 8].y;
        }
        auto twv = [&](int k) { const float* t = &tw1[4 * ((k >> 1) * 32 + lane) + 2 * (k & 1)]; return cpx{t[0], t[1]}; };
        warp_pass1(v, twv(1), twv(2), twv(4), twv(8));
        for (int k1 = 0; k1 < 16; ++k1) memcpy(&xch[xw_write_off(k1, lane)], &v[k1], 8);
    }
    // pass 2 per lane
    static cpx Y[32][16];
    for (int lane = 0; lane < 32; ++lane) {
        const int k1 = warp_k1(lane), p = warp_par(lane);
        const float s = p ? -1.f : 1.f;
        cpx u[16];
        for (int m = 0; m < 8; ++m) {
            float a[4], b[4];
            memcpy(a, &xch[xw_read_off(k1, m)], 16);
            memcpy(b, &xch[xw_read_off(k1, m + 8)], 16);
            const float* t = &ts[4 * (m * 2 + p)];
            u[2 * m] = warp_dif(cpx{a[0], a[1]}, cpx{b[0], b[1]}, s, p != 0, t[0], t[1]);
            u[2 * m + 1] = warp_dif(cpx{a[2], a[3]}, cpx{b[2], b[3]}, s, p != 0, t[2], t[3]);
        }
        Fft<16>::run(u);
        for (int j = 0; j < 16; ++j) Y[lane][j] = u[j];
    }
    // reference: windowed one-sided spectra of the two real channels (float64)
    double err = 0, max_abs = 0, im_dc = 0, im_nyq = 0;
    std::vector<int> seen(257, 0);
    for (int lane = 0; lane < 32; ++lane) {
        const int P = warp_partner(lane);
        for (int j = 0; j <= 8; ++j) {
            if (j == 8 && lane != 0) continue;
            const int k = warp_bin(lane, j);
            if (k > 256) { ++bad_map; continue; }
            const int jm = (j == 8) ? 8 : warp_mirror_reg(lane, j);
            const cpx zf = Y[lane][j], zm = Y[P][jm];
            // the mirror really is bin 512 - k
            if (((512 - k) & 511) != warp_bin(P, jm)) ++bad_map;
            const float r0 = zf.x + zm.x, i0 = zf.y - zm.y, r1 = zf.y + zm.y, i1 = zm.x - zf.x;
            double e0r = 0, e0i = 0, e1r = 0, e1i = 0;
            for (int n = 0; n < N; ++n) {
                const double w = 0.5 - 0.5 * cos(2 * M_PI * n / N), a = -2 * M_PI * double(n) * k / N;
                e0r += w * x0[n] * cos(a); e0i += w * x0[n] * sin(a);
                e1r += w * x1[n] * cos(a); e1i += w * x1[n] * sin(a);
            }
            err = fmax(err, fmax(fmax(fabs(e0r - r0), fabs(e0i - i0)), fmax(fabs(e1r - r1), fabs(e1i - i1))));
            max_abs = fmax(max_abs, fmax(fabs(e0r), fabs(e1r)));
            if (k == 0) im_dc = fmax(fabs(i0), fabs(i1));
            if (k == 256) im_nyq = fmax(fabs(i0), fabs(i1));
            seen[k]++;
        }
    }
    int missing = 0;
    for (int k = 0; k <= 256; ++k) missing += (seen[k] != 1);
    // bins < 128 need only registers j < 4 and mirrors j >= 12 (lane 0: also register 0)
    int bad_prune = 0;
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 16; ++j) {
            const int k = warp_bin(lane, j);
            if (k < 128 && j >= 4) ++bad_prune;
            if (j < 4) {
                const int jm = warp_mirror_reg(lane, j);
                if (!(jm >= 12 || jm == 0)) ++bad_prune;
            }
        }
    printf("{\"err512\": %.3e, \"max_abs\": %.3e, \"im_dc\": %.3e, \"im_nyq\": %.3e, \"bad_map\": %d, "
           "\"missing\": %d, \"bad_prune\": %d}\n", err, max_abs, im_dc, im_nyq, bad_map, missing, bad_prune);
    return 0;
}
