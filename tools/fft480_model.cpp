// fft480_model.cpp -- CPU check of nnnoiseless_b200/csrc/fft480.cuh: the 32-lane / 15-register decomposition of the
// 480-point FFT (index maps, twiddles, bit reversal), the real-FFT split and the inverse pre-twist, against direct DFT
// sums in double precision.  TEST INFRASTRUCTURE: emulates the warp lane by lane.
//   g++ -O2 -o /tmp/fft480_model tools/fft480_model.cpp && /tmp/fft480_model
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../nnnoiseless_b200/csrc/fft480.cuh"
using namespace nnb;

static const double PI = 3.14159265358979323846264338327950288;

// the warp algorithm, lane by lane; z[480] -> Z[480]
static void fft480_warp(const float2* z, float2* Z) {
    static float2 S[15][32];
    for (int b = 0; b < 32; b++) {  // steps 1 + 2
        float2 v[15];
        for (int a = 0; a < 15; a++) v[a] = z[32 * a + b];
        dft15(v);
        for (int k1 = 0; k1 < 15; k1++) {
            const double ang = -2.0 * PI * (double)(b * k1) / 480.0;
            const float2 tw = make_float2((float)cos(ang), (float)sin(ang));
            S[k1][b] = k1 == 0 ? v[k1] : c_mul(v[k1], tw);
        }
    }
    for (int k1 = 0; k1 < 15; k1++) {  // step 3
        float2 r[32];
        for (int b = 0; b < 32; b++) r[b] = S[k1][b];
        fft32_dif(r);
        for (int k2 = 0; k2 < 32; k2++) Z[k1 + 15 * k2] = r[bitrev5(k2)];
    }
}

int main() {
    srand(1);
    double worst = 0;
    // complex FFT
    {
        std::vector<float2> z(480), Z(480);
        for (auto& v : z) v = make_float2((float)(rand() % 20001 - 10000), (float)(rand() % 20001 - 10000));
        fft480_warp(z.data(), Z.data());
        double num = 0, den = 0;
        for (int k = 0; k < 480; k++) {
            double sr = 0, si = 0;
            for (int n = 0; n < 480; n++) {
                const double ang = -2.0 * PI * (double)((long)n * k % 480) / 480.0;
                sr += z[n].x * cos(ang) - z[n].y * sin(ang);
                si += z[n].x * sin(ang) + z[n].y * cos(ang);
            }
            num += (Z[k].x - sr) * (Z[k].x - sr) + (Z[k].y - si) * (Z[k].y - si);
            den += sr * sr + si * si;
        }
        printf("fft480 rel rms error %.3g\n", sqrt(num / den));
        worst = fmax(worst, sqrt(num / den));
    }
    // real forward (960 reals -> 481 bins) and inverse round trip
    {
        std::vector<float> x(960);
        for (auto& v : x) v = (float)(rand() % 20001 - 10000);
        std::vector<float2> z(480), Z(480), X(481);
        for (int n = 0; n < 480; n++) z[n] = make_float2(x[2 * n], x[2 * n + 1]);
        fft480_warp(z.data(), Z.data());
        for (int k = 0; k <= 240; k++) {
            const double ang = -2.0 * PI * k / 960.0;
            float2 r0, r1;
            rfft_split_pair(Z[k], Z[k == 0 ? 0 : 480 - k], make_float2((float)cos(ang), (float)sin(ang)), 1.0f, k == 0, r0, r1);
            X[k] = r0;
            if (k != 240) X[480 - k] = r1;
        }
        double num = 0, den = 0;
        for (int k = 0; k <= 480; k++) {
            double sr = 0, si = 0;
            for (int n = 0; n < 960; n++) {
                const double ang = -2.0 * PI * (double)((long)n * k % 960) / 960.0;
                sr += x[n] * cos(ang);
                si += x[n] * sin(ang);
            }
            num += (X[k].x - sr) * (X[k].x - sr) + (X[k].y - si) * (X[k].y - si);
            den += sr * sr + si * si;
        }
        printf("rfft960 rel rms error %.3g\n", sqrt(num / den));
        worst = fmax(worst, sqrt(num / den));
        // inverse
        std::vector<float2> zi(480), o(480);
        for (int k = 0; k <= 240; k++) {
            const double ang = -2.0 * PI * k / 960.0;
            float2 z0, z1;
            irfft_pretwist_pair(X[k], X[480 - k], make_float2((float)cos(ang), (float)sin(ang)), k == 0, z0, z1);
            zi[k] = z0;
            if (k != 0 && k != 240) zi[480 - k] = z1;
        }
        fft480_warp(zi.data(), o.data());
        num = den = 0;
        for (int n = 0; n < 480; n++) {
            const double y0 = o[n].x / 960.0, y1 = -o[n].y / 960.0;  // unnormalised inverse: 960 x
            num += (y0 - x[2 * n]) * (y0 - x[2 * n]) + (y1 - x[2 * n + 1]) * (y1 - x[2 * n + 1]);
            den += x[2 * n] * x[2 * n] + x[2 * n + 1] * x[2 * n + 1];
        }
        printf("irfft960(rfft960(x)) / 960 rel rms error %.3g\n", sqrt(num / den));
        worst = fmax(worst, sqrt(num / den));
    }
    if (worst > 1e-6) {
        printf("FAIL\n");
        return 1;
    }
    printf("ok\n");
    return 0;
}
