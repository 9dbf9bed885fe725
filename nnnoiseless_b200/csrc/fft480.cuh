// fft480.cuh -- register-resident pieces of the warp-cooperative 480-point complex FFT (forward, e^{-i}).
//
// One warp transforms one stream: 480 = 32 x 15.
//   step 1  lane b holds z[32 a + b], a = 0..14, and runs a 15-point DFT on them (Good-Thomas 3 x 5: no twiddles);
//   step 2  Y[b][k1] *= W480^(b k1);
//   step 3  a transpose through per-warp shared memory gives lane k1 the 32 values Y[.][k1]; a 32-point FFT in
//           registers yields Z[k1 + 15 k2], k2 = 0..31.
// No block barrier anywhere (only __syncwarp around the transpose).  The functions below are plain inlineable code on
// register arrays with compile-time indices; they also compile as host code (tools/fft480_model.cpp checks the index
// maps and the numerics against a direct DFT on the CPU before any GPU time is spent).
//
// Replaces the block-cooperative Stockham FFT of round 1 (easyfft / rustfft call sites src/features.rs:264,290: an
// unnormalised forward transform and an unnormalised inverse; the reference's FFT crates are not under /root/reference,
// any correct f32 FFT is within the stated tolerance -- tests/test_oracle_golden.py pins the oracle's).
#pragma once

#ifdef __CUDACC__
#define FFT_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define FFT_HD inline
struct float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

namespace nnb {

FFT_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFT_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFT_HD float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 3-point DFT in place (forward): out[k] = sum_a in[a] W3^(a k)
FFT_HD void dft3(float2& a0, float2& a1, float2& a2) {
    const float s = 0.86602540378443864676f;
    const float2 t1 = c_add(a1, a2), d = c_sub(a1, a2);
    const float2 m1 = make_float2(a0.x - 0.5f * t1.x, a0.y - 0.5f * t1.y);
    a0 = c_add(a0, t1);
    a1 = make_float2(m1.x + s * d.y, m1.y - s * d.x);
    a2 = make_float2(m1.x - s * d.y, m1.y + s * d.x);
}

// 5-point DFT in place (forward)
FFT_HD void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 t1 = c_add(a1, a4), t2 = c_add(a2, a3);
    const float2 d1 = c_sub(a1, a4), d2 = c_sub(a2, a3);
    const float2 u1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
    const float2 u2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
    const float2 v1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
    const float2 v2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
    a0 = c_add(a0, c_add(t1, t2));
    a1 = make_float2(u1.x + v1.y, u1.y - v1.x);
    a4 = make_float2(u1.x - v1.y, u1.y + v1.x);
    a2 = make_float2(u2.x + v2.y, u2.y - v2.x);
    a3 = make_float2(u2.x - v2.y, u2.y + v2.x);
}

// 15-point DFT in place, natural order in and out: v[k] = sum_a v[a] W15^(a k).
// Good-Thomas: a = (5 a1 + 3 a2) mod 15, k = (10 k1 + 6 k2) mod 15  =>  W15^(a k) = W3^(a1 k1) W5^(a2 k2).
FFT_HD void dft15(float2 (&v)[15]) {
    float2 t[3][5];
#pragma unroll
    for (int a1 = 0; a1 < 3; a1++) {
#pragma unroll
        for (int a2 = 0; a2 < 5; a2++) t[a1][a2] = v[(5 * a1 + 3 * a2) % 15];
        dft5(t[a1][0], t[a1][1], t[a1][2], t[a1][3], t[a1][4]);  // over a2 -> index k2
    }
#pragma unroll
    for (int k2 = 0; k2 < 5; k2++) {
        dft3(t[0][k2], t[1][k2], t[2][k2]);  // over a1 -> index k1
#pragma unroll
        for (int k1 = 0; k1 < 3; k1++) v[(10 * k1 + 6 * k2) % 15] = t[k1][k2];
    }
}

// cos / sin of 2 pi j / 32, j = 0..8
#define FFT32_C0 1.0f
#define FFT32_C1 0.98078528040323044913f
#define FFT32_C2 0.92387953251128675613f
#define FFT32_C3 0.83146961230254523708f
#define FFT32_C4 0.70710678118654752440f
#define FFT32_C5 0.55557023301960222474f
#define FFT32_C6 0.38268343236508977173f
#define FFT32_C7 0.19509032201612826785f
#define FFT32_C8 0.0f

// W32^j = (cos(2 pi j / 32), -sin(2 pi j / 32)), j = 0..15, as compile-time constants
FFT_HD float2 w32(int j) {
    const float c[17] = {FFT32_C0, FFT32_C1, FFT32_C2, FFT32_C3, FFT32_C4, FFT32_C5, FFT32_C6, FFT32_C7, FFT32_C8,
                         -FFT32_C7, -FFT32_C6, -FFT32_C5, -FFT32_C4, -FFT32_C3, -FFT32_C2, -FFT32_C1, -1.0f};
    // sin(2 pi j / 32) = cos(2 pi (8 - j) / 32)
    const float sn = j <= 8 ? c[8 - j] : c[j - 8];
    return make_float2(c[j], -sn);
}

// (u - w) * W32^j with the trivial cases written out (j is a compile-time constant after unrolling)
FFT_HD float2 twiddle_mul32(float2 d, int j) {
    if (j == 0) return d;
    if (j == 8) return make_float2(d.y, -d.x);  // * -i
    if (j == 4) return make_float2(FFT32_C4 * (d.x + d.y), FFT32_C4 * (d.y - d.x));   // * (1 - i) / sqrt 2
    if (j == 12) return make_float2(FFT32_C4 * (d.y - d.x), -FFT32_C4 * (d.x + d.y));  // * (-1 - i) / sqrt 2
    return c_mul(d, w32(j));
}

FFT_HD constexpr int bitrev5(int k) { return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4); }

// 32-point FFT in place (forward), radix-2 decimation in frequency: natural order in, BIT-REVERSED order out:
// X[k] ends up in v[bitrev5(k)].
FFT_HD void fft32_dif(float2 (&v)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1, tstep = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const float2 u = v[blk + j], w = v[blk + j + half];
                v[blk + j] = c_add(u, w);
                v[blk + j + half] = twiddle_mul32(c_sub(u, w), j * tstep);
            }
        }
    }
}

}  // namespace nnb

namespace nnb {

// Even/odd split of the 480-point complex FFT Z of z[n] = x[2n] + i x[2n+1] into bins k and 480-k of the 960-point
// real FFT, scaled by wn (src/features.rs:290-295).  zc = Z[480-k] (Z[0] for k = 0), tw = exp(-2 pi i k / 960).
// For k = 0 the two results are bins 0 and 480 (imaginary parts exactly 0).
FFT_HD void rfft_split_pair(float2 zk, float2 zc, float2 tw, float wn, bool k_is_zero, float2& r0, float2& r1) {
    const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
    const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);
    const float tx = di * tw.x + dr * tw.y, ty = di * tw.y - dr * tw.x;
    r0 = make_float2((er + tx) * wn, (ei + ty) * wn);
    r1 = make_float2((er - tx) * wn, (ty - ei) * wn);
    if (k_is_zero) {
        r0.y = 0.0f;
        r1.y = 0.0f;
    }
}

// Inverse (unnormalised, src/features.rs:263-275): from bins xk = X[k], xc = X[480-k] build conj(Z[k]) and
// conj(Z[480-k]) of the 480-point complex sequence whose FORWARD FFT, conjugated, is the time signal pair-packed:
// y[2n] = Re o[n], y[2n+1] = -Im o[n].  Imaginary parts of DC / Nyquist are ignored (k = 0), as realfft does.
FFT_HD void irfft_pretwist_pair(float2 xk, float2 xc, float2 w, bool k_is_zero, float2& z0, float2& z1) {
    const float xi = k_is_zero ? 0.0f : xk.y, yi = k_is_zero ? 0.0f : xc.y;
    const float sr = xk.x + xc.x, si = xi - yi;
    const float dr = xk.x - xc.x, di = xi + yi;
    const float tx = dr * w.x + di * w.y, ty = di * w.x - dr * w.y;
    z0 = make_float2(sr - ty, -(si + tx));  // conj(Z[k])
    z1 = make_float2(sr + ty, si - tx);     // conj(Z[480-k])
}

}  // namespace nnb
