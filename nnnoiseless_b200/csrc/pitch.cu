// pitch.cu -- ORDER-EXACT pitch analysis (src/pitch.rs:45-489), 32 streams per thread block.
//
// Compiled with -fmad=false and written with explicit round-to-nearest intrinsics: every f32 operation
// is rounded like the reference's scalar code and every sum runs in the reference's order, so the pitch
// period (an integer) is bit-identical to the reference restatement for every frame.
//
// Why 32 streams per block: the pitch path alternates strictly sequential recurrences (5-lag
// autocorrelation sums, Levinson, running energies with a clamp per step, best/second-best selection,
// the k = 2..15 sub-harmonic ladder) with small dense sums (147-lag cross-correlation, ~40 inner
// products of 480).  With one stream per block the recurrences run on one lane of a warp (measured:
// 18.6 of 32 lanes active, 25.7k warp-instructions per stream).  Here every recurrence runs
// LANE-PER-STREAM (32 streams advance in lock-step in one warp, operands fetched as float4 rows of the
// shared-memory tile), while the dense sums are spread over (stream, lag-group) lane-tasks packed
// densely into warps and use register sliding windows (one LDS.128 per 16 multiply-adds).
//
// Shared-memory tile (dynamic, ~208 KB, one block per SM):
//   P   [32][868]  2x-decimated, LPC-whitened history (pitch_buf); row stride 868 = 16B aligned and
//                  = 4 (mod 32) so that lane-per-stream float4 reads are bank-conflict free
//   Y4  [32][436]  its even samples (the 4x-decimated signal); later reused for the fine running
//                  energies yn2 [32][297] and then for yy_lookup [32][387]
//   XC  [32][149]  coarse cross-correlation      YN4 [32][149]  coarse running energy
#include "common.cuh"

namespace nnb {

namespace {

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }

constexpr int SB = 32;    // streams per block
constexpr int NT = 512;   // threads per block
constexpr int NW = NT / 32;
constexpr int PB = PITCH_BUF_SIZE / 2;                                 // 864
constexpr int MAXP = PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD;          // 588
constexpr int N4 = PITCH_FRAME_SIZE / 4;                               // 240
constexpr int NL4 = MAXP / 4;                                          // 147 coarse lags
constexpr int NL2 = MAXP / 2;                                          // 294 fine lags
constexpr int HALF_MAX = PITCH_MAX_PERIOD / 2;                         // 384
constexpr int HALF_N = PITCH_FRAME_SIZE / 2;                           // 480
constexpr int MIN_PERIOD2 = PITCH_MIN_PERIOD / 2;                      // 30

constexpr int P_LD = 868;
constexpr int Y4_LD = 436;
constexpr int XC_LD = 149;
constexpr int YN2_LD = 297;
constexpr int YY_LD = 387;
constexpr int IPR_LD = 31;
constexpr int FX_LD = 11;
constexpr int NGRP = (NL4 + 3) / 4;  // 37 lag groups of 4

constexpr int OFF_P = 0;
constexpr int OFF_Y4 = OFF_P + SB * P_LD;
constexpr int OFF_XC = OFF_Y4 + SB * Y4_LD;
constexpr int OFF_YN4 = OFF_XC + SB * XC_LD;
constexpr int OFF_AC = OFF_YN4 + SB * XC_LD;     // [5][32]
constexpr int OFF_LPC = OFF_AC + 5 * SB;         // [5][32]
constexpr int OFF_XX = OFF_LPC + 5 * SB;         // [32]
constexpr int OFF_PG = OFF_XX + SB;              // [32]
constexpr int OFF_IPR = OFF_PG + SB;             // [32][31]
constexpr int OFF_FX = OFF_IPR + SB * IPR_LD;    // [32][11]
constexpr int OFF_SI = OFF_FX + SB * FX_LD;      // int [5][32]: best4, second4, pitch_idx/t0, t, task counters
constexpr int SMEM_FLOATS = OFF_SI + 5 * SB;
static_assert(SB * YN2_LD <= SB * Y4_LD && SB * YY_LD <= SB * Y4_LD, "yn2 / yy must fit in the Y4 region");
static_assert(SMEM_FLOATS * 4 <= 227 * 1024, "shared-memory tile too large");

__constant__ int c_second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};  // src/pitch.rs:489

__device__ __forceinline__ float pitch_gain(float xy, float xx, float yy) {
    return __fdiv_rn(xy, __fsqrt_rn(fa(1.0f, fm(xx, yy))));  // src/pitch.rs:485-487
}

// Selection step of find_best_pitch (src/pitch.rs:383-400).
struct BestTwo {
    float best_num = -1.0f, second_num = -1.0f, best_den = 0.0f, second_den = 0.0f;
    int best = 0, second = 1;
    __device__ __forceinline__ void consider(int i, float corr, float ysq) {
        if (corr > 0.0f) {
            float num = fm(corr, corr);
            if (fm(num, second_den) > fm(second_num, ysq)) {
                if (fm(num, best_den) > fm(best_num, ysq)) {
                    second_num = best_num;
                    second_den = best_den;
                    second = best;
                    best_num = num;
                    best_den = ysq;
                    best = i;
                } else {
                    second_num = num;
                    second_den = ysq;
                    second = i;
                }
            }
        }
    }
};

// celt_autocorr lag K for one stream (lane-per-stream): sum_{j<860} p[j] p[j+K] in order, then the tail
// sum_{i=K+860}^{863} p[i] p[i-K] (src/pitch.rs:433-446, 296-363).  row = 217 float4.
template <int K>
__device__ __forceinline__ float autocorr_lag(const float4* __restrict__ row) {
    float c = 0.0f;
    float4 w0 = row[0];
#pragma unroll 5
    for (int m = 0; m < (PB - 4) / 4; m++) {
        const float4 w1 = row[m + 1];
        const float e[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int d = 0; d < 4; d++) c = fa(c, fm(e[d], e[d + K]));
        w0 = w1;
    }
    // w0 = p[860..863]
    const float e[4] = {w0.x, w0.y, w0.z, w0.w};
    float d = 0.0f;
#pragma unroll
    for (int i = K; i < 4; i++) d = fa(d, fm(e[i], e[i - K]));
    return fa(c, d);
}

// inner_prod(x, y, 480) of src/pitch.rs:225-244 for one (stream, lag) lane-task:
// xr = aligned float4 row of x (pbuf + 384), y = pbuf + 384 - lag (unaligned scalars).
__device__ __forceinline__ float inner_prod_480(const float4* __restrict__ xr, const float* __restrict__ y) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
    for (int m = 0; m < HALF_N / 4; m++) {
        const float4 x = xr[m];
        s0 = fa(s0, fm(x.x, y[4 * m]));
        s1 = fa(s1, fm(x.y, y[4 * m + 1]));
        s2 = fa(s2, fm(x.z, y[4 * m + 2]));
        s3 = fa(s3, fm(x.w, y[4 * m + 3]));
    }
    return fa(fa(fa(s0, s1), s2), s3);
}

__global__ void __launch_bounds__(NT, 1) pitch32_kernel(const float* __restrict__ hist, int32_t* __restrict__ last_period,
                                                        float* __restrict__ last_gain, int32_t* __restrict__ pitch_out,
                                                        int n_streams, int hbase) {
    extern __shared__ __align__(16) float sm[];
    float* P = sm + OFF_P;
    float* Y4 = sm + OFF_Y4;
    float* XC = sm + OFF_XC;
    float* YN4 = sm + OFF_YN4;
    float* AC = sm + OFF_AC;
    float* LPC = sm + OFF_LPC;
    float* XX = sm + OFF_XX;
    float* PG = sm + OFF_PG;
    float* IPR = sm + OFF_IPR;
    float* FX = sm + OFF_FX;
    int* SI = reinterpret_cast<int*>(sm + OFF_SI);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = blockIdx.x * SB;
    const int ns = min(SB, n_streams - s0);

    // ---- Ph1: pitch_downsample part 1 (src/pitch.rs:455-458); rows of absent streams are zero ----
    for (int r = warp; r < SB; r += NW) {
        float* prow = P + r * P_LD;
        if (r < ns) {
            const float* h = hist + (size_t)(s0 + r) * HIST_CAP;
            for (int i = lane; i < PB; i += 32) {
                int p1 = hbase + 2 * i;
                int pa = p1 - 1, pb = p1 + 1;
                if (p1 >= HIST_CAP) p1 -= HIST_CAP;
                if (pa >= HIST_CAP) pa -= HIST_CAP;
                if (pb >= HIST_CAP) pb -= HIST_CAP;
                float v;
                if (i == 0) v = fm(fa(fm(h[pb], 0.5f), h[p1]), 0.5f);
                else v = fm(fa(fm(fa(h[pa], h[pb]), 0.5f), h[p1]), 0.5f);
                prow[i] = v;
            }
        } else {
            for (int i = lane; i < PB; i += 32) prow[i] = 0.0f;
        }
        if (lane < 4) {
            prow[PB + lane] = 0.0f;
            Y4[r * Y4_LD + PB / 2 + lane] = 0.0f;
        }
    }
    if (tid == 0) SI[4 * SB] = 0;  // xcorr task counter
    if (tid == 1) SI[4 * SB + 1] = 0;
    __syncthreads();

    // ---- Ph2: celt_autocorr, warp k = lag k, lane = stream ----
    if (warp < 5) {
        const float4* row = reinterpret_cast<const float4*>(P + lane * P_LD);
        float v;
        switch (warp) {
            case 0: v = autocorr_lag<0>(row); break;
            case 1: v = autocorr_lag<1>(row); break;
            case 2: v = autocorr_lag<2>(row); break;
            case 3: v = autocorr_lag<3>(row); break;
            default: v = autocorr_lag<4>(row); break;
        }
        AC[warp * SB + lane] = v;
    }
    __syncthreads();

    // ---- Ph3: noise floor, lag window, LPC(4), bandwidth expansion, extra zero (src/pitch.rs:462-480, 257-292) ----
    if (warp == 0) {
        float a[5];
#pragma unroll
        for (int i = 0; i < 5; i++) a[i] = AC[i * SB + lane];
        a[0] = fm(a[0], 1.0001f);
#pragma unroll
        for (int i = 1; i < 5; i++) {
            float w = fm(0.008f, (float)i);
            a[i] = fs(a[i], fm(fm(a[i], w), w));
        }
        float lpc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (a[0] != 0.0f) {
            float error = a[0];
            bool live = true;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (live) {
                    float rr = 0.0f;
#pragma unroll
                    for (int j = 0; j < i; j++) rr = fa(rr, fm(lpc[j], a[i - j]));
                    rr = fa(rr, a[i + 1]);
                    float r = __fdiv_rn(-rr, error);
                    lpc[i] = r;
#pragma unroll
                    for (int j = 0; j < (i + 1) / 2; j++) {
                        float t1 = lpc[j], t2 = lpc[i - 1 - j];
                        lpc[j] = fa(t1, fm(r, t2));
                        lpc[i - 1 - j] = fa(t2, fm(r, t1));
                    }
                    error = fs(error, fm(fm(r, r), error));
                    if (error < fm(0.001f, a[0])) live = false;  // "bail out once we get 30 dB gain"
                }
            }
        }
        float tmp = 1.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            tmp = fm(tmp, 0.9f);
            lpc[i] = fm(lpc[i], tmp);
        }
        LPC[0 * SB + lane] = fa(lpc[0], 0.8f);
        LPC[1 * SB + lane] = fa(lpc[1], fm(0.8f, lpc[0]));
        LPC[2 * SB + lane] = fa(lpc[2], fm(0.8f, lpc[1]));
        LPC[3 * SB + lane] = fa(lpc[3], fm(0.8f, lpc[2]));
        LPC[4 * SB + lane] = fm(0.8f, lpc[3]);
    }
    __syncthreads();

    // ---- Ph4: fir5_in_place (src/pitch.rs:407-429) + second decimation (src/pitch.rs:74-79).
    // One warp per row, 32-sample chunks from the END of the row backwards, so the 5 older inputs a chunk
    // needs are still un-filtered when it is processed. ----
    for (int r = warp; r < SB; r += NW) {
        float* prow = P + r * P_LD;
        const float n0 = LPC[0 * SB + r], n1 = LPC[1 * SB + r], n2 = LPC[2 * SB + r], n3 = LPC[3 * SB + r], n4 = LPC[4 * SB + r];
        for (int c = PB / 32 - 1; c >= 0; c--) {
            const int i = 32 * c + lane;
            const float x = prow[i];
            const float m0 = i >= 1 ? prow[i - 1] : 0.0f;
            const float m1 = i >= 2 ? prow[i - 2] : 0.0f;
            const float m2 = i >= 3 ? prow[i - 3] : 0.0f;
            const float m3 = i >= 4 ? prow[i - 4] : 0.0f;
            const float m4 = i >= 5 ? prow[i - 5] : 0.0f;
            const float o = fa(fa(fa(fa(fa(x, fm(n0, m0)), fm(n1, m1)), fm(n2, m2)), fm(n3, m3)), fm(n4, m4));
            __syncwarp();
            prow[i] = o;
            if ((lane & 1) == 0) Y4[r * Y4_LD + (i >> 1)] = o;
        }
    }
    __syncthreads();

    // ---- Ph5: coarse xcorr (all warps, dynamic lane-task groups) + coarse running energy (warp 14) + xx (warp 15) ----
    if (warp == NW - 2) {
        // y_sq_norm of find_best_pitch(xcorr, y_lp4, 240) (src/pitch.rs:379-382, 401-402); YN4[i] = value seen at lag i
        const float4* row = reinterpret_cast<const float4*>(Y4 + lane * Y4_LD);
        float y = 1.0f;
#pragma unroll 4
        for (int m = 0; m < N4 / 4; m++) {
            const float4 v = row[m];
            y = fa(y, fm(v.x, v.x));
            y = fa(y, fm(v.y, v.y));
            y = fa(y, fm(v.z, v.z));
            y = fa(y, fm(v.w, v.w));
        }
        float* out = YN4 + lane * XC_LD;
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < NGRP; m++) {
            const float4 va = row[N4 / 4 + m], vb = row[m];
            const float a[4] = {va.x, va.y, va.z, va.w}, b[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int d = 0; d < 4; d++) {
                y = fmaxf(fa(y, fs(fm(a[d], a[d]), fm(b[d], b[d]))), 1.0f);
                if (4 * m + d + 1 < XC_LD) out[4 * m + d + 1] = y;
            }
        }
    } else if (warp == NW - 1) {
        // xx = inner_prod(x, x, 480) with its four interleaved accumulators (src/pitch.rs:133, 225-244)
        const float4* xr = reinterpret_cast<const float4*>(P + lane * P_LD + HALF_MAX);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
        for (int m = 0; m < HALF_N / 4; m++) {
            const float4 x = xr[m];
            a0 = fa(a0, fm(x.x, x.x));
            a1 = fa(a1, fm(x.y, x.y));
            a2 = fa(a2, fm(x.z, x.z));
            a3 = fa(a3, fm(x.w, x.w));
        }
        XX[lane] = fa(fa(fa(a0, a1), a2), a3);
    }
    // coarse xcorr (src/pitch.rs:82, 296-363): lane-task = (stream, group of 4 consecutive lags); every
    // accumulator sums x_lp4[j] * y_lp4[lag + j] with j ascending, operands via a sliding register window.
    for (;;) {
        int T = 0;
        if (lane == 0) T = atomicAdd(&SI[4 * SB], 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= NGRP) break;
        const int L = T * 32 + lane;  // 37 * 32 lane-tasks = 32 streams x 37 groups
        const int s = L / NGRP, g = L - s * NGRP;
        const float4* xr = reinterpret_cast<const float4*>(Y4 + s * Y4_LD + HALF_MAX / 2);
        const float4* yr = reinterpret_cast<const float4*>(Y4 + s * Y4_LD + 4 * g);
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
        float4 w = yr[0];
#pragma unroll 2
        for (int m = 0; m < N4 / 4; m++) {
            const float4 x = xr[m];
            const float4 wn = yr[m + 1];
            const float e[8] = {w.x, w.y, w.z, w.w, wn.x, wn.y, wn.z, wn.w};
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                c0 = fa(c0, fm(xv[u], e[u]));
                c1 = fa(c1, fm(xv[u], e[u + 1]));
                c2 = fa(c2, fm(xv[u], e[u + 2]));
                c3 = fa(c3, fm(xv[u], e[u + 3]));
            }
            w = wn;
        }
        float* o = XC + s * XC_LD + 4 * g;
        o[0] = c0;
        o[1] = c1;
        o[2] = c2;
        if (4 * g + 3 < NL4) o[3] = c3;
    }
    __syncthreads();

    // ---- Ph6: coarse best/second (warp 0, serial over lags, lane = stream) + fine running energy (warp 1) ----
    float* YN2 = Y4;  // the 4x-decimated copy is dead from here on
    if (warp == 0) {
        BestTwo b2;
        const float* xc = XC + lane * XC_LD;
        const float* yn = YN4 + lane * XC_LD;
#pragma unroll 3
        for (int i = 0; i < NL4; i++) b2.consider(i, xc[i], yn[i]);
        SI[0 * SB + lane] = b2.best;
        SI[1 * SB + lane] = b2.second;
    } else if (warp == 1) {
        // y_sq_norm of find_best_pitch(xcorr, y, 480): YN2[i] = value seen at fine lag i
        const float4* row = reinterpret_cast<const float4*>(P + lane * P_LD);
        float y = 1.0f;
#pragma unroll 4
        for (int m = 0; m < HALF_N / 4; m++) {
            const float4 v = row[m];
            y = fa(y, fm(v.x, v.x));
            y = fa(y, fm(v.y, v.y));
            y = fa(y, fm(v.z, v.z));
            y = fa(y, fm(v.w, v.w));
        }
        float* out = YN2 + lane * YN2_LD;
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < (NL2 + 3) / 4; m++) {
            const float4 va = row[HALF_N / 4 + m], vb = row[m];
            const float a[4] = {va.x, va.y, va.z, va.w}, b[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int d = 0; d < 4; d++) {
                y = fmaxf(fa(y, fs(fm(a[d], a[d]), fm(b[d], b[d]))), 1.0f);
                if (4 * m + d + 1 < YN2_LD) out[4 * m + d + 1] = y;
            }
        }
    }
    __syncthreads();

    // ---- Ph7: fine search, 10 candidate lags per stream (src/pitch.rs:88-96); lane-task = (stream, candidate) ----
    for (int L = tid; L < SB * 10; L += NT) {
        const int s = L / 10, c = L - s * 10;
        const int best4 = SI[0 * SB + s], second4 = SI[1 * SB + s];
        const int i = (c < 5) ? (2 * best4 - 2 + c) : (2 * second4 - 2 + (c - 5));
        float v = 0.0f;
        if (i >= 0 && i < NL2) {
            const float* prow = P + s * P_LD;
            v = fmaxf(inner_prod_480(reinterpret_cast<const float4*>(prow + HALF_MAX), prow + i), -1.0f);
        }
        FX[s * FX_LD + c] = v;
    }
    __syncthreads();

    // ---- Ph8: fine best + pseudo-interpolation (src/pitch.rs:97-114), lane = stream ----
    if (warp == 0) {
        const int best4 = SI[0 * SB + lane], second4 = SI[1 * SB + lane];
        const float* fx = FX + lane * FX_LD;
        const float* yn = YN2 + lane * YN2_LD;
        const int loA = 2 * best4 - 2, loB = 2 * second4 - 2;
        // xcorr value at fine lag i: inside either 5-wide window it is the computed value, elsewhere 0
        auto xcf = [&](int i) -> float {
            if (i < 0 || i >= NL2) return 0.0f;
            if (i >= loA && i <= loA + 4) return fx[i - loA];
            if (i >= loB && i <= loB + 4) return fx[5 + i - loB];
            return 0.0f;
        };
        BestTwo b2;
        // lags outside the windows have xcorr 0 and can never be selected: scan the windows in ascending order
        const int c0 = min(loA, loB), c1 = max(loA, loB);
        const int lo0 = max(c0, 0), hi0 = min(c0 + 4, NL2 - 1);
        const int lo1 = max(max(c1, 0), hi0 + 1), hi1 = min(c1 + 4, NL2 - 1);
        for (int i = lo0; i <= hi0; i++) b2.consider(i, xcf(i), yn[i]);
        for (int i = lo1; i <= hi1; i++) b2.consider(i, xcf(i), yn[i]);
        const int best = b2.best;
        int offset = 0;
        if (best > 0 && best < NL2 - 1) {
            const float a = xcf(best - 1), b = xcf(best), c = xcf(best + 1);
            if (fs(c, a) > fm(0.7f, fs(b, a))) offset = 1;
            else if (fs(a, c) > fm(0.7f, fs(b, c))) offset = -1;
        }
        const int pitch_idx = PITCH_MAX_PERIOD - (2 * best - offset);  // src/pitch.rs:49,114
        SI[2 * SB + lane] = min(pitch_idx / 2, HALF_MAX - 1);          // t0 of remove_doubling
    }
    __syncthreads();

    // ---- Ph9: remove_doubling inner products (src/pitch.rs:134,167-168) + yy_lookup chain (warp 15 first) ----
    float* YY = Y4;  // yn2 is dead from here on
    if (warp == NW - 1) {
        // yy_lookup (src/pitch.rs:135-142): stored clamped at 0, carried unclamped; i = 1..384 descending rows
        const float4* row = reinterpret_cast<const float4*>(P + lane * P_LD);
        float* out = YY + lane * YY_LD;
        float y = XX[lane];
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < HALF_MAX / 4; m++) {
            const float4 va = row[HALF_MAX / 4 - 1 - m];              // p[380-4m .. 383-4m]
            const float4 vb = row[(HALF_MAX + HALF_N) / 4 - 1 - m];   // p[860-4m .. 863-4m]
            const float a[4] = {va.w, va.z, va.y, va.x}, b[4] = {vb.w, vb.z, vb.y, vb.x};
#pragma unroll
            for (int d = 0; d < 4; d++) {
                y = fa(y, fs(fm(a[d], a[d]), fm(b[d], b[d])));
                out[4 * m + d + 1] = fmaxf(y, 0.0f);
            }
        }
    }
    // lane-task = (stream, q): q = 1: xy(t0); q = 2 + 2(k-2) + {0,1}: t1(k), t1b(k), k = 2..15
    for (;;) {
        int T = 0;
        if (lane == 0) T = atomicAdd(&SI[4 * SB + 1], 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= 29) break;  // 29 * 32 lane-tasks
        const int L = T * 32 + lane;
        const int s = L / 29, q = 1 + (L - s * 29);
        const int t0 = SI[2 * SB + s];
        int lagq = -1;
        if (q == 1) lagq = t0;
        else {
            const int k = 2 + ((q - 2) >> 1);
            const int t1 = (2 * t0 + k) / (2 * k);
            if (t1 >= MIN_PERIOD2) {
                if (((q - 2) & 1) == 0) lagq = t1;
                else if (k == 2) lagq = (t1 + t0 > HALF_MAX) ? t0 : t0 + t1;
                else lagq = (2 * c_second_check[k] * t0 + k) / (2 * k);
            }
        }
        float v = 0.0f;
        if (lagq >= 0) {
            const float* prow = P + s * P_LD;
            v = inner_prod_480(reinterpret_cast<const float4*>(prow + HALF_MAX), prow + HALF_MAX - lagq);
        }
        IPR[s * IPR_LD + q] = v;
    }
    __syncthreads();

    // ---- Ph10: the sub-harmonic ladder (src/pitch.rs:144-203), lane = stream ----
    if (warp == 0) {
        const float* ipr = IPR + lane * IPR_LD;
        const float* yy = YY + lane * YY_LD;
        const int t0 = SI[2 * SB + lane];
        const float xx = XX[lane];
        float xy = ipr[1];
        float yyv = yy[t0];
        int prev_period = 0;
        float lg = 0.0f;
        if (lane < ns) {
            prev_period = last_period[s0 + lane] / 2;
            lg = last_gain[s0 + lane];
        }
        float best_xy = xy, best_yy = yyv;
        const float g0 = pitch_gain(xy, xx, yyv);
        float g = g0;
        int t = t0;
        for (int k = 2; k <= 15; k++) {
            const int t1 = (2 * t0 + k) / (2 * k);
            if (t1 < MIN_PERIOD2) break;
            int t1b;
            if (k == 2) t1b = (t1 + t0 > HALF_MAX) ? t0 : t0 + t1;
            else t1b = (2 * c_second_check[k] * t0 + k) / (2 * k);
            xy = fm(fa(ipr[2 + 2 * (k - 2)], ipr[3 + 2 * (k - 2)]), 0.5f);
            yyv = fm(fa(yy[t1], yy[t1b]), 0.5f);
            const float g1 = pitch_gain(xy, xx, yyv);
            const int d = abs(t1 - prev_period);
            float cont;
            if (d <= 1) cont = lg;
            else if (d <= 2 && 5 * k * k < t0) cont = fm(lg, 0.5f);
            else cont = 0.0f;
            float thresh;
            if (t1 < 3 * MIN_PERIOD2) thresh = fmaxf(fs(fm(0.85f, g0), cont), 0.4f);
            else if (t1 < 2 * MIN_PERIOD2) thresh = fmaxf(fs(fm(0.9f, g0), cont), 0.5f);  // dead branch, as in the reference
            else thresh = fmaxf(fs(fm(0.7f, g0), cont), 0.3f);
            if (g1 > thresh) {
                best_xy = xy;
                best_yy = yyv;
                t = t1;
                g = g1;
            }
        }
        best_xy = fmaxf(best_xy, 0.0f);
        float pg = (best_yy <= best_xy) ? 1.0f : __fdiv_rn(best_xy, fa(best_yy, 1.0f));
        pg = fminf(pg, g);
        SI[3 * SB + lane] = t;
        PG[lane] = pg;
    }
    __syncthreads();

    // ---- Ph11: +-1 refinement (src/pitch.rs:205-218): three inner products per stream ----
    for (int L = tid; L < SB * 3; L += NT) {
        const int s = L / 3, c = L - s * 3;
        const int t = SI[3 * SB + s];
        const float* prow = P + s * P_LD;
        IPR[s * IPR_LD + c] = inner_prod_480(reinterpret_cast<const float4*>(prow + HALF_MAX), prow + HALF_MAX - (t + c - 1));
    }
    __syncthreads();
    if (warp == 0 && lane < ns) {
        const float* ipr = IPR + lane * IPR_LD;
        const float x_0 = ipr[0], x_1 = ipr[1], x_2 = ipr[2];
        const int t = SI[3 * SB + lane];
        int offset = 0;
        if (fs(x_2, x_0) > fm(0.7f, fs(x_1, x_0))) offset = 1;
        else if (fs(x_0, x_2) > fm(0.7f, fs(x_1, x_2))) offset = -1;
        const int tf = max(2 * t + offset, PITCH_MIN_PERIOD);
        pitch_out[s0 + lane] = tf;
        last_period[s0 + lane] = tf;
        last_gain[s0 + lane] = PG[lane];
    }
}

}  // namespace

cudaError_t launch_pitch(const BatchBuffers& b, int slot, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(float) * SMEM_FLOATS;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pitch32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int grid = (b.n_streams + SB - 1) / SB;
    pitch32_kernel<<<grid, NT, smem, st>>>(b.hist, b.last_period, b.last_gain, b.pitch, b.n_streams, hist_base(slot));
    return cudaGetLastError();
}

}  // namespace nnb
