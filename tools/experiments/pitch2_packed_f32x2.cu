// EXPERIMENT, NOT BUILT INTO THE LIBRARY (round 1): the pitch kernel on packed f32x2 arithmetic, two streams per lane.
// Result on B200: bit-exact (all tests/test_gpu_parity.py cases pass with it) but SLOWER than the scalar kernel,
// 1.21 ms vs 0.92 ms per 65,536-stream frame (gpurun 2026-09-23).  Why: FADD2 / FFMA2 issue at half rate (each
// occupies the FP32 pipe for two cycles whatever the number of active lanes), so the non-FMA multiply-add costs the
// same pipe time packed or scalar; the dense phases were already FP-pipe bound while they run, the serial phases are
// latency bound (no gain), and lane-per-PAIR warps have only 8 of 32 lanes active yet still pay two pipe cycles per
// instruction.  Kept for the two findings a later round can build on:
//   * ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 despite the explicit .rn (the scalar forms are
//     left alone); the multiply has to be written fma(a, b, -0.0) with the -0.0 taken from constant memory (see mul2);
//   * measured latencies / rates of the packed instructions: tools/ubench/f32x2.cu.
// To try it: copy to nnnoiseless_b200/csrc/pitch2.cu, add ("pitch2.cu", ["-fmad=false"]) to build.py's UNITS and call
// launch_pitch2 from host.cu's launch_stage.
//
// pitch2.cu -- ORDER-EXACT pitch analysis (src/pitch.rs:45-489) on PACKED f32x2 arithmetic: two streams per lane.
//
// Same algorithm, phases and summation order as pitch.cu (which stays as the scalar reference implementation,
// NNB_PITCH_V1=1); what changes is the data layout and the instruction set.  Blackwell executes add/mul/sub on
// PAIRS of f32 (`add.rn.f32x2` -> FADD2, `mul.rn.f32x2` -> FMUL2), each half rounded to nearest exactly like the
// scalar instruction (no FMA contraction: -fmad=false and explicit .rn).  Measured on B200 (tools/ubench/f32x2.cu):
// dependent FADD2 latency 4.5 cycles (FADD 4.7), non-FMA throughput 90 lane-MAC/cycle/SM packed vs 52 scalar, and a
// single warp sustains 21 vs 11 lane-MAC/cycle.  So every shared-memory array holds PAIRS (stream 2p, stream 2p+1)
// as one 64-bit element, every recurrence runs lane-per-PAIR, and every dense sum is a (pair, lag group) lane-task:
// half the FP and half the address instructions per stream, identical per-stream results.
//
// Where the two streams of a pair need different addresses (the fine windows around each stream's own coarse
// candidates, the sub-harmonic lags T/k, the final +-1 refinement) the x operand is still shared (one 128-bit load =
// two pairs) and the y operand is gathered per half with 32-bit loads into the two halves of a pair register;
// data-dependent selections (find_best_pitch, the threshold ladder) run lane-per-stream on the pair layout.
//
// Shared-memory tile per block of SB = 16 streams = SP = 8 pairs (64-bit elements, ~106 KB -> two blocks per SM):
//   P2  [SP][870]  2x-decimated, LPC-whitened history; row stride 870 (x 8 B) keeps lane-per-pair 128-bit reads
//                  conflict-free (870 / 2 odd)
//   Y42 [SP][438]  its even samples; later reused for the fine running energies yn2 [SP][297], then yy [SP][387]
//   XC2 / YN42 [SP][149]  coarse cross-correlation / running energy
#include "common.cuh"

namespace nnb {

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ u64 pk(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float hi(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// There is no packed multiply in the ISA: ptxas lowers mul.rn.f32x2 to FFMA2(a, b, 0) and then (CUDA 12.9) merges it
// with a following add.rn.f32x2 into ONE fused FFMA2 -- a contraction the explicit .rn forbids and the scalar path
// never does.  Writing the product as fma(a, b, -0.0) with the -0.0 pair read from constant memory (a value ptxas
// cannot see) keeps it a separate, correctly rounded multiply: a*b + (-0.0) == RN(a*b) for every a*b, signed zeros
// included, and FFMA2 followed by FADD2 cannot be merged.
__constant__ u64 c_negzero2 = 0x8000000080000000ull;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c_negzero2));
    return r;
}
__device__ __forceinline__ u64 splat(float a) { return pk(a, a); }
__device__ __forceinline__ u64 max2(u64 v, float m) { return pk(fmaxf(lo(v), m), fmaxf(hi(v), m)); }

#ifndef PITCH_SB
#define PITCH_SB 16
#endif
#ifndef PITCH_NT
#define PITCH_NT 256
#endif
constexpr int SB = PITCH_SB;  // streams per block
constexpr int SP = SB / 2;    // pairs per block (lane-per-pair phases use lanes 0..SP-1, mirrored on the others)
constexpr int NT = PITCH_NT;
constexpr int NW = NT / 32;
static_assert(SB % 2 == 0 && SB <= 32 && (32 % SB) == 0 && NW >= 5, "phase-to-warp assignment assumes >= 5 warps");
constexpr int PB = PITCH_BUF_SIZE / 2;                         // 864
constexpr int MAXP = PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD;  // 588
constexpr int N4 = PITCH_FRAME_SIZE / 4;                       // 240
constexpr int NL4 = MAXP / 4;                                  // 147 coarse lags
constexpr int NL2 = MAXP / 2;                                  // 294 fine lags
constexpr int HALF_MAX = PITCH_MAX_PERIOD / 2;                 // 384
constexpr int HALF_N = PITCH_FRAME_SIZE / 2;                   // 480
constexpr int MIN_PERIOD2 = PITCH_MIN_PERIOD / 2;              // 30

// row strides in 64-bit elements
constexpr int P_LD = 870;
constexpr int Y4_LD = 438;
constexpr int XC_LD = 149;
constexpr int YN2_LD = 297;
constexpr int YY_LD = 387;
constexpr int IPR_LD = 31;
constexpr int FX_LD = 11;
constexpr int NGRP = (NL4 + 3) / 4;  // 37 lag groups of 4
constexpr int NQMAX = 29;            // inner products per stream in remove_doubling: 1 + 2 * 14

// offsets in 64-bit elements
constexpr int OFF_P = 0;
constexpr int OFF_Y4 = OFF_P + SP * P_LD;
constexpr int OFF_XC = OFF_Y4 + SP * Y4_LD;
constexpr int OFF_YN4 = OFF_XC + SP * XC_LD;
constexpr int OFF_AC = OFF_YN4 + SP * XC_LD;   // [5][SP]
constexpr int OFF_LPC = OFF_AC + 5 * SP;       // [5][SP]
constexpr int OFF_XX = OFF_LPC + 5 * SP;       // [SP]
constexpr int OFF_IPR = OFF_XX + SP;           // [SP][31]
constexpr int OFF_FX = OFF_IPR + SP * IPR_LD;  // [SP][11]
constexpr int OFF_INT = OFF_FX + SP * FX_LD;   // ints from here
// int region: SI [4][SB] (best4, second4, t0, spare), CTR [4], TASK [SP * 29], LAG [SB][29]
constexpr int INT_SI = 0;
constexpr int INT_CTR = INT_SI + 4 * SB;
constexpr int INT_TASK = INT_CTR + 4;
constexpr int INT_LAG = INT_TASK + SP * NQMAX;
constexpr int INT_TOTAL = INT_LAG + SB * NQMAX;
constexpr size_t SMEM_BYTES = (size_t)OFF_INT * 8 + (size_t)INT_TOTAL * 4;
static_assert(YN2_LD <= Y4_LD && YY_LD <= Y4_LD, "yn2 / yy must fit in the Y4 region");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "tile must fit in one SM");
constexpr int BLOCKS_PER_SM = (int)((227 * 1024) / (SMEM_BYTES + 1024));

__constant__ int c_second_check2[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};  // src/pitch.rs:489

__device__ __forceinline__ float pitch_gain(float xy, float xx, float yy) {
    return __fdiv_rn(xy, __fsqrt_rn(fa(1.0f, fm(xx, yy))));  // src/pitch.rs:485-487
}

// Selection step of find_best_pitch (src/pitch.rs:383-400) with selects instead of branches.
struct BestTwo {
    float best_num = -1.0f, second_num = -1.0f, best_den = 0.0f, second_den = 0.0f;
    int best = 0, second = 1;
    __device__ __forceinline__ void consider(int i, float corr, float ysq) {
        const float num = fm(corr, corr);
        const bool c2 = (corr > 0.0f) && (fm(num, second_den) > fm(second_num, ysq));
        const bool c1 = c2 && (fm(num, best_den) > fm(best_num, ysq));
        second_num = c1 ? best_num : (c2 ? num : second_num);
        second_den = c1 ? best_den : (c2 ? ysq : second_den);
        second = c1 ? best : (c2 ? i : second);
        best_num = c1 ? num : best_num;
        best_den = c1 ? ysq : best_den;
        best = c1 ? i : best;
    }
};

// 128-bit shared-memory access = two consecutive pairs
__device__ __forceinline__ void ld2(const u64* p, u64& a, u64& b) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void st2(u64* p, u64 a, u64 b) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b); }

// celt_autocorr lag K for one pair: sum_{j<860} p[j] p[j+K] in order, then the tail sum_{i=K+860}^{863} p[i] p[i-K]
// (src/pitch.rs:433-446, 296-363).
template <int K>
__device__ __forceinline__ u64 autocorr_lag(const u64* __restrict__ row) {
    u64 c = splat(0.0f);
    u64 e[8];
    ld2(row, e[0], e[1]);
    ld2(row + 2, e[2], e[3]);
#pragma unroll 5
    for (int m = 0; m < (PB - 4) / 4; m++) {
        ld2(row + 4 * m + 4, e[4], e[5]);
        ld2(row + 4 * m + 6, e[6], e[7]);
#pragma unroll
        for (int d = 0; d < 4; d++) c = add2(c, mul2(e[d], e[d + K]));
#pragma unroll
        for (int d = 0; d < 4; d++) e[d] = e[4 + d];
    }
    u64 t = splat(0.0f);  // e[0..3] = p[860..863]
#pragma unroll
    for (int i = K; i < 4; i++) t = add2(t, mul2(e[i], e[i - K]));
    return add2(c, t);
}

// y operand of a pair whose halves sit at different lags: element j of half A is ya[2 j], of half B yb[2 j]
// (ya / yb already point at the right half of the right pair element).
__device__ __forceinline__ u64 gather(const float* __restrict__ ya, const float* __restrict__ yb, int j) { return pk(ya[2 * j], yb[2 * j]); }

// inner_prod(x, y, 480) of src/pitch.rs:225-244 (four interleaved accumulators) for both halves of a pair.
__device__ __forceinline__ u64 inner_prod_480(const u64* __restrict__ xr, const float* __restrict__ ya, const float* __restrict__ yb) {
    u64 s0 = splat(0.0f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll 4
    for (int m = 0; m < HALF_N / 4; m++) {
        u64 x0, x1, x2, x3;
        ld2(xr + 4 * m, x0, x1);
        ld2(xr + 4 * m + 2, x2, x3);
        s0 = add2(s0, mul2(x0, gather(ya, yb, 4 * m)));
        s1 = add2(s1, mul2(x1, gather(ya, yb, 4 * m + 1)));
        s2 = add2(s2, mul2(x2, gather(ya, yb, 4 * m + 2)));
        s3 = add2(s3, mul2(x3, gather(ya, yb, 4 * m + 3)));
    }
    return add2(add2(add2(s0, s1), s2), s3);
}

// NLAG consecutive lags of inner_prod(x, y + lag, 480) for both halves of a pair with ONE sliding register window
// over y: acc[c][u] is the reference's accumulator u of lag c, y read once.
template <int NLAG>
__device__ __forceinline__ void inner_prod_window(const u64* __restrict__ xr, const float* __restrict__ ya, const float* __restrict__ yb,
                                                  u64* out) {
    u64 acc[NLAG][4];
#pragma unroll
    for (int c = 0; c < NLAG; c++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[c][u] = splat(0.0f);
    u64 w[8];
#pragma unroll
    for (int u = 0; u < 4; u++) w[u] = gather(ya, yb, u);
#pragma unroll 2
    for (int m = 0; m < HALF_N / 4; m++) {
        u64 xv[4];
        ld2(xr + 4 * m, xv[0], xv[1]);
        ld2(xr + 4 * m + 2, xv[2], xv[3]);
#pragma unroll
        for (int u = 0; u < 4; u++) w[4 + u] = gather(ya, yb, 4 * m + 4 + u);
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NLAG; c++) acc[c][u] = add2(acc[c][u], mul2(xv[u], w[u + c]));
#pragma unroll
        for (int u = 0; u < 4; u++) w[u] = w[4 + u];
    }
#pragma unroll
    for (int c = 0; c < NLAG; c++) out[c] = add2(add2(add2(acc[c][0], acc[c][1]), acc[c][2]), acc[c][3]);
}

#ifdef PITCH_PROFILE
__device__ unsigned long long g_pitch2_prof[16];
#define PPROF(k)                                                                   \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            const long long now_ = clock64();                                      \
            atomicAdd(&g_pitch2_prof[k], (unsigned long long)(now_ - pprof_t_));   \
            pprof_t_ = now_;                                                       \
        }                                                                          \
    } while (0)
#else
#define PPROF(k)
#endif

__global__ void __launch_bounds__(NT, (BLOCKS_PER_SM * NT <= 512) ? BLOCKS_PER_SM : 512 / NT)
    pitch2_kernel(const float* __restrict__ hist, int32_t* __restrict__ last_period, float* __restrict__ last_gain,
                  int32_t* __restrict__ pitch_out, int n_streams, int hbase) {
    extern __shared__ __align__(16) u64 sm2[];
    u64* P = sm2 + OFF_P;
    u64* Y4 = sm2 + OFF_Y4;
    u64* XC = sm2 + OFF_XC;
    u64* YN4 = sm2 + OFF_YN4;
    u64* AC = sm2 + OFF_AC;
    u64* LPC = sm2 + OFF_LPC;
    u64* XX = sm2 + OFF_XX;
    u64* IPR = sm2 + OFF_IPR;
    u64* FX = sm2 + OFF_FX;
    int* SI = reinterpret_cast<int*>(sm2 + OFF_INT) + INT_SI;
    int* CTR = reinterpret_cast<int*>(sm2 + OFF_INT) + INT_CTR;  // [0] xcorr task counter, [1] rd task counter, [2] number of rd tasks
    int* TASK = reinterpret_cast<int*>(sm2 + OFF_INT) + INT_TASK;
    int* LAG = reinterpret_cast<int*>(sm2 + OFF_INT) + INT_LAG;
    // scalar (one stream) view of a pair array: element i of stream s = base_f[(pair * LD + i) * 2 + (s & 1)]
    const float* Pf = reinterpret_cast<const float*>(P);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lp = lane % SP;  // lane-per-pair phases (lanes >= SP mirror lanes < SP: same reads, same writes)
    const int ls = lane % SB;  // lane-per-stream phases
    const int s0 = blockIdx.x * SB;
    const int ns = min(SB, n_streams - s0);
#ifdef PITCH_PROFILE
    long long pprof_t_ = clock64();
#endif

    // ---- Ph1: pitch_downsample part 1 (src/pitch.rs:455-458), one warp per pair row, both streams' 128-bit loads
    // issued before use.  x_lp[m] = .5 (.5 (x[2m-1] + x[2m+1]) + x[2m]); for m = 0 the reference drops x[-1]:
    // adding -0.0f instead is the identity on every float, signed zeros included. ----
    for (int r = warp; r < SP; r += NW) {
        u64* prow = P + r * P_LD;
        const bool onA = 2 * r < ns, onB = 2 * r + 1 < ns;
        const float* hA = hist + (size_t)(s0 + 2 * r) * HIST_CAP;
        const float* hB = hA + HIST_CAP;
        const u64 half = splat(0.5f);
        constexpr int NQ = 7;  // float4 per lane per half row (2 x 7 x 32 >= 432)
#pragma unroll 1
        for (int hh = 0; hh < 2; hh++) {
            float4 va[NQ], vb[NQ];
            float qa[NQ], qb[NQ];
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const int m = lane + 32 * (NQ * hh + k);
                va[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                vb[k] = va[k];
                qa[k] = -0.0f;
                qb[k] = -0.0f;
                if (m < PITCH_BUF_SIZE / 4) {
                    int pos = hbase + 4 * m;  // hbase is a multiple of 4: a float4 never straddles the ring wrap
                    if (pos >= HIST_CAP) pos -= HIST_CAP;
                    int pp = hbase + 4 * m - 1;
                    if (pp >= HIST_CAP) pp -= HIST_CAP;
                    if (onA) {
                        va[k] = __ldg(reinterpret_cast<const float4*>(hA + pos));
                        if (m > 0) qa[k] = __ldg(hA + pp);
                    }
                    if (onB) {
                        vb[k] = __ldg(reinterpret_cast<const float4*>(hB + pos));
                        if (m > 0) qb[k] = __ldg(hB + pp);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NQ; k++) {
                const int m = lane + 32 * (NQ * hh + k);
                if (m < PITCH_BUF_SIZE / 4) {
                    const u64 x0 = pk(va[k].x, vb[k].x), x1 = pk(va[k].y, vb[k].y), x2 = pk(va[k].z, vb[k].z), x3 = pk(va[k].w, vb[k].w);
                    const u64 o0 = mul2(add2(mul2(add2(pk(qa[k], qb[k]), x1), half), x0), half);
                    const u64 o1 = mul2(add2(mul2(add2(x1, x3), half), x2), half);
                    st2(prow + 2 * m, o0, o1);
                }
            }
        }
        if (lane < 4) {
            prow[PB + lane] = 0ull;
            Y4[r * Y4_LD + PB / 2 + lane] = 0ull;
        }
    }
    if (tid < 4) CTR[tid] = 0;
    __syncthreads();
    PPROF(0);

    // ---- Ph2: celt_autocorr, warp k = lag k, lane = pair ----
    for (int k = warp; k < 5; k += NW) {
        const u64* row = P + lp * P_LD;
        u64 v;
        switch (k) {
            case 0: v = autocorr_lag<0>(row); break;
            case 1: v = autocorr_lag<1>(row); break;
            case 2: v = autocorr_lag<2>(row); break;
            case 3: v = autocorr_lag<3>(row); break;
            default: v = autocorr_lag<4>(row); break;
        }
        AC[k * SP + lp] = v;
    }
    __syncthreads();
    PPROF(1);

    // ---- Ph3: noise floor, lag window, LPC(4), bandwidth expansion, extra zero (src/pitch.rs:462-480, 257-292);
    // data-dependent early exit: lane = stream ----
    if (warp == 0) {
        const float* ACf = reinterpret_cast<const float*>(AC);
        float* LPCf = reinterpret_cast<float*>(LPC);
        float a[5];
#pragma unroll
        for (int i = 0; i < 5; i++) a[i] = ACf[i * SB + ls];  // (i * SP + pair) * 2 + half = i * SB + stream
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
        LPCf[0 * SB + ls] = fa(lpc[0], 0.8f);
        LPCf[1 * SB + ls] = fa(lpc[1], fm(0.8f, lpc[0]));
        LPCf[2 * SB + ls] = fa(lpc[2], fm(0.8f, lpc[1]));
        LPCf[3 * SB + ls] = fa(lpc[3], fm(0.8f, lpc[2]));
        LPCf[4 * SB + ls] = fm(0.8f, lpc[3]);
    }
    __syncthreads();
    PPROF(2);

    // ---- Ph4: fir5_in_place (src/pitch.rs:407-429) + second decimation (src/pitch.rs:74-79).  One warp per pair
    // row, four samples per lane, 128-sample rounds from the END of the row backwards, so the five older inputs a
    // round needs are still un-filtered when it runs. ----
    for (int r = warp; r < SP; r += NW) {
        u64* prow = P + r * P_LD;
        const u64 nc[5] = {LPC[0 * SP + r], LPC[1 * SP + r], LPC[2 * SP + r], LPC[3 * SP + r], LPC[4 * SP + r]};
        for (int rd = (PB / 4 + 31) / 32 - 1; rd >= 0; rd--) {
            const int q = 32 * rd + lane;
            const bool on = q < PB / 4;
            u64 e[9];
            u64 o[4];
            if (on) {
                ld2(prow + 4 * q, e[5], e[6]);
                ld2(prow + 4 * q + 2, e[7], e[8]);
                if (q >= 1) {
                    ld2(prow + 4 * q - 4, e[1], e[2]);
                    ld2(prow + 4 * q - 2, e[3], e[4]);
                } else {
                    e[1] = e[2] = e[3] = e[4] = 0ull;
                }
                e[0] = q >= 2 ? prow[4 * q - 5] : 0ull;
#pragma unroll
                for (int d = 0; d < 4; d++)
                    o[d] = add2(add2(add2(add2(add2(e[5 + d], mul2(nc[0], e[4 + d])), mul2(nc[1], e[3 + d])), mul2(nc[2], e[2 + d])),
                                     mul2(nc[3], e[1 + d])),
                                mul2(nc[4], e[d]));
            }
            __syncwarp();
            if (on) {
                st2(prow + 4 * q, o[0], o[1]);
                st2(prow + 4 * q + 2, o[2], o[3]);
                st2(Y4 + r * Y4_LD + 2 * q, o[0], o[2]);
            }
        }
    }
    __syncthreads();
    PPROF(3);

    // ---- Ph5: coarse running energy (warp NW-2) + xx (warp NW-1), then coarse xcorr on all warps ----
    if (warp == NW - 2) {
        // y_sq_norm of find_best_pitch(xcorr, y_lp4, 240) (src/pitch.rs:379-382, 401-402); YN4[i] = value seen at lag i
        const u64* row = Y4 + lp * Y4_LD;
        u64 y = splat(1.0f);
#pragma unroll 4
        for (int m = 0; m < N4 / 2; m++) {
            u64 v0, v1;
            ld2(row + 2 * m, v0, v1);
            y = add2(y, mul2(v0, v0));
            y = add2(y, mul2(v1, v1));
        }
        u64* out = YN4 + lp * XC_LD;
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < (NL4 + 1) / 2; m++) {
            u64 a0, a1, b0, b1;
            ld2(row + N4 + 2 * m, a0, a1);
            ld2(row + 2 * m, b0, b1);
            y = max2(add2(y, sub2(mul2(a0, a0), mul2(b0, b0))), 1.0f);
            if (2 * m + 1 < XC_LD) out[2 * m + 1] = y;
            y = max2(add2(y, sub2(mul2(a1, a1), mul2(b1, b1))), 1.0f);
            if (2 * m + 2 < XC_LD) out[2 * m + 2] = y;
        }
    } else if (warp == NW - 1) {
        // xx = inner_prod(x, x, 480) with its four interleaved accumulators (src/pitch.rs:133, 225-244)
        const u64* xr = P + lp * P_LD + HALF_MAX;
        u64 a0 = splat(0.0f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll 4
        for (int m = 0; m < HALF_N / 4; m++) {
            u64 x0, x1, x2, x3;
            ld2(xr + 4 * m, x0, x1);
            ld2(xr + 4 * m + 2, x2, x3);
            a0 = add2(a0, mul2(x0, x0));
            a1 = add2(a1, mul2(x1, x1));
            a2 = add2(a2, mul2(x2, x2));
            a3 = add2(a3, mul2(x3, x3));
        }
        XX[lp] = add2(add2(add2(a0, a1), a2), a3);
    }
    // coarse xcorr (src/pitch.rs:82, 296-363): lane-task = (pair, group of 4 consecutive lags); every accumulator
    // sums x_lp4[j] * y_lp4[lag + j] with j ascending, operands via a sliding register window.
    for (;;) {
        int T = 0;
        if (lane == 0) T = atomicAdd(&CTR[0], 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T * 32 >= SP * NGRP) break;
        const int L = T * 32 + lane;
        if (L < SP * NGRP) {
            const int s = L / NGRP, g = L - s * NGRP;
            const u64* xr = Y4 + s * Y4_LD + HALF_MAX / 2;
            const u64* yr = Y4 + s * Y4_LD + 4 * g;
            u64 c0 = splat(0.0f), c1 = c0, c2 = c0, c3 = c0;
            u64 e[8];
            ld2(yr, e[0], e[1]);
            ld2(yr + 2, e[2], e[3]);
#pragma unroll 2
            for (int m = 0; m < N4 / 4; m++) {
                u64 xv[4];
                ld2(xr + 4 * m, xv[0], xv[1]);
                ld2(xr + 4 * m + 2, xv[2], xv[3]);
                ld2(yr + 4 * m + 4, e[4], e[5]);
                ld2(yr + 4 * m + 6, e[6], e[7]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    c0 = add2(c0, mul2(xv[u], e[u]));
                    c1 = add2(c1, mul2(xv[u], e[u + 1]));
                    c2 = add2(c2, mul2(xv[u], e[u + 2]));
                    c3 = add2(c3, mul2(xv[u], e[u + 3]));
                }
#pragma unroll
                for (int u = 0; u < 4; u++) e[u] = e[4 + u];
            }
            u64* o = XC + s * XC_LD + 4 * g;
            o[0] = c0;
            o[1] = c1;
            o[2] = c2;
            if (4 * g + 3 < NL4) o[3] = c3;
        }
    }
    __syncthreads();
    PPROF(4);

    // ---- Ph6a: warp 0: coarse best/second (serial over lags, lane = stream; src/pitch.rs:83-84).
    // warp 1: fine running energy, lane = pair.  The 4x-decimated copy is dead: YN2 reuses it. ----
    u64* YN2 = Y4;
    if (warp == 0) {
        BestTwo b2;
        const float* xc = reinterpret_cast<const float*>(XC + (ls >> 1) * XC_LD) + (ls & 1);
        const float* yn = reinterpret_cast<const float*>(YN4 + (ls >> 1) * XC_LD) + (ls & 1);
#pragma unroll 7
        for (int i = 0; i < NL4; i++) b2.consider(i, xc[2 * i], yn[2 * i]);
        SI[0 * SB + ls] = b2.best;
        SI[1 * SB + ls] = b2.second;
    } else if (warp == 1) {
        // y_sq_norm of find_best_pitch(xcorr, y, 480): YN2[i] = value seen at fine lag i
        const u64* row = P + lp * P_LD;
        u64 y = splat(1.0f);
#pragma unroll 4
        for (int m = 0; m < HALF_N / 2; m++) {
            u64 v0, v1;
            ld2(row + 2 * m, v0, v1);
            y = add2(y, mul2(v0, v0));
            y = add2(y, mul2(v1, v1));
        }
        u64* out = YN2 + lp * YN2_LD;
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < NL2 / 2; m++) {
            u64 a0, a1, b0, b1;
            ld2(row + HALF_N + 2 * m, a0, a1);
            ld2(row + 2 * m, b0, b1);
            y = max2(add2(y, sub2(mul2(a0, a0), mul2(b0, b0))), 1.0f);
            out[2 * m + 1] = y;
            y = max2(add2(y, sub2(mul2(a1, a1), mul2(b1, b1))), 1.0f);
            out[2 * m + 2] = y;
        }
    }
    __syncthreads();
    PPROF(5);

    // ---- Ph6b: the two 5-lag fine windows of every stream (src/pitch.rs:88-96), each split into a 3-lag and a
    // 2-lag sliding window so that two warps share the work.  lane-task = (pair, window): lags i0c .. i0c+4 of each
    // half, i0c = window start clamped into the valid range; which of them count as candidates is decided in Ph8. ----
    if (warp == 2 || warp == 3) {
        for (int L = lane; L < 2 * SP; L += 32) {
            const int p = L >> 1, wdw = L & 1;
            const int i0a = min(max(2 * SI[wdw * SB + 2 * p] - 2, 0), NL2 - 5);
            const int i0b = min(max(2 * SI[wdw * SB + 2 * p + 1] - 2, 0), NL2 - 5);
            const u64* xr = P + p * P_LD + HALF_MAX;
            const float* rowf = Pf + (size_t)p * P_LD * 2;
            if (warp == 2) {
                u64 out[3];
                inner_prod_window<3>(xr, rowf + 2 * i0a, rowf + 2 * i0b + 1, out);
#pragma unroll
                for (int c = 0; c < 3; c++) FX[p * FX_LD + wdw * 5 + c] = max2(out[c], -1.0f);
            } else {
                u64 out[2];
                inner_prod_window<2>(xr, rowf + 2 * (i0a + 3), rowf + 2 * (i0b + 3) + 1, out);
#pragma unroll
                for (int c = 0; c < 2; c++) FX[p * FX_LD + wdw * 5 + 3 + c] = max2(out[c], -1.0f);
            }
        }
    }
    __syncthreads();
    PPROF(6);

    // ---- Ph8: fine best + pseudo-interpolation (src/pitch.rs:97-114), lane = stream ----
    if (warp == 0) {
        const int best4 = SI[0 * SB + ls], second4 = SI[1 * SB + ls];
        const float* fx = reinterpret_cast<const float*>(FX + (ls >> 1) * FX_LD) + (ls & 1);
        const float* yn = reinterpret_cast<const float*>(YN2 + (ls >> 1) * YN2_LD) + (ls & 1);
        const int cA = 2 * best4, cB = 2 * second4;
        const int baseA = min(max(cA - 2, 0), NL2 - 5), baseB = min(max(cB - 2, 0), NL2 - 5);
        // xcorr at fine lag i: computed iff |i - 2 best| <= 2 or |i - 2 second| <= 2 (src/pitch.rs:90-95), else 0
        auto xcf = [&](int i) -> float {
            if (i < 0 || i >= NL2) return 0.0f;
            if (abs(i - cA) <= 2) return fx[2 * (i - baseA)];
            if (abs(i - cB) <= 2) return fx[2 * (5 + i - baseB)];
            return 0.0f;
        };
        BestTwo b2;
        // lags outside the windows have xcorr 0 and can never be selected: scan the windows in ascending order
        const int c0 = min(cA, cB) - 2, c1 = max(cA, cB) - 2;
        const int lo0 = max(c0, 0), hi0 = min(c0 + 4, NL2 - 1);
        const int lo1 = max(max(c1, 0), hi0 + 1), hi1 = min(c1 + 4, NL2 - 1);
        for (int i = lo0; i <= hi0; i++) b2.consider(i, xcf(i), yn[2 * i]);
        for (int i = lo1; i <= hi1; i++) b2.consider(i, xcf(i), yn[2 * i]);
        const int best = b2.best;
        int offset = 0;
        if (best > 0 && best < NL2 - 1) {
            const float a = xcf(best - 1), b = xcf(best), c = xcf(best + 1);
            if (fs(c, a) > fm(0.7f, fs(b, a))) offset = 1;
            else if (fs(a, c) > fm(0.7f, fs(b, c))) offset = -1;
        }
        const int pitch_idx = PITCH_MAX_PERIOD - (2 * best - offset);  // src/pitch.rs:49,114
        const int t0 = min(pitch_idx / 2, HALF_MAX - 1);                // t0 of remove_doubling
        SI[2 * SB + ls] = t0;
        // Lags of the inner products remove_doubling will need (src/pitch.rs:134,152-168): q = 0: xy(t0); then for
        // k = 2.. while t1 >= min_period: q = 2k-3: t1, q = 2k-2: t1b.  A pair runs max(nA, nB) packed inner products;
        // the shorter half repeats t0 (result unused).
        int nk = 0;
        int* lg = LAG + ls * NQMAX;
        lg[0] = t0;
        for (int k = 2; k <= 15; k++) {
            const int t1 = (2 * t0 + k) / (2 * k);
            const bool liveK = (nk == k - 2) && t1 >= MIN_PERIOD2;
            if (liveK) nk++;
            const int t1b = (k == 2) ? ((t1 + t0 > HALF_MAX) ? t0 : t0 + t1) : (2 * c_second_check2[k] * t0 + k) / (2 * k);
            lg[2 * k - 3] = liveK ? t1 : t0;
            lg[2 * k - 2] = liveK ? t1b : t0;
        }
        const int nko = __shfl_xor_sync(0xffffffffu, nk, 1);
        const bool owner = lane < SB && (lane & 1) == 0;
        const int n = owner ? 1 + 2 * max(nk, nko) : 0;
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) CTR[2] = incl;
        if (owner) {
            int* tk = TASK + (incl - n);
            for (int q = 0; q < n; q++) tk[q] = ((lane >> 1) << 8) | q;  // entry = pair << 8 | q
        }
    }
    __syncthreads();
    PPROF(7);

    // ---- Ph9: yy_lookup chain (warp NW-1 first) + remove_doubling inner products on all warps ----
    u64* YY = Y4;  // yn2 is dead from here on
    if (warp == NW - 1) {
        // yy_lookup (src/pitch.rs:135-142): stored clamped at 0, carried unclamped; i = 1..384 walks the rows downwards
        const u64* row = P + lp * P_LD;
        u64* out = YY + lp * YY_LD;
        u64 y = XX[lp];
        out[0] = y;
#pragma unroll 2
        for (int m = 0; m < HALF_MAX / 2; m++) {
            u64 a0, a1, b0, b1;
            ld2(row + HALF_MAX - 2 - 2 * m, a0, a1);             // p[382-2m], p[383-2m]
            ld2(row + HALF_MAX + HALF_N - 2 - 2 * m, b0, b1);    // p[862-2m], p[863-2m]
            y = add2(y, sub2(mul2(a1, a1), mul2(b1, b1)));
            out[2 * m + 1] = max2(y, 0.0f);
            y = add2(y, sub2(mul2(a0, a0), mul2(b0, b0)));
            out[2 * m + 2] = max2(y, 0.0f);
        }
    }
    // lane-task = one entry of the compacted list (pair, q): IPR[pair][q] = inner_prod(x, x - lag, 480) per half
    {
        const int ntask = CTR[2];
        for (;;) {
            int T = 0;
            if (lane == 0) T = atomicAdd(&CTR[1], 1);
            T = __shfl_sync(0xffffffffu, T, 0);
            if (T * 32 >= ntask) break;
            const int L = T * 32 + lane;
            if (L < ntask) {
                const int e = TASK[L];
                const int p = e >> 8, q = e & 255;
                const int lagA = LAG[(2 * p) * NQMAX + q], lagB = LAG[(2 * p + 1) * NQMAX + q];
                const float* rowf = Pf + (size_t)p * P_LD * 2;
                IPR[p * IPR_LD + q] = inner_prod_480(P + p * P_LD + HALF_MAX, rowf + 2 * (HALF_MAX - lagA), rowf + 2 * (HALF_MAX - lagB) + 1);
            }
        }
    }
    __syncthreads();
    PPROF(8);

    // ---- Ph10-12: the sub-harmonic ladder (src/pitch.rs:144-203, lane = stream), the +-1 refinement (205-218, packed:
    // both lanes of a pair run the pair's window) and the result; no further block-level synchronisation ----
    if (warp == 0) {
        const float* ipr = reinterpret_cast<const float*>(IPR + (ls >> 1) * IPR_LD) + (ls & 1);
        const float* yy = reinterpret_cast<const float*>(YY + (ls >> 1) * YY_LD) + (ls & 1);
        const int t0 = SI[2 * SB + ls];
        const float xx = reinterpret_cast<const float*>(XX)[ls];
        float xy = ipr[0];
        float yyv = yy[2 * t0];
        int prev_period = 0;
        float lg = 0.0f;
        if (ls < ns) {
            prev_period = last_period[s0 + ls] / 2;
            lg = last_gain[s0 + ls];
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
            else t1b = (2 * c_second_check2[k] * t0 + k) / (2 * k);
            xy = fm(fa(ipr[2 * (2 * k - 3)], ipr[2 * (2 * k - 2)]), 0.5f);
            yyv = fm(fa(yy[2 * t1], yy[2 * t1b]), 0.5f);
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

        // xcorr at lags t-1, t, t+1 of both halves: one sliding window starting at lag t+1 (lowest address)
        __syncwarp();
        const int tother = __shfl_xor_sync(0xffffffffu, t, 1);
        const int tA = (ls & 1) ? tother : t, tB = (ls & 1) ? t : tother;
        u64 xc3[3];
        const int p = ls >> 1;
        const float* rowf = Pf + (size_t)p * P_LD * 2;
        inner_prod_window<3>(P + p * P_LD + HALF_MAX, rowf + 2 * (HALF_MAX - (tA + 1)), rowf + 2 * (HALF_MAX - (tB + 1)) + 1, xc3);
        const bool odd = (ls & 1) != 0;
        const float x_0 = odd ? hi(xc3[2]) : lo(xc3[2]), x_1 = odd ? hi(xc3[1]) : lo(xc3[1]), x_2 = odd ? hi(xc3[0]) : lo(xc3[0]);
        int offset = 0;  // window slot c <-> lag t + 1 - c
        if (fs(x_2, x_0) > fm(0.7f, fs(x_1, x_0))) offset = 1;
        else if (fs(x_0, x_2) > fm(0.7f, fs(x_1, x_2))) offset = -1;
        const int tf = max(2 * t + offset, PITCH_MIN_PERIOD);
        if (lane < ns) {
            pitch_out[s0 + lane] = tf;
            last_period[s0 + lane] = tf;
            last_gain[s0 + lane] = pg;
        }
    }
    PPROF(9);
}

}  // namespace

#ifdef PITCH_PROFILE
extern "C" void nnb_pitch2_prof_read(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_pitch2_prof, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_pitch2_prof, z, sizeof z);
    }
}
#endif

cudaError_t launch_pitch2(const BatchBuffers& b, int slot, cudaStream_t st) {
    static unsigned long long attr_devs = 0;  // bit d: attribute set on device d
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((attr_devs >> dev) & 1ull)) {
        e = cudaFuncSetAttribute(pitch2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_devs |= 1ull << dev;
    }
    const int grid = (b.n_streams + SB - 1) / SB;
    pitch2_kernel<<<grid, NT, SMEM_BYTES, st>>>(b.hist, b.last_period, b.last_gain, b.pitch, b.n_streams, hist_base(slot));
    return cudaGetLastError();
}

}  // namespace nnb
