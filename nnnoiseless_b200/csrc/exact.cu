// exact.cu -- the ORDER-EXACT part of the path: high-pass biquad + pitch analysis.
//
// Compiled with -fmad=false and written with explicit round-to-nearest intrinsics so that every
// f32/f64 operation is rounded exactly like the reference's scalar Rust code and summed in the
// same order.  Result: the pitch period (an integer) is bit-identical to the reference's
// restatement (oracle) for every frame.
//
// Reference: src/features.rs:97-110 (shift_and_filter_input, find_pitch), src/util.rs:68-107
// (Biquad), src/pitch.rs:45-489 (PitchFinder and helpers).
#include "common.cuh"

namespace nnb {

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }

// ================================================================================================
// K1: high-pass biquad, one lane per stream (src/util.rs:95-107: f64 arithmetic, f32 state).
// 32 streams per block; the [32][480] tile is staged through shared memory so that global
// traffic is coalesced 128-bit while the serial recurrence walks rows conflict-free (stride 481).
// ================================================================================================
constexpr int HP_STREAMS = 32;
constexpr int HP_THREADS = 128;
constexpr int HP_LD = FRAME_SIZE + 1;

__global__ void __launch_bounds__(HP_THREADS) hp_filter_kernel(const float* __restrict__ in, long stream_stride,
                                                               float* __restrict__ hist, float* __restrict__ hp_mem,
                                                               int n_streams, int slot, int vec_ok) {
    extern __shared__ float tile[];  // [HP_STREAMS][HP_LD]
    const int s0 = blockIdx.x * HP_STREAMS;
    const int tid = threadIdx.x;
    const int ns = min(HP_STREAMS, n_streams - s0);

    if (vec_ok) {
        for (int idx = tid; idx < ns * (FRAME_SIZE / 4); idx += HP_THREADS) {
            int row = idx / (FRAME_SIZE / 4), c4 = idx % (FRAME_SIZE / 4);
            float4 v = __ldg(reinterpret_cast<const float4*>(in + (long)(s0 + row) * stream_stride) + c4);
            float* t = tile + row * HP_LD + 4 * c4;
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        }
    } else {
        for (int idx = tid; idx < ns * FRAME_SIZE; idx += HP_THREADS) {
            int row = idx / FRAME_SIZE, c = idx % FRAME_SIZE;
            tile[row * HP_LD + c] = in[(long)(s0 + row) * stream_stride + c];
        }
    }
    __syncthreads();

    if (tid < ns) {
        const double a0 = (double)-1.99599f, a1 = (double)0.99600f, b0 = (double)-2.0f, b1 = (double)1.0f;
        float m0 = hp_mem[2 * (s0 + tid)], m1 = hp_mem[2 * (s0 + tid) + 1];
        float* row = tile + tid * HP_LD;
#pragma unroll 4
        for (int i = 0; i < FRAME_SIZE; i++) {
            double x64 = (double)row[i];
            double y64 = __dadd_rn(x64, (double)m0);
            double t0 = __dsub_rn(__dmul_rn(b0, x64), __dmul_rn(a0, y64));
            double t1 = __dsub_rn(__dmul_rn(b1, x64), __dmul_rn(a1, y64));
            m0 = __double2float_rn(__dadd_rn((double)m1, t0));
            m1 = __double2float_rn(t1);
            row[i] = __double2float_rn(y64);
        }
        hp_mem[2 * (s0 + tid)] = m0;
        hp_mem[2 * (s0 + tid) + 1] = m1;
    }
    __syncthreads();

    // hist rows are 16-byte aligned (HIST_CAP*4 and slot*480*4 are multiples of 16)
    for (int idx = tid; idx < ns * (FRAME_SIZE / 4); idx += HP_THREADS) {
        int row = idx / (FRAME_SIZE / 4), c4 = idx % (FRAME_SIZE / 4);
        const float* t = tile + row * HP_LD + 4 * c4;
        float4 v = make_float4(t[0], t[1], t[2], t[3]);
        reinterpret_cast<float4*>(hist + (size_t)(s0 + row) * HIST_CAP + slot * FRAME_SIZE)[c4] = v;
    }
}

cudaError_t launch_hp_filter(const BatchBuffers& b, const float* in, long stream_stride, int slot, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(float) * HP_STREAMS * HP_LD;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(hp_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int vec_ok = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (stream_stride % 4 == 0);
    int grid = (b.n_streams + HP_STREAMS - 1) / HP_STREAMS;
    hp_filter_kernel<<<grid, HP_THREADS, smem, st>>>(in, stream_stride, b.hist, b.hp_mem, b.n_streams, slot, vec_ok);
    return cudaGetLastError();
}

// ================================================================================================
// K2: pitch analysis, one block (5 warps) per stream.
// ================================================================================================
constexpr int PT = 160;                                                // threads
constexpr int PB = PITCH_BUF_SIZE / 2;                                 // 864
constexpr int MAXP = PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD;          // 588
constexpr int N4 = PITCH_FRAME_SIZE / 4;                               // 240
constexpr int NY4 = N4 + MAXP / 4;                                     // 387
constexpr int NL4 = MAXP / 4;                                          // 147 coarse lags
constexpr int NL2 = MAXP / 2;                                          // 294 fine lags
constexpr int HALF_MAX = PITCH_MAX_PERIOD / 2;                         // 384
constexpr int HALF_N = PITCH_FRAME_SIZE / 2;                           // 480

__constant__ int c_second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};  // src/pitch.rs:489

// src/pitch.rs:372-405 (serial by construction: running energy with a clamp at every step)
__device__ void find_best_pitch(const float* xcorr, int nlag, const float* ys, int len, int* best, int* second) {
    float best_num = -1.0f, second_best_num = -1.0f;
    float best_den = 0.0f, second_best_den = 0.0f;
    int best_pitch = 0, second_best_pitch = 1;
    float y_sq_norm = 1.0f;
    for (int j = 0; j < len; j++) y_sq_norm = fa(y_sq_norm, fm(ys[j], ys[j]));
    for (int i = 0; i < nlag; i++) {
        float corr = xcorr[i];
        if (corr > 0.0f) {
            float num = fm(corr, corr);
            if (fm(num, second_best_den) > fm(second_best_num, y_sq_norm)) {
                if (fm(num, best_den) > fm(best_num, y_sq_norm)) {
                    second_best_num = best_num;
                    second_best_den = best_den;
                    second_best_pitch = best_pitch;
                    best_num = num;
                    best_den = y_sq_norm;
                    best_pitch = i;
                } else {
                    second_best_num = num;
                    second_best_den = y_sq_norm;
                    second_best_pitch = i;
                }
            }
        }
        float a = ys[i + len], b = ys[i];
        y_sq_norm = fa(y_sq_norm, fs(fm(a, a), fm(b, b)));
        y_sq_norm = fmaxf(y_sq_norm, 1.0f);
    }
    *best = best_pitch;
    *second = second_best_pitch;
}

// One of the four interleaved accumulators of inner_prod (src/pitch.rs:225-237): terms a, a+4, ...
__device__ __forceinline__ float inner_prod_lane(const float* xs, const float* ys, int n, int a) {
    float s = 0.0f;
    for (int i = a; i < n; i += 4) s = fa(s, fm(xs[i], ys[i]));
    return s;
}
// combine as ((s0+s1)+s2)+s3 (src/pitch.rs:239); the four partial sums live in 4 adjacent lanes
__device__ __forceinline__ float inner_prod_combine(float s, int lane) {
    int base = lane & ~3;
    float s0 = __shfl_sync(0xffffffffu, s, base);
    float s1 = __shfl_sync(0xffffffffu, s, base + 1);
    float s2 = __shfl_sync(0xffffffffu, s, base + 2);
    float s3 = __shfl_sync(0xffffffffu, s, base + 3);
    return fa(fa(fa(s0, s1), s2), s3);
}

__device__ __forceinline__ float pitch_gain(float xy, float xx, float yy) {
    return __fdiv_rn(xy, __fsqrt_rn(fa(1.0f, fm(xx, yy))));  // src/pitch.rs:485-487
}

__global__ void __launch_bounds__(PT) pitch_kernel(const float* __restrict__ hist, int32_t* __restrict__ last_period,
                                                   float* __restrict__ last_gain, int32_t* __restrict__ pitch_out,
                                                   int hbase) {
    __shared__ float pbuf[PB];
    __shared__ float xlp4[N4];
    __shared__ float ylp4[NY4 + 1];
    __shared__ float xc[NL2 + 2];
    __shared__ float yy[HALF_MAX + 4];
    __shared__ float ac[8];
    __shared__ float lpc2[8];
    __shared__ float ipr[32];  // inner products of remove_doubling
    __shared__ int s_i[4];

    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const float* h = hist + (size_t)s * HIST_CAP;

    // ---- pitch_downsample part 1 (src/pitch.rs:455-458) ----
    for (int i = tid; i < PB; i += PT) {
        int p1 = hbase + 2 * i;  // ring position of x[2i]
        int pa = p1 - 1, pb = p1 + 1;
        if (p1 >= HIST_CAP) p1 -= HIST_CAP;
        if (pa >= HIST_CAP) pa -= HIST_CAP;
        if (pb >= HIST_CAP) pb -= HIST_CAP;
        float v;
        if (i == 0) {
            v = fm(fa(fm(h[pb], 0.5f), h[p1]), 0.5f);
        } else {
            v = fm(fa(fm(fa(h[pa], h[pb]), 0.5f), h[p1]), 0.5f);
        }
        pbuf[i] = v;
    }
    __syncthreads();

    // ---- celt_autocorr, 5 lags, sequential sums (src/pitch.rs:433-446 + 296-363) ----
    if (tid < 5) {
        const int k = tid, fast_n = PB - 4;
        float c = 0.0f;
        for (int j = 0; j < fast_n; j++) c = fa(c, fm(pbuf[j], pbuf[j + k]));
        float d = 0.0f;
        for (int i = k + fast_n; i < PB; i++) d = fa(d, fm(pbuf[i], pbuf[i - k]));
        ac[k] = fa(c, d);
    }
    __syncthreads();

    // ---- noise floor, lag window, LPC(4), bandwidth expansion, add a zero (src/pitch.rs:462-480, 257-292) ----
    if (tid == 0) {
        float a[5];
        for (int i = 0; i < 5; i++) a[i] = ac[i];
        a[0] = fm(a[0], 1.0001f);
        for (int i = 1; i < 5; i++) {
            float w = fm(0.008f, (float)i);
            a[i] = fs(a[i], fm(fm(a[i], w), w));
        }
        float lpc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (a[0] != 0.0f) {
            float error = a[0];
            for (int i = 0; i < 4; i++) {
                float rr = 0.0f;
                for (int j = 0; j < i; j++) rr = fa(rr, fm(lpc[j], a[i - j]));
                rr = fa(rr, a[i + 1]);
                float r = __fdiv_rn(-rr, error);
                lpc[i] = r;
                for (int j = 0; j < (i + 1) / 2; j++) {
                    float t1 = lpc[j], t2 = lpc[i - 1 - j];
                    lpc[j] = fa(t1, fm(r, t2));
                    lpc[i - 1 - j] = fa(t2, fm(r, t1));
                }
                error = fs(error, fm(fm(r, r), error));
                if (error < fm(0.001f, a[0])) break;
            }
        }
        float tmp = 1.0f;
        for (int i = 0; i < 4; i++) {
            tmp = fm(tmp, 0.9f);
            lpc[i] = fm(lpc[i], tmp);
        }
        lpc2[0] = fa(lpc[0], 0.8f);
        lpc2[1] = fa(lpc[1], fm(0.8f, lpc[0]));
        lpc2[2] = fa(lpc[2], fm(0.8f, lpc[1]));
        lpc2[3] = fa(lpc[3], fm(0.8f, lpc[2]));
        lpc2[4] = fm(0.8f, lpc[3]);
    }
    __syncthreads();

    // ---- fir5_in_place (src/pitch.rs:407-429): zero initial memory, out-of-place through registers ----
    {
        const float n0 = lpc2[0], n1 = lpc2[1], n2 = lpc2[2], n3 = lpc2[3], n4 = lpc2[4];
        float outv[(PB + PT - 1) / PT];
#pragma unroll
        for (int r = 0; r < (PB + PT - 1) / PT; r++) {
            int i = tid + r * PT;
            float o = 0.0f;
            if (i < PB) {
                float m0 = i >= 1 ? pbuf[i - 1] : 0.0f;
                float m1 = i >= 2 ? pbuf[i - 2] : 0.0f;
                float m2 = i >= 3 ? pbuf[i - 3] : 0.0f;
                float m3 = i >= 4 ? pbuf[i - 4] : 0.0f;
                float m4 = i >= 5 ? pbuf[i - 5] : 0.0f;
                o = fa(fa(fa(fa(fa(pbuf[i], fm(n0, m0)), fm(n1, m1)), fm(n2, m2)), fm(n3, m3)), fm(n4, m4));
            }
            outv[r] = o;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < (PB + PT - 1) / PT; r++) {
            int i = tid + r * PT;
            if (i < PB) pbuf[i] = outv[r];
        }
    }
    __syncthreads();

    // ---- pitch_search: second decimation (src/pitch.rs:74-79) ----
    for (int j = tid; j < N4; j += PT) xlp4[j] = pbuf[HALF_MAX + 2 * j];
    for (int j = tid; j < NY4; j += PT) ylp4[j] = pbuf[2 * j];
    __syncthreads();

    // ---- coarse xcorr: one lane per lag, j ascending (src/pitch.rs:82, 296-363) ----
    if (tid < NL4) {
        float c = 0.0f;
        const float* y = ylp4 + tid;
#pragma unroll 8
        for (int j = 0; j < N4; j++) c = fa(c, fm(xlp4[j], y[j]));
        xc[tid] = c;
    }
    __syncthreads();

    if (tid == 0) {
        int b, sc;
        find_best_pitch(xc, NL4, ylp4, N4, &b, &sc);
        s_i[0] = b;
        s_i[1] = sc;
    }
    __syncthreads();

    // ---- fine search around the two candidates (src/pitch.rs:88-96) ----
    {
        const int best = s_i[0], second = s_i[1];
        __syncthreads();
        for (int i = tid; i < NL2; i += PT) xc[i] = 0.0f;
        __syncthreads();
        if (tid < 64) {  // two full warps so the shuffles below are convergent
            int c = tid >> 2, a = tid & 3;
            int i = (c < 5) ? (2 * best - 2 + c) : (2 * second - 2 + (c - 5));
            bool valid = (c < 10) && i >= 0 && i < NL2;
            float sacc = 0.0f;
            if (valid) sacc = inner_prod_lane(pbuf + HALF_MAX, pbuf + i, HALF_N, a);
            float sum = inner_prod_combine(sacc, lane);
            if (valid && a == 0) xc[i] = fmaxf(sum, -1.0f);
        }
    }
    __syncthreads();

    if (tid == 0) {
        int best, dummy;
        find_best_pitch(xc, NL2, pbuf, HALF_N, &best, &dummy);
        int offset = 0;
        if (best > 0 && best < NL2 - 1) {
            float a = xc[best - 1], b = xc[best], c = xc[best + 1];
            if (fs(c, a) > fm(0.7f, fs(b, a))) offset = 1;
            else if (fs(a, c) > fm(0.7f, fs(b, c))) offset = -1;
        }
        s_i[2] = PITCH_MAX_PERIOD - (2 * best - offset);  // src/pitch.rs:49,114
    }
    __syncthreads();

    // ---- remove_doubling (src/pitch.rs:118-221) ----
    const int pitch_idx = s_i[2];
    const int t0 = min(pitch_idx / 2, HALF_MAX - 1);
    const int min_period = PITCH_MIN_PERIOD / 2;
    const float* x0 = pbuf + HALF_MAX;
    // inner products: q = 0: xx; q = 1: xy(t0); q = 2 + 2(k-2) + {0,1}: t1(k), t1b(k), k = 2..15
    if (tid < 128) {  // four full warps
        int q = tid >> 2, a = tid & 3;
        int lagq = -1;
        if (q == 0) lagq = 0;
        else if (q == 1) lagq = t0;
        else if (q < 30) {
            int k = 2 + ((q - 2) >> 1);
            int t1 = (2 * t0 + k) / (2 * k);
            if (t1 >= min_period) {
                if (((q - 2) & 1) == 0) lagq = t1;
                else if (k == 2) lagq = (t1 + t0 > HALF_MAX) ? t0 : t0 + t1;
                else lagq = (2 * c_second_check[k] * t0 + k) / (2 * k);
            }
        }
        float sacc = 0.0f;
        if (lagq >= 0) sacc = inner_prod_lane(x0, x0 - lagq, HALF_N, a);
        float sum = inner_prod_combine(sacc, lane);
        if (a == 0 && q < 30) ipr[q] = sum;
    }
    __syncthreads();

    if (tid == 0) {
        const float xx = ipr[0];
        float xy = ipr[1];
        // yy_lookup (src/pitch.rs:135-142): running energy, stored clamped, carried unclamped
        yy[0] = xx;
        float yyv = xx;
        for (int i = 1; i <= HALF_MAX; i++) {
            float a = x0[-i], b = x0[HALF_N - i];
            yyv = fa(yyv, fs(fm(a, a), fm(b, b)));
            yy[i] = fmaxf(yyv, 0.0f);
        }
        const int prev_period = last_period[s] / 2;
        const float lg = last_gain[s];
        yyv = yy[t0];
        float best_xy = xy, best_yy = yyv;
        const float g0 = pitch_gain(xy, xx, yyv);
        float g = g0;
        int t = t0;
        for (int k = 2; k <= 15; k++) {
            int t1 = (2 * t0 + k) / (2 * k);
            if (t1 < min_period) break;
            int t1b;
            if (k == 2) t1b = (t1 + t0 > HALF_MAX) ? t0 : t0 + t1;
            else t1b = (2 * c_second_check[k] * t0 + k) / (2 * k);
            xy = fm(fa(ipr[2 + 2 * (k - 2)], ipr[3 + 2 * (k - 2)]), 0.5f);
            yyv = fm(fa(yy[t1], yy[t1b]), 0.5f);
            float g1 = pitch_gain(xy, xx, yyv);
            int d = abs(t1 - prev_period);
            float cont;
            if (d <= 1) cont = lg;
            else if (d <= 2 && 5 * k * k < t0) cont = fm(lg, 0.5f);
            else cont = 0.0f;
            float thresh;
            if (t1 < 3 * min_period) thresh = fmaxf(fs(fm(0.85f, g0), cont), 0.4f);
            else if (t1 < 2 * min_period) thresh = fmaxf(fs(fm(0.9f, g0), cont), 0.5f);  // dead branch, as in the reference
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
        s_i[3] = t;
        ac[7] = pg;
    }
    __syncthreads();

    if (tid < 32) {
        const int t = s_i[3];
        int c = tid >> 2, a = tid & 3;
        float sacc = 0.0f;
        if (c < 3) sacc = inner_prod_lane(x0, x0 - (t + c - 1), HALF_N, a);
        float sum = inner_prod_combine(sacc, lane);
        float x_0 = __shfl_sync(0xffffffffu, sum, 0);
        float x_1 = __shfl_sync(0xffffffffu, sum, 4);
        float x_2 = __shfl_sync(0xffffffffu, sum, 8);
        if (tid == 0) {
            int offset = 0;
            if (fs(x_2, x_0) > fm(0.7f, fs(x_1, x_0))) offset = 1;
            else if (fs(x_0, x_2) > fm(0.7f, fs(x_1, x_2))) offset = -1;
            int tf = max(2 * t + offset, PITCH_MIN_PERIOD);
            pitch_out[s] = tf;
            last_period[s] = tf;
            last_gain[s] = ac[7];
        }
    }
}

cudaError_t launch_pitch(const BatchBuffers& b, int slot, cudaStream_t st) {
    pitch_kernel<<<b.n_streams, PT, 0, st>>>(b.hist, b.last_period, b.last_gain, b.pitch, hist_base(slot));
    return cudaGetLastError();
}

}  // namespace nnb
