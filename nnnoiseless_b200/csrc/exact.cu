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
// K2: pitch analysis, one block (6 warps) per stream.
//
// The reference's pitch path alternates data-parallel sums with strictly sequential recurrences
// (running energies with a clamp per step, best/second-best selection).  Order-exactness forbids
// re-associating them, so the kernel overlaps them instead: while warps 0-4 compute the coarse
// cross-correlation (one lane per lag, j ascending), warp 5 walks ALL THREE running-energy chains
// (coarse y_sq_norm, fine y_sq_norm, yy_lookup) in SIMT lock-step on three lanes.  The k = 2..15
// sub-harmonic test of remove_doubling has no loop-carried arithmetic, so it is evaluated by 14
// lanes at once and resolved with one ballot.
// ================================================================================================
constexpr int PT = 192;                                                // threads
constexpr int PB = PITCH_BUF_SIZE / 2;                                 // 864
constexpr int MAXP = PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD;          // 588
constexpr int N4 = PITCH_FRAME_SIZE / 4;                               // 240
constexpr int NL4 = MAXP / 4;                                          // 147 coarse lags
constexpr int NL2 = MAXP / 2;                                          // 294 fine lags
constexpr int HALF_MAX = PITCH_MAX_PERIOD / 2;                         // 384
constexpr int HALF_N = PITCH_FRAME_SIZE / 2;                           // 480

__constant__ int c_second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};  // src/pitch.rs:489

// One of the four interleaved accumulators of inner_prod (src/pitch.rs:225-237): terms a, a+4, ...
__device__ __forceinline__ float inner_prod_lane(const float* xs, const float* ys, int n, int a) {
    float s = 0.0f;
#pragma unroll 8
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

// Selection step of find_best_pitch (src/pitch.rs:383-400) for lag i with correlation corr and the
// running energy ysq valid at that lag.
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

__global__ void __launch_bounds__(PT) pitch_kernel(const float* __restrict__ hist, int32_t* __restrict__ last_period,
                                                   float* __restrict__ last_gain, int32_t* __restrict__ pitch_out,
                                                   int hbase) {
    __shared__ float pbuf[PB];           // 2x-decimated, LPC-whitened history (src/pitch.rs pitch_buf)
    __shared__ float pe[PB / 2];         // its even samples = the 4x-decimated signal y_lp4 (x_lp4 = pe + 192)
    __shared__ float xc[NL2 + 2];        // xcorr (147 coarse, then 294 fine)
    __shared__ float yn4[NL4 + 1];       // y_sq_norm seen at coarse lag i
    __shared__ float yn2[NL2 + 2];       // y_sq_norm seen at fine lag i
    __shared__ float yy[HALF_MAX + 4];   // yy_lookup
    __shared__ float ac[8];
    __shared__ float lpc2[8];
    __shared__ float ipr[32];            // inner products of remove_doubling
    __shared__ int s_i[4];

    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
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
#pragma unroll 8
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
#pragma unroll
            for (int i = 0; i < 4; i++) {
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
                if (error < fm(0.001f, a[0])) break;
            }
        }
        float tmp = 1.0f;
#pragma unroll
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
        constexpr int R = (PB + PT - 1) / PT;
        float outv[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
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
        for (int r = 0; r < R; r++) {
            int i = tid + r * PT;
            if (i < PB) {
                pbuf[i] = outv[r];
                if ((i & 1) == 0) pe[i >> 1] = outv[r];  // second decimation (src/pitch.rs:74-79)
            }
        }
    }
    __syncthreads();

    if (warp < 5) {
        // ---- coarse xcorr: one lane per lag, j ascending (src/pitch.rs:82, 296-363) ----
        if (tid < NL4) {
            float c = 0.0f;
            const float* x = pe + HALF_MAX / 2;
            const float* y = pe + tid;
#pragma unroll 8
            for (int j = 0; j < N4; j++) c = fa(c, fm(x[j], y[j]));
            xc[tid] = c;
        }
    } else {
        // ---- warp 5: the three running-energy chains, lock-step on lanes 4 (coarse), 5 (fine), 0 (yy) ----
        // init sums: lanes 0-3 = the four accumulators of xx = inner_prod(x, x, 480); lane 4 = 1 + sum y_lp4^2
        // (240 terms, src/pitch.rs:379-382); lane 5 = 1 + sum y^2 (480 terms)
        const float* ip;
        int ist, icnt;
        float acc;
        if (lane < 4) { ip = pbuf + HALF_MAX + lane; ist = 4; icnt = HALF_N / 4; acc = 0.0f; }
        else if (lane == 4) { ip = pe; ist = 1; icnt = N4; acc = 1.0f; }
        else if (lane == 5) { ip = pbuf; ist = 1; icnt = HALF_N; acc = 1.0f; }
        else { ip = pbuf; ist = 0; icnt = 0; acc = 0.0f; }
#pragma unroll 4
        for (int it = 0; it < HALF_N; it++) {
            if (it < icnt) {
                float v = ip[it * ist];
                acc = fa(acc, fm(v, v));
            }
        }
        const float xx = inner_prod_combine(acc, lane);  // valid on lanes 0-3
        // chains: y <- y + (a^2 - b^2); coarse/fine clamp the carried value at 1 (src/pitch.rs:401-402),
        // yy_lookup stores max(y, 0) but carries y unclamped (src/pitch.rs:138-142)
        const float *pa, *pb;
        float* po;
        int cst, ccnt, clamp1;
        float y;
        if (lane == 4) { pa = pe + N4; pb = pe; po = yn4; cst = 1; ccnt = NL4; clamp1 = 1; y = acc; }
        else if (lane == 5) { pa = pbuf + HALF_N; pb = pbuf; po = yn2; cst = 1; ccnt = NL2; clamp1 = 1; y = acc; }
        else if (lane == 0) { pa = pbuf + HALF_MAX - 1; pb = pbuf + HALF_MAX + HALF_N - 1; po = yy; cst = -1; ccnt = HALF_MAX; clamp1 = 0; y = xx; }
        else { pa = pbuf; pb = pbuf; po = yy; cst = 0; ccnt = 0; clamp1 = 0; y = 0.0f; }
        if (ccnt > 0) po[0] = y;
        if (lane == 0) ipr[0] = xx;
#pragma unroll 4
        for (int it = 0; it < HALF_MAX; it++) {
            if (it < ccnt) {
                float a = pa[it * cst], b = pb[it * cst];
                float yn = fa(y, fs(fm(a, a), fm(b, b)));
                float st = clamp1 ? fmaxf(yn, 1.0f) : fmaxf(yn, 0.0f);
                po[it + 1] = st;
                y = clamp1 ? st : yn;
            }
        }
    }
    __syncthreads();

    // ---- coarse best / second best (src/pitch.rs:83-84, 372-405), serial over the 147 lags ----
    if (tid == 0) {
        BestTwo b2;
#pragma unroll 4
        for (int i = 0; i < NL4; i++) b2.consider(i, xc[i], yn4[i]);
        s_i[0] = b2.best;
        s_i[1] = b2.second;
    }
    __syncthreads();

    // ---- fine search around the two candidates (src/pitch.rs:88-96) ----
    const int best4 = s_i[0], second4 = s_i[1];
    __syncthreads();
    for (int i = tid; i < NL2; i += PT) xc[i] = 0.0f;
    __syncthreads();
    if (tid < 64) {  // two full warps so the shuffles are convergent
        int c = tid >> 2, a = tid & 3;
        int i = (c < 5) ? (2 * best4 - 2 + c) : (2 * second4 - 2 + (c - 5));
        bool valid = (c < 10) && i >= 0 && i < NL2;
        float sacc = 0.0f;
        if (valid) sacc = inner_prod_lane(pbuf + HALF_MAX, pbuf + i, HALF_N, a);
        float sum = inner_prod_combine(sacc, lane);
        if (valid && a == 0) xc[i] = fmaxf(sum, -1.0f);
    }
    __syncthreads();

    if (tid == 0) {
        // find_best_pitch over the 294 fine lags: every lag outside the two 5-wide windows has xcorr 0 and
        // cannot be selected (corr > 0 fails), so visiting the windows in ascending order is the same scan.
        BestTwo b2;
        int c0 = 2 * min(best4, second4), c1 = 2 * max(best4, second4);
        int lo0 = max(c0 - 2, 0), hi0 = min(c0 + 2, NL2 - 1);
        int lo1 = max(max(c1 - 2, 0), hi0 + 1), hi1 = min(c1 + 2, NL2 - 1);
        for (int i = lo0; i <= hi0; i++) b2.consider(i, xc[i], yn2[i]);
        for (int i = lo1; i <= hi1; i++) b2.consider(i, xc[i], yn2[i]);
        const int best = b2.best;
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
    // inner products: q = 1: xy(t0); q = 2 + 2(k-2) + {0,1}: t1(k), t1b(k), k = 2..15   (q = 0, xx, is in ipr[0])
    if (tid < 128) {  // four full warps
        int q = tid >> 2, a = tid & 3;
        int lagq = -1;
        if (q == 1) lagq = t0;
        else if (q >= 2 && q < 30) {
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
        if (a == 0 && q >= 1 && q < 30) ipr[q] = sum;
    }
    __syncthreads();

    if (tid < 32) {
        // lane k (2..15) evaluates sub-harmonic k; nothing in the reference's loop body depends on earlier
        // iterations except "the last k that passes wins", which a ballot resolves.
        const float xx = ipr[0], xy0 = ipr[1];
        const float yy0 = yy[t0];
        const float g0 = pitch_gain(xy0, xx, yy0);
        const int prev_period = last_period[s] / 2;
        const float lg = last_gain[s];
        const int k = lane;
        bool pass = false;
        float xy = 0.0f, yyv = 0.0f, g1 = 0.0f;
        int t1 = 0;
        if (k >= 2 && k <= 15) {
            t1 = (2 * t0 + k) / (2 * k);
            if (t1 >= min_period) {
                int t1b;
                if (k == 2) t1b = (t1 + t0 > HALF_MAX) ? t0 : t0 + t1;
                else t1b = (2 * c_second_check[k] * t0 + k) / (2 * k);
                xy = fm(fa(ipr[2 + 2 * (k - 2)], ipr[3 + 2 * (k - 2)]), 0.5f);
                yyv = fm(fa(yy[t1], yy[t1b]), 0.5f);
                g1 = pitch_gain(xy, xx, yyv);
                int d = abs(t1 - prev_period);
                float cont;
                if (d <= 1) cont = lg;
                else if (d <= 2 && 5 * k * k < t0) cont = fm(lg, 0.5f);
                else cont = 0.0f;
                float thresh;
                if (t1 < 3 * min_period) thresh = fmaxf(fs(fm(0.85f, g0), cont), 0.4f);
                else if (t1 < 2 * min_period) thresh = fmaxf(fs(fm(0.9f, g0), cont), 0.5f);  // dead branch, as in the reference
                else thresh = fmaxf(fs(fm(0.7f, g0), cont), 0.3f);
                pass = g1 > thresh;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        float best_xy = xy0, best_yy = yy0, g = g0;
        int t = t0;
        if (mask) {
            const int kk = 31 - __clz(mask);
            best_xy = __shfl_sync(0xffffffffu, xy, kk);
            best_yy = __shfl_sync(0xffffffffu, yyv, kk);
            g = __shfl_sync(0xffffffffu, g1, kk);
            t = __shfl_sync(0xffffffffu, t1, kk);
        } else {
            // keep the shuffles convergent
            (void)__shfl_sync(0xffffffffu, xy, 0); (void)__shfl_sync(0xffffffffu, yyv, 0);
            (void)__shfl_sync(0xffffffffu, g1, 0); (void)__shfl_sync(0xffffffffu, t1, 0);
        }
        best_xy = fmaxf(best_xy, 0.0f);
        float pg = (best_yy <= best_xy) ? 1.0f : __fdiv_rn(best_xy, fa(best_yy, 1.0f));
        pg = fminf(pg, g);

        // final +-1 refinement (src/pitch.rs:205-218)
        int c = lane >> 2, a = lane & 3;
        float sacc = 0.0f;
        if (c < 3) sacc = inner_prod_lane(x0, x0 - (t + c - 1), HALF_N, a);
        float sum = inner_prod_combine(sacc, lane);
        float x_0 = __shfl_sync(0xffffffffu, sum, 0);
        float x_1 = __shfl_sync(0xffffffffu, sum, 4);
        float x_2 = __shfl_sync(0xffffffffu, sum, 8);
        if (lane == 0) {
            int offset = 0;
            if (fs(x_2, x_0) > fm(0.7f, fs(x_1, x_0))) offset = 1;
            else if (fs(x_0, x_2) > fm(0.7f, fs(x_1, x_2))) offset = -1;
            int tf = max(2 * t + offset, PITCH_MIN_PERIOD);
            pitch_out[s] = tf;
            last_period[s] = tf;
            last_gain[s] = pg;
        }
    }
}

cudaError_t launch_pitch(const BatchBuffers& b, int slot, cudaStream_t st) {
    pitch_kernel<<<b.n_streams, PT, 0, st>>>(b.hist, b.last_period, b.last_gain, b.pitch, hist_base(slot));
    return cudaGetLastError();
}

}  // namespace nnb
