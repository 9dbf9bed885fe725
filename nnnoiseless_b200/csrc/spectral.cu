// spectral.cu -- frequency-domain half of the path: windowed 960-point real FFTs of the current
// and pitch-lagged windows, Bark-band energies/correlations, the 42 input features, and (after the
// RNN) pitch filtering, band-gain interpolation, inverse FFT and overlap-add.
//
// Reference: src/features.rs:115-298, src/lib.rs:65-162, src/denoise.rs:95-116.
// f32 with FMA contraction allowed: these stages are compared to the oracle within tolerance.
#include "common.cuh"

namespace nnb {

constexpr int ST = 128;  // threads per block (one block per stream)

// ---- 480-point complex Stockham FFT (forward, e^{-i}), radices 4,4,5,6 ---------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <int R>
__device__ __forceinline__ void butterfly(float2* a) {
    if (R == 2) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else if (R == 4) {
        float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        float2 t2 = cadd(a[1], a[3]), t3 = csub(a[1], a[3]);
        a[0] = cadd(t0, t2);
        a[1] = make_float2(t1.x + t3.y, t1.y - t3.x);
        a[2] = csub(t0, t2);
        a[3] = make_float2(t1.x - t3.y, t1.y + t3.x);
    } else if (R == 3) {
        const float s = 0.86602540378443864676f;
        float2 t1 = cadd(a[1], a[2]), d = csub(a[1], a[2]);
        float2 m1 = make_float2(a[0].x - 0.5f * t1.x, a[0].y - 0.5f * t1.y);
        a[0] = cadd(a[0], t1);
        a[1] = make_float2(m1.x + s * d.y, m1.y - s * d.x);
        a[2] = make_float2(m1.x - s * d.y, m1.y + s * d.x);
    } else if (R == 6) {
        // 6 = 2 x 3: E = DFT3(a0, a2, a4), O = DFT3(a1, a3, a5);  X[k] = E[k] + W6^k O[k],  X[k + 3] = E[k] - W6^k O[k]
        const float h = 0.5f, s = 0.86602540378443864676f;
        float2 e[3] = {a[0], a[2], a[4]}, o[3] = {a[1], a[3], a[5]};
        butterfly<3>(e);
        butterfly<3>(o);
        const float2 t1 = make_float2(h * o[1].x + s * o[1].y, h * o[1].y - s * o[1].x);    // o1 * (1/2, -s)
        const float2 t2 = make_float2(s * o[2].y - h * o[2].x, -h * o[2].y - s * o[2].x);   // o2 * (-1/2, -s)
        a[0] = cadd(e[0], o[0]);
        a[3] = csub(e[0], o[0]);
        a[1] = cadd(e[1], t1);
        a[4] = csub(e[1], t1);
        a[2] = cadd(e[2], t2);
        a[5] = csub(e[2], t2);
    } else if (R == 5) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
        float2 d1 = csub(a[1], a[4]), d2 = csub(a[2], a[3]);
        float2 u1 = make_float2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
        float2 u2 = make_float2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
        float2 v1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
        float2 v2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
        a[0] = cadd(a[0], cadd(t1, t2));
        a[1] = make_float2(u1.x + v1.y, u1.y - v1.x);
        a[4] = make_float2(u1.x - v1.y, u1.y + v1.x);
        a[2] = make_float2(u2.x + v2.y, u2.y - v2.x);
        a[3] = make_float2(u2.x - v2.y, u2.y + v2.x);
    }
}

// One Stockham DIF pass of radix R on sub-length N with stride S = 480 / N (product of previous radices).
template <int R, int N, int S>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ x, float2* __restrict__ y,
                                              const float2* __restrict__ tw) {
    constexpr int M = N / R;
    for (int b = threadIdx.x; b < 480 / R; b += ST) {
        const int p = b / S, q = b - p * S;
        float2 a[R];
#pragma unroll
        for (int k = 0; k < R; k++) a[k] = x[q + S * (p + k * M)];
        butterfly<R>(a);
        y[q + S * (R * p)] = a[0];
#pragma unroll
        for (int j = 1; j < R; j++) {
            // twiddle exp(-2 pi i j p / N) = tw480[j p S];  j p S < 480 because p < N / R
            y[q + S * (R * p + j)] = (N == R) ? a[j] : cmul(a[j], __ldg(&tw[j * p * S]));
        }
    }
    __syncthreads();
}

// forward FFT of a[480]; FOUR passes, so the result lands back in a (b is the pong buffer).  Both buffers in shared
// memory; tw = tw480 table.  (The last two radices 3 and 2 are one radix-6 pass: one barrier and one trip through
// shared memory less.)
__device__ void fft480(float2* a, float2* b, const float2* tw) {
    stockham_pass<4, 480, 1>(a, b, tw);
    stockham_pass<4, 120, 4>(b, a, tw);
    stockham_pass<5, 30, 16>(a, b, tw);
    stockham_pass<6, 6, 80>(b, a, tw);
}

// Two independent 480-point FFTs advanced together (same indices, same twiddles, half the barriers).
template <int R, int N, int S>
__device__ __forceinline__ void stockham_pass2(const float2* __restrict__ x0, float2* __restrict__ y0, const float2* __restrict__ x1,
                                               float2* __restrict__ y1, const float2* __restrict__ tw) {
    constexpr int M = N / R;
    for (int b = threadIdx.x; b < 480 / R; b += ST) {
        const int p = b / S, q = b - p * S;
        float2 a[R], c[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            a[k] = x0[q + S * (p + k * M)];
            c[k] = x1[q + S * (p + k * M)];
        }
        butterfly<R>(a);
        butterfly<R>(c);
        y0[q + S * (R * p)] = a[0];
        y1[q + S * (R * p)] = c[0];
#pragma unroll
        for (int j = 1; j < R; j++) {
            if (N == R) {
                y0[q + S * (R * p + j)] = a[j];
                y1[q + S * (R * p + j)] = c[j];
            } else {
                const float2 w = __ldg(&tw[j * p * S]);
                y0[q + S * (R * p + j)] = cmul(a[j], w);
                y1[q + S * (R * p + j)] = cmul(c[j], w);
            }
        }
    }
    __syncthreads();
}
// forward FFTs of a0[480] and a1[480]; results land back in a0 and a1.
__device__ void fft480x2(float2* a0, float2* b0, float2* a1, float2* b1, const float2* tw) {
    stockham_pass2<4, 480, 1>(a0, b0, a1, b1, tw);
    stockham_pass2<4, 120, 4>(b0, a0, b1, a1, tw);
    stockham_pass2<5, 30, 16>(a0, b0, a1, b1, tw);
    stockham_pass2<6, 6, 80>(b0, a0, b1, a1, tw);
}

// Band-weighted sums (src/lib.rs:65-82) driven by DeviceTables::bt_*: the 800 weighted terms are dealt to BT_LANES = 96
// lanes (3 warps), <= 9 CONSECUTIVE bins of one band each (straight-line, predicated); the lanes of a band are
// contiguous, so a segmented shuffle reduction leaves each band's sum (per warp) in its first lane; bands that straddle a
// warp boundary are completed from the per-warp partials in a fixed order (deterministic).
// NS = 3: ex = |X|^2, ep = |P|^2, exp = Re(X conj P) in one sweep (x, p: spectra in shared memory);
// NS = 1: only |X|^2.  part: shared scratch [NS][3][NB_BANDS].  All threads must call; contains two barriers.
template <int NS>
__device__ __forceinline__ void band_sums(const float2* xs, const float2* ps, const DeviceTables* __restrict__ tab, float* part,
                                          float* o0, float* o1, float* o2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    static_assert(BT_LANES == 96, "three warps of band lanes");
    if (tid < BT_LANES) {
        const int t0 = tab->bt_lane_start[tid], n = tab->bt_lane_start[tid + 1] - t0;
        const int bin0 = tab->bt_bin[t0];
        const int band = tab->bt_lane_band[tid];
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            if (i < n) {
                const float w = __ldg(&tab->bt_w[t0 + i]);
                const float2 x = xs[bin0 + i];
                a0 = fmaf(w, x.x * x.x + x.y * x.y, a0);
                if (NS == 3) {
                    const float2 p = ps[bin0 + i];
                    a1 = fmaf(w, p.x * p.x + p.y * p.y, a1);
                    a2 = fmaf(w, x.x * p.x + x.y * p.y, a2);
                }
            }
        }
        // segmented reduction towards the first lane of each band within the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int ob = __shfl_down_sync(0xffffffffu, band, off);
            const float v0 = __shfl_down_sync(0xffffffffu, a0, off);
            const float v1 = NS == 3 ? __shfl_down_sync(0xffffffffu, a1, off) : 0.0f;
            const float v2 = NS == 3 ? __shfl_down_sync(0xffffffffu, a2, off) : 0.0f;
            if (lane + off < 32 && ob == band) {
                a0 += v0;
                if (NS == 3) {
                    a1 += v1;
                    a2 += v2;
                }
            }
        }
        const int prev = __shfl_up_sync(0xffffffffu, band, 1);
        if (lane == 0 || prev != band) {
            part[warp * NB_BANDS + band] = a0;
            if (NS == 3) {
                part[(3 + warp) * NB_BANDS + band] = a1;
                part[(6 + warp) * NB_BANDS + band] = a2;
            }
        }
    }
    __syncthreads();
    if (tid < NS * 32) {
        const int which = tid >> 5, b = tid & 31;
        if (b < NB_BANDS) {
            const int w0 = tab->bt_band_lane[b] >> 5, w1 = (tab->bt_band_lane[b + 1] - 1) >> 5;
            float acc = 0.0f;
            for (int w = w0; w <= w1; w++) acc += part[(which * 3 + w) * NB_BANDS + b];
            if (b == 0 || b == NB_BANDS - 1) acc *= 2.0f;
            float* o = which == 0 ? o0 : (which == 1 ? o1 : o2);
            o[b] = acc;
        }
    }
    __syncthreads();
}

// interp_band_gain for one bin (src/lib.rs:84-97); bins >= 400 get 0.
__device__ __forceinline__ float interp_gain(const float* g, const DeviceTables* __restrict__ tab, int idx) {
    if (idx >= NB_BINS_BANDED) return 0.0f;
    int b = tab->band_of[idx];
    float f = tab->band_frac[idx];
    return (1.0f - f) * g[b] + f * g[b + 1];
}

// Window * history -> FFT input (src/features.rs:281-290).  The 960 samples starting at ring position `start` are
// fetched into registers first (issue_loads) and multiplied/stored later (store), so that the DRAM latency of the
// second (pitch-lagged) window hides behind other work.
struct WindowLoad {
    float4 hv[2], wv[2];   // aligned path: samples 4q..4q+3 of q = tid, tid + 128
    float2 hs[4], ws[4];   // unaligned path: complex element n = tid + 128 it
    bool aligned;
    __device__ __forceinline__ void issue_loads(const float* __restrict__ h, int start, const DeviceTables* __restrict__ tab) {
        aligned = (start & 3) == 0;
        if (aligned) {
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int q = threadIdx.x + it * ST;
                hv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                wv[it] = hv[it];
                if (q < WINDOW_SIZE / 4) {
                    int pos = start + 4 * q;  // a float4 never straddles the ring wrap (start and HIST_CAP are multiples of 4)
                    if (pos >= HIST_CAP) pos -= HIST_CAP;
                    hv[it] = __ldg(reinterpret_cast<const float4*>(h + pos));
                    wv[it] = __ldg(reinterpret_cast<const float4*>(tab->window) + q);
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int n = threadIdx.x + it * ST;
                hs[it] = make_float2(0.f, 0.f);
                ws[it] = hs[it];
                if (n < 480) {
                    int p0 = start + 2 * n;
                    if (p0 >= HIST_CAP) p0 -= HIST_CAP;
                    int p1 = p0 + 1;
                    if (p1 >= HIST_CAP) p1 -= HIST_CAP;
                    hs[it] = make_float2(__ldg(h + p0), __ldg(h + p1));
                    ws[it] = __ldg(reinterpret_cast<const float2*>(tab->window) + n);
                }
            }
        }
    }
    __device__ __forceinline__ void store(float2* a) const {
        if (aligned) {
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int q = threadIdx.x + it * ST;
                if (q < WINDOW_SIZE / 4)
                    reinterpret_cast<float4*>(a)[q] = make_float4(hv[it].x * wv[it].x, hv[it].y * wv[it].y, hv[it].z * wv[it].z, hv[it].w * wv[it].w);
            }
        } else {
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int n = threadIdx.x + it * ST;
                if (n < 480) a[n] = make_float2(hs[it].x * ws[it].x, hs[it].y * ws[it].y);
            }
        }
    }
};

// even/odd split of a 480-point complex FFT b into the 481 bins of the 960-point real FFT, scaled by wnorm
// (src/features.rs:290-295); bins k and 480-k together (tw960[480-k] = -conj(tw960[k])).  No barrier inside.
__device__ __forceinline__ void rfft_post(const float2* b, float2* xs, const DeviceTables* __restrict__ tab) {
    const float wn = tab->wnorm;
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int k = threadIdx.x + it * ST;
        if (k <= 240) {
            const float2 zk = b[k], zc = b[k == 0 ? 0 : 480 - k];
            const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
            const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);
            const float2 tw = __ldg(&tab->tw960[k]);
            // bin k: E + (di, -dr) * tw ;  bin 480-k: conj(E) + (di, dr) * (-tw.x, tw.y)
            const float tx = di * tw.x + dr * tw.y, ty = di * tw.y - dr * tw.x;
            float2 r0 = make_float2((er + tx) * wn, (ei + ty) * wn);
            float2 r1 = make_float2((er - tx) * wn, (ty - ei) * wn);
            if (k == 0) {
                r0.y = 0.0f;
                r1.y = 0.0f;
            }
            xs[k] = r0;
            if (k != 240) xs[480 - k] = r1;
        }
    }
}

// ================================================================================================
// K3: analysis -- X, P, band energies, features (src/features.rs:115-219)
// ================================================================================================
__global__ void __launch_bounds__(ST, 10) analysis_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab, int hbase) {
    __shared__ __align__(16) float2 xa[FREQ_SIZE + 1];  // X: FFT ping buffer, then the 481-bin spectrum
    __shared__ __align__(16) float2 xb[480];            // X: FFT pong buffer
    __shared__ __align__(16) float2 pa[FREQ_SIZE + 1];  // P: the same for the pitch-lagged window
    __shared__ __align__(16) float2 pb[480];
    __shared__ float part[3 * BT_LANES];
    __shared__ float s_ex[NB_BANDS], s_ep[NB_BANDS], s_exp[NB_BANDS];
    __shared__ float s_feat[NB_FEATURES];
    __shared__ float s_ceps[CEPS_MEM][NB_BANDS];
    __shared__ float s_dist[CEPS_MEM][CEPS_MEM];

    const int s = blockIdx.x, tid = threadIdx.x;
    const float* h = bb.hist + (size_t)s * HIST_CAP;
    const int pitch = bb.pitch[s];

    // X = rfft(window * input_mem[768..1728]),  P = rfft(window * input_mem[768-pitch .. 1728-pitch]): all global
    // loads are issued before the first use, then both transforms advance together.
    int start = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE);
    if (start >= HIST_CAP) start -= HIST_CAP;
    WindowLoad wx, wp;
    wx.issue_loads(h, start, tab);
    start = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE) - pitch;  // >= 0 since pitch <= 768
    if (start >= HIST_CAP) start -= HIST_CAP;
    wp.issue_loads(h, start, tab);
    wx.store(xa);
    wp.store(pa);
    __syncthreads();
    fft480x2(xa, xb, pa, pb, tab->tw480);
    float2* xs = xa;
    float2* ps = pa;
    rfft_post(xa, xs, tab);  // in place: bins k and 480 - k are read and written by the same thread
    rfft_post(pa, ps, tab);
    __syncthreads();

    float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
    float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
    for (int k = tid; k <= 480; k += ST) Xg[k] = xs[k];
    for (int k = tid; k < NB_BINS_BANDED; k += ST) Pg[k] = ps[k];
    band_sums<3>(xs, ps, tab, part, s_ex, s_ep, s_exp);

    // ---- features (src/features.rs:134-219): 22 bands = one warp, lane = band; no block barriers from here on ----
    if (tid >= 32) return;
    const int lane = tid;
    const bool bl = lane < NB_BANDS;
    const float ex = bl ? s_ex[lane] : 0.0f, ep = bl ? s_ep[lane] : 0.0f;
    const float xpn = bl ? s_exp[lane] / sqrtf(0.001f + ex * ep) : 0.0f;
    if (bl) {
        bb.ex[(size_t)s * NB_BANDS + lane] = ex;
        bb.ep[(size_t)s * NB_BANDS + lane] = ep;
        bb.exp[(size_t)s * NB_BANDS + lane] = xpn;
    }
    // cepstral ring (8 x 22): 6 elements per lane, in flight while the log energies are computed
    float* cg = bb.ceps_mem + (size_t)s * CEPS_MEM * NB_BANDS;
    const int mem_id = bb.ceps_id[s];
    float cr[6];
#pragma unroll
    for (int k = 0; k < 6; k++) cr[k] = (lane + 32 * k < CEPS_MEM * NB_BANDS) ? cg[lane + 32 * k] : 0.0f;
    // log band energies with the sequential follower (src/features.rs:147-158) and the silence test (:160)
    const float lg = bl ? log10f(1e-2f + ex) : 0.0f;
    float ly = 0.0f, log_max = -2.0f, follow = -2.0f, e = 0.0f;
#pragma unroll
    for (int k = 0; k < NB_BANDS; k++) {
        const float v = fmaxf(fmaxf(__shfl_sync(0xffffffffu, lg, k), log_max - 7.0f), follow - 1.5f);
        if (lane == k) ly = v;
        log_max = fmaxf(log_max, v);
        follow = fmaxf(follow - 1.5f, v);
        e += __shfl_sync(0xffffffffu, ex, k);
    }
    float* featg = bb.features + (size_t)s * NB_FEATURES;
    if (e < 0.04f) {  // silent frame: zero features, cepstral ring untouched (src/features.rs:160-166)
        featg[lane] = 0.0f;
        if (lane + 32 < NB_FEATURES) featg[lane + 32] = 0.0f;
        if (lane == 0) bb.silence[s] = 1;
        return;
    }
    // both DCTs (src/lib.rs:139-148) share the table: lane i accumulates output i over j in order
    const double dct_scale = 0.30151134457776362265;  // sqrt(2/22), src/lib.rs:146
    float sum_ly = 0.0f, sum_xp = 0.0f;
#pragma unroll
    for (int j = 0; j < NB_BANDS; j++) {
        const float d = bl ? __ldg(&tab->dct[j * NB_BANDS + lane]) : 0.0f;
        sum_ly += __shfl_sync(0xffffffffu, ly, j) * d;
        sum_xp += __shfl_sync(0xffffffffu, xpn, j) * d;
    }
    float ceps = (float)((double)sum_ly * dct_scale);
    float pcor = (float)((double)sum_xp * dct_scale);
    if (lane == 0) {
        ceps -= 12.0f;
        pcor -= 1.3f;
    }
    if (lane == 1) {
        ceps -= 4.0f;
        pcor -= 0.9f;
    }
    // ring -> shared memory, with the new row in place
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (lane + 32 * k < CEPS_MEM * NB_BANDS) (&s_ceps[0][0])[lane + 32 * k] = cr[k];
    __syncwarp();
    if (bl) {
        s_ceps[mem_id][lane] = ceps;
        cg[mem_id * NB_BANDS + lane] = ceps;
        s_feat[lane] = ceps;
    }
    __syncwarp();
    if (lane < NB_DELTA_CEPS) {
        const int c1 = (mem_id < 1) ? CEPS_MEM + mem_id - 1 : mem_id - 1;
        const int c2 = (mem_id < 2) ? CEPS_MEM + mem_id - 2 : mem_id - 2;
        const float a = s_ceps[mem_id][lane], b = s_ceps[c1][lane], c = s_ceps[c2][lane];
        s_feat[lane] = a + b + c;
        s_feat[NB_BANDS + lane] = a - c;
        s_feat[NB_BANDS + NB_DELTA_CEPS + lane] = a - 2.0f * b + c;
        s_feat[NB_BANDS + 2 * NB_DELTA_CEPS + lane] = pcor;
    }
    // spectral variability (src/features.rs:199-216): pairwise squared distances of the 8 ring rows, two pairs per lane
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
        const int pr = lane + 32 * h2, i = pr >> 3, j = pr & 7;
        float dist = 0.0f;
#pragma unroll
        for (int k = 0; k < NB_BANDS; k++) {
            const float t = s_ceps[i][k] - s_ceps[j][k];
            dist += t * t;
        }
        s_dist[i][j] = dist;
    }
    __syncwarp();
    float md = 1e15f;
    if (lane < CEPS_MEM) {
#pragma unroll
        for (int j = 0; j < CEPS_MEM; j++)
            if (j != lane) md = fminf(md, s_dist[lane][j]);
    }
    float sv = 0.0f;
#pragma unroll
    for (int i = 0; i < CEPS_MEM; i++) sv += __shfl_sync(0xffffffffu, md, i);  // i = 0..7 in order, like the reference
    if (lane == 0) {
        s_feat[NB_BANDS + 3 * NB_DELTA_CEPS] = 0.01f * ((float)pitch - 300.0f);
        s_feat[NB_BANDS + 3 * NB_DELTA_CEPS + 1] = sv / (float)CEPS_MEM - 2.1f;
        bb.ceps_id[s] = (mem_id + 1 == CEPS_MEM) ? 0 : mem_id + 1;
        bb.silence[s] = 0;
    }
    __syncwarp();
    featg[lane] = s_feat[lane];
    if (lane + 32 < NB_FEATURES) featg[lane + 32] = s_feat[lane + 32];
}

cudaError_t launch_analysis(const BatchBuffers& b, const DeviceTables* tab, int slot, cudaStream_t st) {
    analysis_kernel<<<b.n_streams, ST, 0, st>>>(b, tab, hist_base(slot));
    return cudaGetLastError();
}

// ================================================================================================
// K5: synthesis -- pitch filter, gain floor, band-gain interpolation, inverse FFT, overlap-add
// (src/denoise.rs:102-115, src/features.rs:223-275)
// ================================================================================================
// TOut = float, or short: clamp to the int16 range then round half away from zero (what both reference front-ends do:
// src/nnnoiseless.rs:152 `clamp().round() as i16`, test_data/rnnoise_demo.c:53 roundf).
__device__ __forceinline__ short to_pcm16(float v) { return (short)roundf(fminf(fmaxf(v, -32768.0f), 32767.0f)); }

template <typename TOut>
__global__ void __launch_bounds__(ST) synthesis_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab,
                                                       TOut* __restrict__ out, long stream_stride, long sample_stride,
                                                       float* __restrict__ vad_out) {
    __shared__ __align__(16) float2 xs[FREQ_SIZE + 1];
    __shared__ __align__(16) float2 fa_[480];
    __shared__ __align__(16) float2 fb_[480];
    __shared__ float part[BT_LANES];
    __shared__ float s_g[NB_BANDS], s_r[NB_BANDS], s_ne[NB_BANDS], s_ex[NB_BANDS];

    const int s = blockIdx.x, tid = threadIdx.x;
    // every global load of the block is issued up front (one DRAM round trip instead of a chain of dependent ones)
    const float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
    const float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
    float2 xv[4], pv[4];
    int bidx[4];
    float bfr[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const int k = tid + it * ST;
        xv[it] = k <= 480 ? __ldg(Xg + k) : make_float2(0.f, 0.f);
        pv[it] = k < NB_BINS_BANDED ? __ldg(Pg + k) : make_float2(0.f, 0.f);
        bidx[it] = k < NB_BINS_BANDED ? __ldg(&tab->band_of[k]) : 0;
        bfr[it] = k < NB_BINS_BANDED ? __ldg(&tab->band_frac[k]) : 0.0f;
    }
    float b_e = 0.f, b_g = 0.f, b_ex = 0.f, b_ep = 0.f, b_lg = 0.f;
    if (tid < NB_BANDS) {
        b_e = bb.exp[(size_t)s * NB_BANDS + tid];
        b_g = bb.gains[(size_t)s * NB_BANDS + tid];
        b_ex = bb.ex[(size_t)s * NB_BANDS + tid];
        b_ep = bb.ep[(size_t)s * NB_BANDS + tid];
        b_lg = bb.lastg[(size_t)s * NB_BANDS + tid];
    }
    const float vad_in = bb.vad[s];
    const int silent = bb.silence[s];
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const int k = tid + it * ST;
        if (k <= 480) xs[k] = xv[it];
    }

    if (!silent) {
        if (tid < NB_BANDS) {
            // r (src/features.rs:226-235)
            const float e = b_e, g = b_g, ex = b_ex, ep = b_ep;
            float r;
            if (e > g) {
                r = 1.0f;
            } else {
                float e2 = e * e, g2 = g * g;
                r = e2 * (1.0f - g2) / (0.001f + g2 * (1.0f - e2));
            }
            r = (r < 0.0f) ? 0.0f : r;
            r = (r > 1.0f) ? 1.0f : r;
            r = sqrtf(r);
            r *= sqrtf(ex / (1e-8f + ep));
            s_r[tid] = r;
            s_ex[tid] = ex;
            // gain floor (src/denoise.rs:106-109)
            float gg = fmaxf(g, 0.6f * b_lg);
            s_g[tid] = gg;
            bb.lastg[(size_t)s * NB_BANDS + tid] = gg;
        }
        __syncthreads();
        // x += rf * p  (bin 0 is the real-valued DC offset; its imaginary part stays 0)
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int k = tid + it * ST;
            if (k < NB_BINS_BANDED) {
                const float rf = (1.0f - bfr[it]) * s_r[bidx[it]] + bfr[it] * s_r[bidx[it] + 1];
                float2 x = xs[k];
                x.x += pv[it].x * rf;
                if (k > 0) x.y += pv[it].y * rf;
                xs[k] = x;
            }
        }
        __syncthreads();
        band_sums<1>(xs, xs, tab, part, s_ne, s_ne, s_ne);
        if (tid < NB_BANDS) s_r[tid] = sqrtf(s_ex[tid] / (1e-8f + s_ne[tid]));
        __syncthreads();
        // x *= rf2 ; x *= gf   (bins >= 400 are zeroed by both interpolations)
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int k = tid + it * ST;
            if (k <= 480) {
                float2 x = xs[k];
                if (k < NB_BINS_BANDED) {
                    const float m1 = (1.0f - bfr[it]) * s_r[bidx[it]] + bfr[it] * s_r[bidx[it] + 1];
                    const float m2 = (1.0f - bfr[it]) * s_g[bidx[it]] + bfr[it] * s_g[bidx[it] + 1];
                    x.x = (x.x * m1) * m2;
                    x.y = (x.y * m1) * m2;
                } else {
                    x = make_float2(0.f, 0.f);
                }
                xs[k] = x;
            }
        }
    }
    __syncthreads();

    // ---- inverse real FFT (unnormalised), src/features.rs:263-275: Z = 2E + i 2O, fed conjugated to the forward FFT.
    // Z[k] and Z[480-k] are built together (tw960[480-k] = -conj tw960[k]).
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int k = tid + it * ST;
        if (k <= 240) {
            const float2 xk = xs[k], xc = xs[480 - k];
            const float xi = (k == 0) ? 0.0f : xk.y, yi = (k == 0) ? 0.0f : xc.y;  // DC / Nyquist imaginary parts are ignored
            const float sr = xk.x + xc.x, si = xi - yi;
            const float dr = xk.x - xc.x, di = xi + yi;
            const float2 w = __ldg(&tab->tw960[k]);
            // t = (dr, di) * conj(w)
            const float tx = dr * w.x + di * w.y, ty = di * w.x - dr * w.y;
            fa_[k] = make_float2(sr - ty, -(si + tx));           // conj(Z[k])
            if (k != 0 && k != 240) {
                // t' = (-dr, di) * (-w) = (dr w.x + di w.y, dr w.y - di w.x) ; Z[480-k] = (sr - t'.y, -si + t'.x)
                const float ux = dr * w.x + di * w.y, uy = dr * w.y - di * w.x;
                fa_[480 - k] = make_float2(sr - uy, -(-si + ux));  // conj(Z[480-k])
            }
        }
    }
    __syncthreads();
    fft480(fa_, fb_, tab->tw480);
    float* sm = bb.synth_mem + (size_t)s * FRAME_SIZE;
    TOut* o = out + (long)s * stream_stride;
    // time samples 4q..4q+3 = (re, -im) of fa_[2q], fa_[2q+1] (the four-pass FFT ends in its input buffer); first half -> output (+ overlap memory),
    // second half -> new overlap memory.  Vector stores when the caller's rows are aligned for them.
    const bool o_vec = sample_stride == 1 && ((reinterpret_cast<uintptr_t>(o) & (4 * sizeof(TOut) - 1)) == 0);
    const long ss = sample_stride;
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int q = tid + it * ST;
        if (q < WINDOW_SIZE / 4) {
            const float4 z = reinterpret_cast<const float4*>(fa_)[q];
            const float4 w = __ldg(reinterpret_cast<const float4*>(tab->window) + q);
            const float4 v = make_float4((z.x * 0.5f) * w.x, (-z.y * 0.5f) * w.y, (z.z * 0.5f) * w.z, (-z.w * 0.5f) * w.w);
            if (q < FRAME_SIZE / 4) {
                const float4 m = reinterpret_cast<const float4*>(sm)[q];
                const float4 r = make_float4(v.x + m.x, v.y + m.y, v.z + m.z, v.w + m.w);
                if (sizeof(TOut) == 4) {
                    float* of = reinterpret_cast<float*>(o);
                    if (o_vec) {
                        reinterpret_cast<float4*>(of)[q] = r;
                    } else {
                        of[(4 * q) * ss] = r.x; of[(4 * q + 1) * ss] = r.y; of[(4 * q + 2) * ss] = r.z; of[(4 * q + 3) * ss] = r.w;
                    }
                } else {
                    short* os = reinterpret_cast<short*>(o);
                    const short p0 = to_pcm16(r.x), p1 = to_pcm16(r.y), p2 = to_pcm16(r.z), p3 = to_pcm16(r.w);
                    if (o_vec) {
                        reinterpret_cast<uint2*>(os)[q] = make_uint2((unsigned)(unsigned short)p0 | ((unsigned)(unsigned short)p1 << 16),
                                                                     (unsigned)(unsigned short)p2 | ((unsigned)(unsigned short)p3 << 16));
                    } else {
                        os[(4 * q) * ss] = p0; os[(4 * q + 1) * ss] = p1; os[(4 * q + 2) * ss] = p2; os[(4 * q + 3) * ss] = p3;
                    }
                }
            } else {
                reinterpret_cast<float4*>(fb_)[q - FRAME_SIZE / 4] = v;  // staged: sm is still being read by other threads
            }
        }
    }
    __syncthreads();
    for (int q = tid; q < FRAME_SIZE / 4; q += ST) reinterpret_cast<float4*>(sm)[q] = reinterpret_cast<const float4*>(fb_)[q];
    if (tid == 0 && vad_out) vad_out[s] = silent ? 0.0f : vad_in;
}

cudaError_t launch_synthesis(const BatchBuffers& b, const DeviceTables* tab, void* out, bool pcm16, long stream_stride, long sample_stride,
                             float* vad_out, cudaStream_t st) {
    if (pcm16) synthesis_kernel<short><<<b.n_streams, ST, 0, st>>>(b, tab, static_cast<short*>(out), stream_stride, sample_stride, vad_out);
    else synthesis_kernel<float><<<b.n_streams, ST, 0, st>>>(b, tab, static_cast<float*>(out), stream_stride, sample_stride, vad_out);
    return cudaGetLastError();
}

}  // namespace nnb
