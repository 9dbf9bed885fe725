// spectral.cu -- frequency-domain half of the path: windowed 960-point real FFTs of the current
// and pitch-lagged windows, Bark-band energies/correlations, the 42 input features, and (after the
// RNN) pitch filtering, band-gain interpolation, inverse FFT and overlap-add.
//
// Reference: src/features.rs:115-298, src/lib.rs:65-162, src/denoise.rs:95-116.
// f32 with FMA contraction allowed: these stages are compared to the oracle within tolerance.
#include "common.cuh"

namespace nnb {

constexpr int ST = 128;  // threads per block (one block per stream)

// ---- 480-point complex Stockham FFT (forward, e^{-i}), radices 4,4,5,3,2 ---------------------------
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

// forward FFT of a[480]; result lands in b.  Both buffers in shared memory; tw = tw480 table.
__device__ void fft480(float2* a, float2* b, const float2* tw) {
    stockham_pass<4, 480, 1>(a, b, tw);
    stockham_pass<4, 120, 4>(b, a, tw);
    stockham_pass<5, 30, 16>(a, b, tw);
    stockham_pass<3, 6, 80>(b, a, tw);
    stockham_pass<2, 2, 240>(a, b, tw);
}

// Band-weighted correlation (src/lib.rs:65-82).  cb[0..400) holds Re(x conj p) per bin.
// Band t = frac-weighted part of segment t-1 plus (1-frac)-weighted part of segment t.  Four lanes per
// band (threads 0..87), partial sums combined with shuffles; call with all threads of warps 0-2.
__device__ void band_sums(const float* cb, const DeviceTables* __restrict__ tab, float* out) {
    const int tid = threadIdx.x;
    if (tid < 96) {
        const int t = tid >> 2, part = tid & 3;
        float acc = 0.0f;
        if (t < NB_BANDS) {
            if (t > 0) {
                const int lo = tab->band_start[t - 1], hi = tab->band_start[t];
                for (int i = lo + part; i < hi; i += 4) acc += __ldg(&tab->band_frac[i]) * cb[i];
            }
            if (t < NB_BANDS - 1) {
                const int lo = tab->band_start[t], hi = tab->band_start[t + 1];
                for (int i = lo + part; i < hi; i += 4) acc += (1.0f - __ldg(&tab->band_frac[i])) * cb[i];
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (t < NB_BANDS && part == 0) out[t] = (t == 0 || t == NB_BANDS - 1) ? 2.0f * acc : acc;
    }
}

// interp_band_gain for one bin (src/lib.rs:84-97); bins >= 400 get 0.
__device__ __forceinline__ float interp_gain(const float* g, const DeviceTables* __restrict__ tab, int idx) {
    if (idx >= NB_BINS_BANDED) return 0.0f;
    int b = tab->band_of[idx];
    float f = tab->band_frac[idx];
    return (1.0f - f) * g[b] + f * g[b + 1];
}

// Windowed real FFT of hist[(start + i)], i < 960 (ring-indexed); writes X[0..480] (scaled by wnorm)
// into xs (shared).  a/b: scratch FFT buffers.  src/features.rs:281-298.
__device__ void windowed_rfft(const float* __restrict__ h, int start, const DeviceTables* __restrict__ tab, float2* a,
                              float2* b, float2* xs) {
    for (int n = threadIdx.x; n < 480; n += ST) {
        int p0 = start + 2 * n;
        if (p0 >= HIST_CAP) p0 -= HIST_CAP;
        int p1 = p0 + 1;
        if (p1 >= HIST_CAP) p1 -= HIST_CAP;
        a[n] = make_float2(h[p0] * tab->window[2 * n], h[p1] * tab->window[2 * n + 1]);
    }
    __syncthreads();
    fft480(a, b, tab->tw480);
    const float wn = tab->wnorm;
    for (int k = threadIdx.x; k <= 480; k += ST) {
        float2 zk = b[k == 480 ? 0 : k];
        float2 zc = b[k == 0 ? 0 : 480 - k];
        float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
        float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);
        float2 t = cmul(make_float2(di, -dr), tab->tw960[k]);
        float2 r = make_float2((er + t.x) * wn, (ei + t.y) * wn);
        if (k == 0 || k == 480) r.y = 0.0f;
        xs[k] = r;
    }
    __syncthreads();
}

// ================================================================================================
// K3: analysis -- X, P, band energies, features (src/features.rs:115-219)
// ================================================================================================
__global__ void __launch_bounds__(ST) analysis_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab, int hbase) {
    __shared__ float2 fa_[FREQ_SIZE + 1];  // FFT ping buffer; also receives the 481-bin P spectrum
    __shared__ float2 fb_[480];
    __shared__ float2 xs[FREQ_SIZE + 1];
    __shared__ float cb[NB_BINS_BANDED];
    __shared__ float s_ex[NB_BANDS], s_ep[NB_BANDS], s_exp[NB_BANDS], s_tmp[NB_BANDS], s_ly[NB_BANDS];
    __shared__ float s_feat[NB_FEATURES];
    __shared__ float s_ceps[CEPS_MEM][NB_BANDS];
    __shared__ float s_dist[CEPS_MEM][CEPS_MEM];
    __shared__ int s_flag;

    const int s = blockIdx.x, tid = threadIdx.x;
    const float* h = bb.hist + (size_t)s * HIST_CAP;
    const int pitch = bb.pitch[s];

    // ---- X = rfft(window * input_mem[768..1728]) ----
    int start = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE);
    if (start >= HIST_CAP) start -= HIST_CAP;
    windowed_rfft(h, start, tab, fa_, fb_, xs);
    float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
    for (int k = tid; k <= 480; k += ST) Xg[k] = xs[k];
    for (int k = tid; k < NB_BINS_BANDED; k += ST) cb[k] = xs[k].x * xs[k].x + xs[k].y * xs[k].y;
    __syncthreads();
    band_sums(cb, tab, s_ex);
    __syncthreads();

    // ---- P = rfft(window * input_mem[768-pitch .. 1728-pitch]) ----
    start = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE) - pitch;  // >= 0 since pitch <= 768
    if (start >= HIST_CAP) start -= HIST_CAP;
    float2* ps = fa_;  // P spectrum ends up in fa_ (free once the FFT result sits in fb_)
    windowed_rfft(h, start, tab, fa_, fb_, ps);
    float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
    for (int k = tid; k < NB_BINS_BANDED; k += ST) {
        Pg[k] = ps[k];
        cb[k] = ps[k].x * ps[k].x + ps[k].y * ps[k].y;
    }
    __syncthreads();
    band_sums(cb, tab, s_ep);
    __syncthreads();
    for (int k = tid; k < NB_BINS_BANDED; k += ST) cb[k] = xs[k].x * ps[k].x + xs[k].y * ps[k].y;
    __syncthreads();
    band_sums(cb, tab, s_exp);
    __syncthreads();

    // ---- features ----
    if (tid < NB_BANDS) {
        s_exp[tid] = s_exp[tid] / sqrtf(0.001f + s_ex[tid] * s_ep[tid]);
        bb.ex[(size_t)s * NB_BANDS + tid] = s_ex[tid];
        bb.ep[(size_t)s * NB_BANDS + tid] = s_ep[tid];
        bb.exp[(size_t)s * NB_BANDS + tid] = s_exp[tid];
    }
    __syncthreads();
    const double dct_scale = 0.30151134457776362265;  // sqrt(2/22), src/lib.rs:146
    if (tid < NB_BANDS) {
        float sum = 0.0f;
        for (int j = 0; j < NB_BANDS; j++) sum += s_exp[j] * tab->dct[j * NB_BANDS + tid];
        s_tmp[tid] = (float)((double)sum * dct_scale);
    }
    if (tid >= 32 && tid < 64) {  // another warp: log band energies, sequential follower (src/features.rs:147-158)
        const int i = tid - 32;
        float lg = (i < NB_BANDS) ? log10f(1e-2f + s_ex[i]) : 0.0f;
        float exi = (i < NB_BANDS) ? s_ex[i] : 0.0f;
        float log_max = -2.0f, follow = -2.0f, e = 0.0f;
#pragma unroll
        for (int k = 0; k < NB_BANDS; k++) {
            float ly = fmaxf(fmaxf(__shfl_sync(0xffffffffu, lg, k), log_max - 7.0f), follow - 1.5f);
            if (i == k) s_ly[k] = ly;
            log_max = fmaxf(log_max, ly);
            follow = fmaxf(follow - 1.5f, ly);
            e += __shfl_sync(0xffffffffu, exi, k);
        }
        if (i == 0) s_flag = (e < 0.04f) ? 1 : 0;
    }
    __syncthreads();
    const int silent = s_flag;
    float* featg = bb.features + (size_t)s * NB_FEATURES;
    if (silent) {
        if (tid < NB_FEATURES) featg[tid] = 0.0f;
        if (tid == 0) bb.silence[s] = 1;
        return;
    }
    // ceps ring
    float* cg = bb.ceps_mem + (size_t)s * CEPS_MEM * NB_BANDS;
    const int mem_id = bb.ceps_id[s];
    for (int i = tid; i < CEPS_MEM * NB_BANDS; i += ST) (&s_ceps[0][0])[i] = cg[i];
    if (tid < NB_BANDS) {
        float sum = 0.0f;
        for (int j = 0; j < NB_BANDS; j++) sum += s_ly[j] * tab->dct[j * NB_BANDS + tid];
        float v = (float)((double)sum * dct_scale);
        if (tid == 0) v -= 12.0f;
        if (tid == 1) v -= 4.0f;
        s_feat[tid] = v;
    }
    __syncthreads();
    if (tid < NB_BANDS) {
        s_ceps[mem_id][tid] = s_feat[tid];
        cg[mem_id * NB_BANDS + tid] = s_feat[tid];
    }
    __syncthreads();
    if (tid < NB_DELTA_CEPS) {
        const int c1 = (mem_id < 1) ? CEPS_MEM + mem_id - 1 : mem_id - 1;
        const int c2 = (mem_id < 2) ? CEPS_MEM + mem_id - 2 : mem_id - 2;
        float a = s_ceps[mem_id][tid], b = s_ceps[c1][tid], c = s_ceps[c2][tid];
        s_feat[tid] = a + b + c;
        s_feat[NB_BANDS + tid] = a - c;
        s_feat[NB_BANDS + NB_DELTA_CEPS + tid] = a - 2.0f * b + c;
        float v = s_tmp[tid];
        if (tid == 0) v -= 1.3f;
        if (tid == 1) v -= 0.9f;
        s_feat[NB_BANDS + 2 * NB_DELTA_CEPS + tid] = v;
    }
    if (tid >= 64 && tid < 128) {
        int i = (tid - 64) >> 3, j = (tid - 64) & 7;
        float dist = 0.0f;
        for (int k = 0; k < NB_BANDS; k++) {
            float t = s_ceps[i][k] - s_ceps[j][k];
            dist += t * t;
        }
        s_dist[i][j] = dist;
    }
    __syncthreads();
    if (tid == 0) {
        float sv = 0.0f;
        for (int i = 0; i < CEPS_MEM; i++) {
            float md = 1e15f;
            for (int j = 0; j < CEPS_MEM; j++)
                if (j != i) md = fminf(md, s_dist[i][j]);
            sv += md;
        }
        s_feat[NB_BANDS + 3 * NB_DELTA_CEPS] = 0.01f * ((float)pitch - 300.0f);
        s_feat[NB_BANDS + 3 * NB_DELTA_CEPS + 1] = sv / (float)CEPS_MEM - 2.1f;
        bb.ceps_id[s] = (mem_id + 1 == CEPS_MEM) ? 0 : mem_id + 1;
        bb.silence[s] = 0;
    }
    __syncthreads();
    if (tid < NB_FEATURES) featg[tid] = s_feat[tid];
}

cudaError_t launch_analysis(const BatchBuffers& b, const DeviceTables* tab, int slot, cudaStream_t st) {
    analysis_kernel<<<b.n_streams, ST, 0, st>>>(b, tab, hist_base(slot));
    return cudaGetLastError();
}

// ================================================================================================
// K5: synthesis -- pitch filter, gain floor, band-gain interpolation, inverse FFT, overlap-add
// (src/denoise.rs:102-115, src/features.rs:223-275)
// ================================================================================================
__global__ void __launch_bounds__(ST) synthesis_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab,
                                                       float* __restrict__ out, long stream_stride,
                                                       float* __restrict__ vad_out) {
    __shared__ float2 xs[FREQ_SIZE + 1];
    __shared__ float2 fa_[480];
    __shared__ float2 fb_[480];
    __shared__ float cb[NB_BINS_BANDED];
    __shared__ float s_g[NB_BANDS], s_r[NB_BANDS], s_ne[NB_BANDS], s_ex[NB_BANDS];

    const int s = blockIdx.x, tid = threadIdx.x;
    const int silent = bb.silence[s];
    const float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
    for (int k = tid; k <= 480; k += ST) xs[k] = Xg[k];

    if (!silent) {
        const float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
        if (tid < NB_BANDS) {
            // r (src/features.rs:226-235)
            float e = bb.exp[(size_t)s * NB_BANDS + tid], g = bb.gains[(size_t)s * NB_BANDS + tid];
            float ex = bb.ex[(size_t)s * NB_BANDS + tid], ep = bb.ep[(size_t)s * NB_BANDS + tid];
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
            float lg = bb.lastg[(size_t)s * NB_BANDS + tid];
            float gg = fmaxf(g, 0.6f * lg);
            s_g[tid] = gg;
            bb.lastg[(size_t)s * NB_BANDS + tid] = gg;
        }
        __syncthreads();
        // x += rf * p ; then new band energies
        for (int k = tid; k < NB_BINS_BANDED; k += ST) {
            float rf = interp_gain(s_r, tab, k);
            float2 p = Pg[k];
            float2 x = xs[k];
            x.x += p.x * rf;
            if (k > 0) x.y += p.y * rf;  // bin 0 is the real-valued DC offset
            xs[k] = x;
            cb[k] = x.x * x.x + x.y * x.y;
        }
        __syncthreads();
        band_sums(cb, tab, s_ne);
        __syncthreads();
        if (tid < NB_BANDS) s_r[tid] = sqrtf(s_ex[tid] / (1e-8f + s_ne[tid]));
        __syncthreads();
        // x *= rf2 ; x *= gf   (bins >= 400 are zeroed by both interpolations)
        for (int k = tid; k <= 480; k += ST) {
            float m1 = interp_gain(s_r, tab, k);
            float m2 = interp_gain(s_g, tab, k);
            float2 x = xs[k];
            x.x = (x.x * m1) * m2;
            x.y = (x.y * m1) * m2;
            xs[k] = x;
        }
    }
    __syncthreads();

    // ---- inverse real FFT (unnormalised), src/features.rs:263-275 ----
    for (int k = tid; k < 480; k += ST) {
        float2 xk = xs[k], xc = xs[480 - k];
        float xi = (k == 0) ? 0.0f : xk.y;
        float yi = (k == 0) ? 0.0f : -xc.y;
        float sr = xk.x + xc.x, si = xi + yi;
        float dr = xk.x - xc.x, di = xi - yi;
        float2 w = tab->tw960[k];
        float2 t = cmul(make_float2(dr, di), make_float2(w.x, -w.y));
        fa_[k] = make_float2(sr - t.y, -(si + t.x));  // conj(Z)
    }
    __syncthreads();
    fft480(fa_, fb_, tab->tw480);
    float* sm = bb.synth_mem + (size_t)s * FRAME_SIZE;
    float* o = out + (long)s * stream_stride;
    for (int n = tid; n < 480; n += ST) {
        // time samples 2n, 2n+1 = (re, -im) of fb_[n]
        float2 z = fb_[n];
        float v0 = (z.x * 0.5f) * tab->window[2 * n];
        float v1 = (-z.y * 0.5f) * tab->window[2 * n + 1];
        // first half -> output (+ overlap memory); second half -> new overlap memory
        if (n < 240) {
            o[2 * n] = v0 + sm[2 * n];
            o[2 * n + 1] = v1 + sm[2 * n + 1];
        }
        fa_[n] = make_float2(v0, v1);
    }
    __syncthreads();
    for (int n = tid; n < 240; n += ST) {
        float2 v = fa_[240 + n];
        sm[2 * n] = v.x;
        sm[2 * n + 1] = v.y;
    }
    if (tid == 0 && vad_out) vad_out[s] = silent ? 0.0f : bb.vad[s];
}

cudaError_t launch_synthesis(const BatchBuffers& b, const DeviceTables* tab, float* out, long stream_stride, float* vad_out,
                             cudaStream_t st) {
    synthesis_kernel<<<b.n_streams, ST, 0, st>>>(b, tab, out, stream_stride, vad_out);
    return cudaGetLastError();
}

}  // namespace nnb
