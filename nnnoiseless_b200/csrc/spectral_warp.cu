// spectral_warp.cu -- frequency-domain half of the path with ONE WARP PER STREAM and no block barrier:
// windowed 960-point real FFTs of the current and pitch-lagged windows, Bark-band energies / correlations, the 42
// input features (analysis), and after the RNN pitch filtering, band-gain interpolation, inverse FFT, overlap-add
// (synthesis).  Reference: src/features.rs:115-298, src/lib.rs:65-162, src/denoise.rs:95-116.
// f32 with FMA contraction allowed: these stages are compared to the oracle within tolerance.
//
// Mapping (fft480.cuh): the 480-point complex FFT behind each 960-point real transform is 32 lanes x 15 registers --
// a 15-point DFT per lane, a twiddle, a transpose through 7.9 KB of per-warp shared memory, a 32-point FFT per lane.
// The two forward transforms of a frame (X and the pitch-lagged P) share every step: lanes 0-14 finish X while lanes
// 15-29 finish P.  Everything between the transforms works on the warp's own buffer under __syncwarp: the even/odd
// split in place (a lane owns bins k and 480-k), the band sums (every lane stays inside one band segment, <= 22 bins),
// the 22-band feature tail (lane = band).  Round 1 spent a 128-thread block per stream with 12 block barriers.
#include "common.cuh"
#include "fft480.cuh"

namespace nnb {

namespace {

constexpr int WPB = 4;             // warps = streams per block
constexpr int RS = 33;             // transpose row stride in float2: lane L reads row L, 2 * 33 = 2 (mod 32): conflict-free
constexpr int ZP_OFF = 495;        // second spectrum inside the warp buffer; 2 * 495 = 30 (mod 32): X and P lanes interleave
constexpr int WBUF = 30 * RS;      // 990 float2 per warp
constexpr int SC_SR = 0, SC_SG = 32, SC_SN = 64, SC_BAND = 96;  // float offsets in the small per-warp scratch
constexpr int WSC = SC_BAND + 21 * 6 + 2;
static_assert(ZP_OFF + FREQ_SIZE <= WBUF, "two 481-bin spectra must fit the warp buffer");

// A warp's first act is a burst of loads that miss to HBM, and with 20-24 warps per SM that latency is only partly
// covered.  Blocks are dispatched in index order, so the stream that will start when this warp's block retires is about
// one resident wave ahead: each warp asks L2 for that stream's rows (prefetch.global.L2, no register, no dependency)
// right after issuing its own loads, turning the next wave's HBM misses into L2 hits.
// (Measured: analysis 0.359 -> 0.325 ms.  The same request in synthesis and in the high-pass kernel made them 2-7 %
// slower -- their inputs were written one or two kernels earlier and largely still sit in L2.)
constexpr int PF_WAVE_A = 148 * 6 * WPB;  // analysis: 6 blocks per SM
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// forward 480-point FFTs of vx (and vp if TWO) held as z[32 a + lane] -> natural-order spectra in buf[0..480) (and
// buf[ZP_OFF..ZP_OFF+480)).  All 32 lanes must call.
template <bool TWO>
__device__ __forceinline__ void fft480_warp(float2 (&vx)[15], float2 (&vp)[15], float2* __restrict__ buf,
                                            const DeviceTables* __restrict__ tab, int lane) {
    dft15(vx);
    if (TWO) dft15(vp);
#pragma unroll
    for (int k1 = 1; k1 < 15; k1++) {
        const float2 tw = __ldg(&tab->twl[k1][lane]);
        vx[k1] = c_mul(vx[k1], tw);
        if (TWO) vp[k1] = c_mul(vp[k1], tw);
    }
#pragma unroll
    for (int k1 = 0; k1 < 15; k1++) {
        buf[k1 * RS + lane] = vx[k1];
        if (TWO) buf[(15 + k1) * RS + lane] = vp[k1];
    }
    __syncwarp();
    float2 r[32];
    const bool on = lane < (TWO ? 30 : 15);
    if (on) {
#pragma unroll
        for (int b = 0; b < 32; b++) r[b] = buf[lane * RS + b];
        fft32_dif(r);
    }
    __syncwarp();
    if (on) {
        const int base = lane < 15 ? lane : ZP_OFF + lane - 15;
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) buf[base + 15 * k2] = r[bitrev5(k2)];
    }
    __syncwarp();
}

// Band sums (src/lib.rs:65-82) over spectra in the warp buffer: lane l covers bins bp_b0[l] .. +bp_n[l] of segment
// bp_seg[l]; the (1 - frac) parts go to band seg, the frac parts to band seg + 1.  NQ = 3: |X|^2, |P|^2, Re(X conj P);
// NQ = 1: |X|^2 only.  Results: lane b < 22 returns band b in o[0..NQ) (first / last band doubled).  sc: >= 21 * 6 + 2 floats.
template <int NQ>
__device__ __forceinline__ void band_sums_warp(const float2* __restrict__ buf, const DeviceTables* __restrict__ tab, float* __restrict__ sc,
                                               int lane, float (&o)[NQ]) {
    const int sg = tab->bp_seg[lane], b0 = tab->bp_b0[lane], n = tab->bp_n[lane], rot = tab->bp_rot[lane], off0 = tab->bp_off[lane];
    const float inv = tab->bp_inv[lane];  // frac = j / size as j * (1 / size): within an ulp of the reference's quotient
    float a[NQ], b[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) a[q] = b[q] = 0.0f;
#pragma unroll
    for (int t = 0; t < BP_MAXBINS; t++) {
        if (t < n) {
            int i = t + rot;  // rotated walk: neighbouring lanes start at different offsets -> different banks
            if (i >= n) i -= n;
            const int k = b0 + i;
            const float f = (float)(off0 + i) * inv, g = 1.0f - f;
            const float2 x = buf[k];
            float e[NQ];
            e[0] = x.x * x.x + x.y * x.y;
            if (NQ == 3) {
                const float2 p = buf[ZP_OFF + k];
                e[1] = p.x * p.x + p.y * p.y;
                e[2] = x.x * p.x + x.y * p.y;
            }
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                a[q] = fmaf(g, e[q], a[q]);
                b[q] = fmaf(f, e[q], b[q]);
            }
        }
    }
    // lanes of one segment are contiguous (at most 4): the first one collects the others' ORIGINAL partial sums in a
    // fixed order (lane + 1, + 2, + 3)
    {
        float a0[NQ], b0v[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            a0[q] = a[q];
            b0v[q] = b[q];
        }
#pragma unroll
        for (int off = 1; off < 4; off++) {
            const int osg = __shfl_down_sync(0xffffffffu, sg, off);
            const bool take = lane + off < 32 && osg == sg;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const float va = __shfl_down_sync(0xffffffffu, a0[q], off), vb = __shfl_down_sync(0xffffffffu, b0v[q], off);
                if (take) {
                    a[q] += va;
                    b[q] += vb;
                }
            }
        }
    }
    const int psg = __shfl_up_sync(0xffffffffu, sg, 1);
    if (lane == 0 || psg != sg) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            sc[SC_BAND + sg * 6 + q] = a[q];
            sc[SC_BAND + sg * 6 + 3 + q] = b[q];
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        float v = 0.0f;
        if (lane < NB_BANDS) {
            if (lane < NB_BANDS - 1) v = sc[SC_BAND + lane * 6 + q];
            if (lane > 0) v += sc[SC_BAND + (lane - 1) * 6 + 3 + q];
            if (lane == 0 || lane == NB_BANDS - 1) v *= 2.0f;
        }
        o[q] = v;
    }
    __syncwarp();
}

// ================================================================================================
// K3: analysis -- X, P, band energies, features (src/features.rs:115-219)
// ================================================================================================
__global__ void __launch_bounds__(WPB * 32, 6) analysis_warp_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab, int hbase) {
    __shared__ __align__(16) float2 sbuf[WPB][WBUF];
    __shared__ float ssc[WPB][WSC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * WPB + warp;
    if (s >= bb.n_streams) return;  // no block barrier below: a warp may leave
    float2* buf = sbuf[warp];
    float* sc = ssc[warp];
    const float* h = bb.hist + (size_t)s * HIST_CAP;
    const int pitch = bb.pitch[s];

    // X = rfft(window * input_mem[768..1728]),  P = rfft(window * input_mem[768-pitch .. 1728-pitch]) (src/features.rs:281-290):
    // lane b takes the complex samples z[32 a + b] = (t[64 a + 2 b], t[64 a + 2 b + 1]); all loads issued before use.
    float2 vx[15], vp[15];
    {
        int sx = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE);  // even (hbase is a multiple of 4): a pair never straddles the wrap
        if (sx >= HIST_CAP) sx -= HIST_CAP;
        int sp = hbase + (PITCH_BUF_SIZE - WINDOW_SIZE) - pitch;  // >= 0 since pitch <= 768
        if (sp >= HIST_CAP) sp -= HIST_CAP;
        const bool even = (pitch & 1) == 0;
        float2 hx[15], hp[15], wv[15];
#pragma unroll
        for (int a = 0; a < 15; a++) {
            const int n2 = 64 * a + 2 * lane;
            int px = sx + n2;
            if (px >= HIST_CAP) px -= HIST_CAP;
            hx[a] = __ldg(reinterpret_cast<const float2*>(h + px));
            wv[a] = __ldg(reinterpret_cast<const float2*>(tab->window + n2));
            int p0 = sp + n2;
            if (p0 >= HIST_CAP) p0 -= HIST_CAP;
            if (even) {
                hp[a] = __ldg(reinterpret_cast<const float2*>(h + p0));
            } else {
                int p1 = p0 + 1;
                if (p1 >= HIST_CAP) p1 -= HIST_CAP;
                hp[a] = make_float2(__ldg(h + p0), __ldg(h + p1));
            }
        }
        {  // next wave: the 1728 ring samples behind both windows (54-55 lines of 128 B) and the cepstral ring (6 lines)
            const int sn = s + PF_WAVE_A;
            if (sn < bb.n_streams) {
                const char* hn = reinterpret_cast<const char*>(bb.hist + (size_t)sn * HIST_CAP);
                const char* cn = reinterpret_cast<const char*>(bb.ceps_mem + (size_t)sn * CEPS_MEM * NB_BANDS);
                int o0 = hbase * 4 + 128 * lane;
                if (o0 >= HIST_CAP * 4) o0 -= HIST_CAP * 4;
                prefetch_l2(hn + o0);
                const int l2 = lane + 32;
                if (l2 < 55) {
                    int o1 = hbase * 4 + 128 * l2;
                    if (o1 >= HIST_CAP * 4) o1 -= HIST_CAP * 4;
                    prefetch_l2(hn + o1);
                } else if (l2 < 61) {
                    prefetch_l2(cn + 128 * (l2 - 55));
                } else if (l2 == 61) {
                    prefetch_l2(bb.pitch + sn);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 15; a++) {
            vx[a] = make_float2(hx[a].x * wv[a].x, hx[a].y * wv[a].y);
            vp[a] = make_float2(hp[a].x * wv[a].x, hp[a].y * wv[a].y);
        }
    }
    fft480_warp<true>(vx, vp, buf, tab, lane);

    // even/odd split into the 481 bins, in place (a lane owns bins k and 480 - k of both spectra), spectra to HBM
    {
        const float wn = tab->wnorm;
        float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
        float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = lane + 32 * j;
            if (k <= 240) {
                const int kc = k == 0 ? 0 : 480 - k;
                const float2 tw = __ldg(&tab->tw960[k]);
                float2 x0, x1, p0, p1;
                rfft_split_pair(buf[k], buf[kc], tw, wn, k == 0, x0, x1);
                rfft_split_pair(buf[ZP_OFF + k], buf[ZP_OFF + kc], tw, wn, k == 0, p0, p1);
                buf[k] = x0;
                buf[ZP_OFF + k] = p0;
                Xg[k] = x0;
                Pg[k] = p0;  // k <= 240 < 400
                if (k != 240) {
                    buf[480 - k] = x1;
                    buf[ZP_OFF + 480 - k] = p1;
                    Xg[480 - k] = x1;
                    if (480 - k < NB_BINS_BANDED) Pg[480 - k] = p1;
                }
            }
        }
    }
    __syncwarp();

    float bs[3];
    band_sums_warp<3>(buf, tab, sc, lane, bs);

    // ---- features (src/features.rs:134-219): 22 bands, lane = band; the warp buffer is free from here on ----
    const bool bl = lane < NB_BANDS;
    const float ex = bs[0], ep = bs[1];
    const float xpn = bl ? bs[2] / sqrtf(0.001f + ex * ep) : 0.0f;
    if (bl) {
        bb.ex[(size_t)s * NB_BANDS + lane] = ex;
        bb.ep[(size_t)s * NB_BANDS + lane] = ep;
        bb.exp[(size_t)s * NB_BANDS + lane] = xpn;
    }
    float* fsc = reinterpret_cast<float*>(buf);            // s_ceps [8][22] | s_dist [8][8] | s_feat [42]
    float* s_ceps = fsc;
    float* s_dist = fsc + CEPS_MEM * NB_BANDS;
    float* s_feat = s_dist + CEPS_MEM * CEPS_MEM;
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
        if (lane + 32 * k < CEPS_MEM * NB_BANDS) s_ceps[lane + 32 * k] = cr[k];
    __syncwarp();
    if (bl) {
        s_ceps[mem_id * NB_BANDS + lane] = ceps;
        cg[mem_id * NB_BANDS + lane] = ceps;
        s_feat[lane] = ceps;
    }
    __syncwarp();
    if (lane < NB_DELTA_CEPS) {
        const int c1 = (mem_id < 1) ? CEPS_MEM + mem_id - 1 : mem_id - 1;
        const int c2 = (mem_id < 2) ? CEPS_MEM + mem_id - 2 : mem_id - 2;
        const float a = s_ceps[mem_id * NB_BANDS + lane], b = s_ceps[c1 * NB_BANDS + lane], c = s_ceps[c2 * NB_BANDS + lane];
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
            const float t = s_ceps[i * NB_BANDS + k] - s_ceps[j * NB_BANDS + k];
            dist += t * t;
        }
        s_dist[i * CEPS_MEM + j] = dist;
    }
    __syncwarp();
    float md = 1e15f;
    if (lane < CEPS_MEM) {
#pragma unroll
        for (int j = 0; j < CEPS_MEM; j++)
            if (j != lane) md = fminf(md, s_dist[lane * CEPS_MEM + j]);
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

// ================================================================================================
// K5: synthesis -- pitch filter, gain floor, band-gain interpolation, inverse FFT, overlap-add
// (src/denoise.rs:102-115, src/features.rs:223-275)
// ================================================================================================
// TOut = float, or short: clamp to the int16 range then round half away from zero (what both reference front-ends do:
// src/nnnoiseless.rs:152 `clamp().round() as i16`, test_data/rnnoise_demo.c:53 roundf).
__device__ __forceinline__ short to_pcm16w(float v) { return (short)roundf(fminf(fmaxf(v, -32768.0f), 32767.0f)); }

template <typename TOut>
__global__ void __launch_bounds__(WPB * 32, 5) synthesis_warp_kernel(BatchBuffers bb, const DeviceTables* __restrict__ tab,
                                                                  TOut* __restrict__ out, long stream_stride, long sample_stride,
                                                                  float* __restrict__ vad_out) {
    __shared__ __align__(16) float2 sbuf[WPB][WBUF];
    __shared__ float ssc[WPB][WSC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * WPB + warp;
    if (s >= bb.n_streams) return;
    float2* buf = sbuf[warp];
    float* sc = ssc[warp];

    // every global load of the stream is issued up front
    const float2* Xg = bb.X + (size_t)s * FREQ_SIZE;
    const float2* Pg = bb.P + (size_t)s * NB_BINS_BANDED;
    float2 xv[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int k = lane + 32 * j;
        xv[j] = k <= 480 ? __ldg(Xg + k) : make_float2(0.f, 0.f);
    }
    const float vad_in = bb.vad[s];
    const int silent = bb.silence[s];
    float* smem_ola = bb.synth_mem + (size_t)s * FRAME_SIZE;
    float4 ola[4];  // overlap memory, float4 q = lane + 32 j < 120
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int q = lane + 32 * j;
        ola[j] = q < FRAME_SIZE / 4 ? __ldg(reinterpret_cast<const float4*>(smem_ola) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!silent) {
        float2 pv[13];
#pragma unroll
        for (int j = 0; j < 13; j++) {
            const int k = lane + 32 * j;
            pv[j] = k < NB_BINS_BANDED ? __ldg(Pg + k) : make_float2(0.f, 0.f);
        }
        float b_ex = 0.0f;
        if (lane < NB_BANDS) {
            const float e = bb.exp[(size_t)s * NB_BANDS + lane], g = bb.gains[(size_t)s * NB_BANDS + lane];
            const float ex = bb.ex[(size_t)s * NB_BANDS + lane], ep = bb.ep[(size_t)s * NB_BANDS + lane];
            const float lastg = bb.lastg[(size_t)s * NB_BANDS + lane];
            b_ex = ex;
            // r (src/features.rs:226-235)
            float r;
            if (e > g) {
                r = 1.0f;
            } else {
                const float e2 = e * e, g2 = g * g;
                r = e2 * (1.0f - g2) / (0.001f + g2 * (1.0f - e2));
            }
            r = (r < 0.0f) ? 0.0f : r;
            r = (r > 1.0f) ? 1.0f : r;
            r = sqrtf(r);
            r *= sqrtf(ex / (1e-8f + ep));
            sc[SC_SR + lane] = r;
            // gain floor (src/denoise.rs:106-109)
            const float gg = fmaxf(g, 0.6f * lastg);
            sc[SC_SG + lane] = gg;
            bb.lastg[(size_t)s * NB_BANDS + lane] = gg;
        }
        __syncwarp();
        // x += rf * p  (bin 0 is the real-valued DC offset; its imaginary part stays 0); spectrum to the warp buffer
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k = lane + 32 * j;
            if (j < 13 && k < NB_BINS_BANDED) {
                const int bi = __ldg(&tab->band_of[k]);
                const float fr = __ldg(&tab->band_frac[k]);
                const float rf = (1.0f - fr) * sc[SC_SR + bi] + fr * sc[SC_SR + bi + 1];
                xv[j].x += pv[j].x * rf;
                if (k > 0) xv[j].y += pv[j].y * rf;
            }
            if (k <= 480) buf[k] = xv[j];
        }
        __syncwarp();
        float ne[1];
        band_sums_warp<1>(buf, tab, sc, lane, ne);
        if (lane < NB_BANDS) sc[SC_SN + lane] = sqrtf(b_ex / (1e-8f + ne[0]));
        __syncwarp();
        // x *= rf2 ; x *= gf   (bins >= 400 are zeroed by both interpolations)
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k = lane + 32 * j;
            if (k <= 480) {
                float2 x = xv[j];
                if (k < NB_BINS_BANDED) {
                    const int bi = __ldg(&tab->band_of[k]);
                    const float fr = __ldg(&tab->band_frac[k]);
                    const float m1 = (1.0f - fr) * sc[SC_SN + bi] + fr * sc[SC_SN + bi + 1];
                    const float m2 = (1.0f - fr) * sc[SC_SG + bi] + fr * sc[SC_SG + bi + 1];
                    x.x = (x.x * m1) * m2;
                    x.y = (x.y * m1) * m2;
                } else {
                    x = make_float2(0.f, 0.f);
                }
                buf[k] = x;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k = lane + 32 * j;
            if (k <= 480) buf[k] = xv[j];
        }
    }
    __syncwarp();

    // ---- inverse real FFT (unnormalised), src/features.rs:263-275: Z = 2E + i 2O, fed conjugated to the forward FFT;
    // a lane builds Z[k] and Z[480-k] together, in place ----
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int k = lane + 32 * j;
        if (k <= 240) {
            float2 z0, z1;
            irfft_pretwist_pair(buf[k], buf[480 - k], __ldg(&tab->tw960[k]), k == 0, z0, z1);
            buf[k] = z0;
            if (k != 0 && k != 240) buf[480 - k] = z1;
        }
    }
    __syncwarp();
    float2 v[15], dummy[15];
#pragma unroll
    for (int a = 0; a < 15; a++) v[a] = buf[32 * a + lane];
    __syncwarp();  // everybody has its inputs: the buffer becomes the transpose area
    fft480_warp<false>(v, dummy, buf, tab, lane);

    // time samples 4q..4q+3 = (re, -im) of buf[2q], buf[2q+1]; first half -> output (+ overlap memory), second half -> new
    // overlap memory (every lane read its part of the old one at the top).  Vector stores when the caller's rows allow.
    TOut* o = out + (long)s * stream_stride;
    const bool o_vec = sample_stride == 1 && ((reinterpret_cast<uintptr_t>(o) & (4 * sizeof(TOut) - 1)) == 0);
    const long ss = sample_stride;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int q = lane + 32 * j;
        if (q < WINDOW_SIZE / 4) {
            const float4 z = reinterpret_cast<const float4*>(buf)[q];
            const float4 w = __ldg(reinterpret_cast<const float4*>(tab->window) + q);
            const float4 t = make_float4((z.x * 0.5f) * w.x, (-z.y * 0.5f) * w.y, (z.z * 0.5f) * w.z, (-z.w * 0.5f) * w.w);
            if (q < FRAME_SIZE / 4) {
                const float4 m = ola[j < 4 ? j : 0];
                const float4 r = make_float4(t.x + m.x, t.y + m.y, t.z + m.z, t.w + m.w);
                if (sizeof(TOut) == 4) {
                    float* of = reinterpret_cast<float*>(o);
                    if (o_vec) {
                        reinterpret_cast<float4*>(of)[q] = r;
                    } else {
                        of[(4 * q) * ss] = r.x; of[(4 * q + 1) * ss] = r.y; of[(4 * q + 2) * ss] = r.z; of[(4 * q + 3) * ss] = r.w;
                    }
                } else {
                    short* os = reinterpret_cast<short*>(o);
                    const short p0 = to_pcm16w(r.x), p1 = to_pcm16w(r.y), p2 = to_pcm16w(r.z), p3 = to_pcm16w(r.w);
                    if (o_vec) {
                        reinterpret_cast<uint2*>(os)[q] = make_uint2((unsigned)(unsigned short)p0 | ((unsigned)(unsigned short)p1 << 16),
                                                                     (unsigned)(unsigned short)p2 | ((unsigned)(unsigned short)p3 << 16));
                    } else {
                        os[(4 * q) * ss] = p0; os[(4 * q + 1) * ss] = p1; os[(4 * q + 2) * ss] = p2; os[(4 * q + 3) * ss] = p3;
                    }
                }
            } else {
                reinterpret_cast<float4*>(smem_ola)[q - FRAME_SIZE / 4] = t;
            }
        }
    }
    if (lane == 0 && vad_out) vad_out[s] = silent ? 0.0f : vad_in;
}

}  // namespace

cudaError_t launch_analysis_warp(const BatchBuffers& b, const DeviceTables* tab, int slot, cudaStream_t st) {
    const int grid = (b.n_streams + WPB - 1) / WPB;
    analysis_warp_kernel<<<grid, WPB * 32, 0, st>>>(b, tab, hist_base(slot));
    return cudaGetLastError();
}

cudaError_t launch_synthesis_warp(const BatchBuffers& b, const DeviceTables* tab, void* out, bool pcm16, long stream_stride,
                                  long sample_stride, float* vad_out, cudaStream_t st) {
    const int grid = (b.n_streams + WPB - 1) / WPB;
    if (pcm16) synthesis_warp_kernel<short><<<grid, WPB * 32, 0, st>>>(b, tab, static_cast<short*>(out), stream_stride, sample_stride, vad_out);
    else synthesis_warp_kernel<float><<<grid, WPB * 32, 0, st>>>(b, tab, static_cast<float*>(out), stream_stride, sample_stride, vad_out);
    return cudaGetLastError();
}

}  // namespace nnb
