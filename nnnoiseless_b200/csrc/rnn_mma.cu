// rnn_mma.cu -- the GRU network (src/rnn.rs:251-379) on tensor cores, batched across streams.
//
// One block advances TS = 32 streams.  With 32 streams side by side every layer is a real dense contraction
// [32 x K] x [K x N] (K up to 224, N up to 192), so it runs as mma.sync.m16n8k16 (f16 inputs, f32 accumulate):
//   * weights are int8 -> exactly representable in f16; pre-packed on the host in B-fragment order, so a warp
//     fetches a whole 16x8 fragment with one coalesced 256-byte load and reuses it for both 16-stream row tiles
//     and for the hi and lo halves of the activations;
//   * activations are f32; each is split x = hi + lo (two f16, ~22 significant bits) and both halves are
//     multiplied -- products are exact in f32, only the summation order differs from the reference's;
//   * per GRU a warp owns the same output columns for z, r and the candidate, so z stays in registers and the
//     update h = z h + (1-z) h~ happens in the accumulator layout without a transpose.
// Activations: src/util.rs:29-53 (table tanh, sigmoid = .5 + .5 tanh(x/2), relu) chosen per layer at run time.
// GRU semantics: src/rnn.rs:292-327 (reset gate applied to the state BEFORE the recurrent product).
#include <cuda_fp16.h>

#include "common.cuh"

namespace nnb {

namespace {

constexpr int TS = 32;      // streams per block = two m16 row tiles
#ifndef RNN_NWARP
#define RNN_NWARP 8
#endif
constexpr int NWARP = RNN_NWARP;
constexpr int NT = NWARP * 32;
constexpr int MAXOT = 16 / NWARP;  // output tiles (8 neurons) per warp: NWARP * MAXOT tiles cover layers up to 128 neurons
constexpr float WEIGHTS_SCALE = 1.0f / 256.0f;

__device__ __forceinline__ float tansig_approx(float x, const float* __restrict__ table) {
    if (!(x < 8.0f)) return 1.0f;
    if (!(x > -8.0f)) return -1.0f;
    float sign = 1.0f;
    if (x < 0.0f) {
        x = -x;
        sign = -1.0f;
    }
    float fi = floorf(0.5f + 25.0f * x);
    x -= 0.04f * fi;
    float y = table[(int)fi];
    float dy = 1.0f - y * y;
    y = y + x * dy * (1.0f - y * x);
    return sign * y;
}
__device__ __forceinline__ float sigmoid_approx(float x, const float* __restrict__ table) {
    return 0.5f + 0.5f * tansig_approx(0.5f * x, table);
}
__device__ __forceinline__ float activate(int act, float x, const float* __restrict__ table) {
    if (act == 0) return tansig_approx(x, table);
    if (act == 1) return sigmoid_approx(x, table);
    return fmaxf(x, 0.0f);
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// x = hi + lo with hi, lo in f16 (|x| is clamped to the f16 range: 65504)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    x = fminf(fmaxf(x, -65504.0f), 65504.0f);
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
// two adjacent columns (col even) of one row
__device__ __forceinline__ void store_pair(__half* Ahi, __half* Alo, int idx, float v0, float v1) {
    __half h0, l0, h1, l1;
    split_f16(v0, h0, l0);
    split_f16(v1, h1, l1);
    *reinterpret_cast<__half2*>(Ahi + idx) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(Alo + idx) = __halves2half2(l0, l1);
}

// acc[i][mt][.] += A[:, phase columns] x W[:, tile tiles[i]]  for i < ntl, both 16-stream row tiles mt.
template <int MAXT>
__device__ __forceinline__ void run_tiles(const MmaPhase& ph, const __half* Ahi, const __half* Alo, int kp, int lane,
                                          const int (&tiles)[MAXT], int ntl, float (&acc)[MAXT][2][4]) {
    const int g = lane >> 2, t = lane & 3;
    const uint2* wp[MAXT];  // this lane's slot in the fragments of tile i; one chunk further = + ntiles * 32
#pragma unroll
    for (int i = 0; i < MAXT; i++) wp[i] = ph.wfrag + (size_t)(i < ntl ? tiles[i] : 0) * 32 + lane;
    const int wstep = ph.ntiles * 32;
    uint2 bn[MAXT];
#pragma unroll
    for (int i = 0; i < MAXT; i++) bn[i] = (i < ntl) ? __ldg(wp[i]) : make_uint2(0u, 0u);
    // row bases of this lane's A fragments: rows g and g + 8 of both row tiles, column 2t of the chunk
    const int rb = g * kp + 2 * t;
    const __half* ah0 = Ahi + rb;
    const __half* al0 = Alo + rb;
    const int r8 = 8 * kp, r16 = 16 * kp;
    const int nch = ph.nchunks;
    for (int kc = 0; kc < nch; kc++) {
        uint2 b[MAXT];
#pragma unroll
        for (int i = 0; i < MAXT; i++) b[i] = bn[i];
        if (kc + 1 < nch) {  // prefetch the next chunk's fragments while this chunk's MMAs run
#pragma unroll
            for (int i = 0; i < MAXT; i++) {
                wp[i] += wstep;
                if (i < ntl) bn[i] = __ldg(wp[i]);
            }
        }
        const int col = ph.col[kc];
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const __half* ph_ = ah0 + col + mt * r16;
            const __half* pl_ = al0 + col + mt * r16;
            ah[mt][0] = *reinterpret_cast<const uint32_t*>(ph_);
            ah[mt][1] = *reinterpret_cast<const uint32_t*>(ph_ + r8);
            ah[mt][2] = *reinterpret_cast<const uint32_t*>(ph_ + 8);
            ah[mt][3] = *reinterpret_cast<const uint32_t*>(ph_ + r8 + 8);
            al[mt][0] = *reinterpret_cast<const uint32_t*>(pl_);
            al[mt][1] = *reinterpret_cast<const uint32_t*>(pl_ + r8);
            al[mt][2] = *reinterpret_cast<const uint32_t*>(pl_ + 8);
            al[mt][3] = *reinterpret_cast<const uint32_t*>(pl_ + r8 + 8);
        }
#pragma unroll
        for (int i = 0; i < MAXT; i++) {
            if (i < ntl) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    mma16816(acc[i][mt], ah[mt], b[i]);
                    mma16816(acc[i][mt], al[mt], b[i]);
                }
            }
        }
    }
}

template <int MAXT>
__device__ __forceinline__ void init_bias(const MmaPhase& ph, int lane, const int (&tiles)[MAXT], int ntl, float (&acc)[MAXT][2][4]) {
    const int t = lane & 3;
#pragma unroll
    for (int i = 0; i < MAXT; i++) {
        float b0 = 0.0f, b1 = 0.0f;
        if (i < ntl) {
            b0 = __ldg(ph.bias + tiles[i] * 8 + 2 * t);
            b1 = __ldg(ph.bias + tiles[i] * 8 + 2 * t + 1);
        }
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            acc[i][mt][0] = b0;
            acc[i][mt][1] = b1;
            acc[i][mt][2] = b0;
            acc[i][mt][3] = b1;
        }
    }
}

// One GRU layer for the block's 32 streams.  c_state: A columns of this layer's state; s_off: its offset in Hf.
__device__ void gru_layer(const MmaPhase& pzr, const MmaPhase& ph, int act, int nn, int c_state, int c_rh, int s_off,
                          __half* Ahi, __half* Alo, int kp, float* Hf, int hs, const float* table) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int ot = (nn + 7) >> 3;  // output tiles of this layer
    int own[MAXOT], cnt = 0;
#pragma unroll
    for (int i = 0; i < MAXOT; i++) {
        own[i] = warp + i * NWARP;
        if (own[i] < ot) cnt = i + 1;
    }
    float zreg[MAXOT][2][4];
    {
        // z | r gates: tiles {z_j, r_j} for the owned output tiles j
        int tiles[2 * MAXOT];
#pragma unroll
        for (int i = 0; i < MAXOT; i++) {
            tiles[2 * i] = own[i];
            tiles[2 * i + 1] = ot + own[i];
        }
        float acc[2 * MAXOT][2][4];
        init_bias<2 * MAXOT>(pzr, lane, tiles, 2 * cnt, acc);
        run_tiles<2 * MAXOT>(pzr, Ahi, Alo, kp, lane, tiles, 2 * cnt, acc);
#pragma unroll
        for (int i = 0; i < MAXOT; i++) {
            if (i < cnt) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {  // rows g and g + 8
                        const int row = mt * 16 + g + 8 * hf, o = own[i] * 8 + 2 * t;
                        float rh[2];
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const float z = sigmoid_approx(WEIGHTS_SCALE * acc[2 * i][mt][2 * hf + c], table);
                            const float r = sigmoid_approx(WEIGHTS_SCALE * acc[2 * i + 1][mt][2 * hf + c], table);
                            zreg[i][mt][2 * hf + c] = z;
                            rh[c] = r * Hf[row * hs + s_off + o + c];  // reset gate scales the previous state
                        }
                        store_pair(Ahi, Alo, row * kp + c_rh + o, rh[0], rh[1]);
                    }
                }
            }
        }
    }
    __syncthreads();
    {
        int tiles[MAXOT];
#pragma unroll
        for (int i = 0; i < MAXOT; i++) tiles[i] = own[i];
        float acc[MAXOT][2][4];
        init_bias<MAXOT>(ph, lane, tiles, cnt, acc);
        run_tiles<MAXOT>(ph, Ahi, Alo, kp, lane, tiles, cnt, acc);
#pragma unroll
        for (int i = 0; i < MAXOT; i++) {
            if (i < cnt) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        const int row = mt * 16 + g + 8 * hf, o = own[i] * 8 + 2 * t;
                        float hn[2];
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const float z = zreg[i][mt][2 * hf + c];
                            const float hh = activate(act, WEIGHTS_SCALE * acc[i][mt][2 * hf + c], table);
                            const float hp = Hf[row * hs + s_off + o + c];
                            hn[c] = (o + c < nn) ? z * hp + (1.0f - z) * hh : 0.0f;
                            Hf[row * hs + s_off + o + c] = hn[c];
                        }
                        store_pair(Ahi, Alo, row * kp + c_state + o, hn[0], hn[1]);
                    }
                }
            }
        }
    }
    __syncthreads();
}

#ifndef RNN_MINB
#define RNN_MINB 1
#endif
__global__ void __launch_bounds__(NT, RNN_MINB) rnn_mma_kernel(BatchBuffers bb, DeviceModelMma m, const DeviceTables* __restrict__ tab) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int kp = m.kp, hs = m.hs;
    __half* Ahi = reinterpret_cast<__half*>(smraw);
    __half* Alo = Ahi + TS * kp;
    float* Hf = reinterpret_cast<float*>(Alo + TS * kp);
    float* table = Hf + TS * hs;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int s0 = blockIdx.x * TS;
    const int ns = min(TS, bb.n_streams - s0);
    const int SS = m.state_size;
    const int so_v = 0, so_n = (m.nv + 7) & ~7, so_d = so_n + ((m.nn + 7) & ~7);  // state offsets inside an Hf row

    // zero A (padding columns must hold finite values) and Hf, load the tanh table
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(smraw);
        const int nz = (TS * kp * 2 * 2 + TS * hs * 4) / 4;
        for (int i = tid; i < nz; i += NT) z[i] = 0u;
        for (int i = tid; i < 201; i += NT) table[i] = tab->tansig[i];
    }
    __syncthreads();
    {
        // features [ns][42] and GRU state [ns][SS] of this block are contiguous in HBM: fetch them with 128-bit loads,
        // all requests of a thread in flight before the first use (full blocks with 16-byte aligned bases)
        const float* fsrc = bb.features + (size_t)s0 * NB_FEATURES;
        const float* ssrc = bb.gru_state + (size_t)s0 * SS;
        const bool vec = ns == TS && ((reinterpret_cast<uintptr_t>(fsrc) | reinterpret_cast<uintptr_t>(ssrc)) & 15) == 0 && (SS & 3) == 0;
        auto put_feat = [&](int e, float v) {
            const int s = e / NB_FEATURES, j = e - s * NB_FEATURES;
            __half hi, lo;
            split_f16(v, hi, lo);
            Ahi[s * kp + m.c_feat + j] = hi;
            Alo[s * kp + m.c_feat + j] = lo;
        };
        auto put_state = [&](int e, float v) {
            const int s = e / SS, j = e - s * SS;
            int ac, ho;
            if (j < m.nv) { ac = m.c_vad + j; ho = so_v + j; }
            else if (j < m.nv + m.nn) { ac = m.c_noise + (j - m.nv); ho = so_n + (j - m.nv); }
            else { ac = m.c_den + (j - m.nv - m.nn); ho = so_d + (j - m.nv - m.nn); }
            Hf[s * hs + ho] = v;
            __half hi, lo;
            split_f16(v, hi, lo);
            Ahi[s * kp + ac] = hi;
            Alo[s * kp + ac] = lo;
        };
        if (vec) {
            constexpr int NF4 = TS * NB_FEATURES / 4;          // 336
            constexpr int FQ = (NF4 + NT - 1) / NT;            // 3 per thread
            float4 fv[FQ];
#pragma unroll
            for (int k = 0; k < FQ; k++) {
                const int q = tid + k * NT;
                fv[k] = q < NF4 ? __ldg(reinterpret_cast<const float4*>(fsrc) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const int ns4 = TS * SS / 4;
            for (int q0 = 0; q0 < ns4; q0 += 8 * NT) {
                float4 sv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int q = q0 + tid + k * NT;
                    sv[k] = q < ns4 ? __ldg(reinterpret_cast<const float4*>(ssrc) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int q = q0 + tid + k * NT;
                    if (q < ns4) {
                        put_state(4 * q, sv[k].x);
                        put_state(4 * q + 1, sv[k].y);
                        put_state(4 * q + 2, sv[k].z);
                        put_state(4 * q + 3, sv[k].w);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < FQ; k++) {
                const int q = tid + k * NT;
                if (q < NF4) {
                    put_feat(4 * q, fv[k].x);
                    put_feat(4 * q + 1, fv[k].y);
                    put_feat(4 * q + 2, fv[k].z);
                    put_feat(4 * q + 3, fv[k].w);
                }
            }
        } else {
            for (int i = tid; i < ns * NB_FEATURES; i += NT) put_feat(i, fsrc[i]);
            for (int i = tid; i < ns * SS; i += NT) put_state(i, ssrc[i]);
        }
    }
    __syncthreads();

    // ---- input_dense (src/rnn.rs:353-355) ----
    {
        const int ntile = (m.nd + 7) >> 3;
        int tiles[MAXOT], cnt = 0;
#pragma unroll
        for (int i = 0; i < MAXOT; i++) {
            tiles[i] = warp + i * NWARP;
            if (tiles[i] < ntile) cnt = i + 1;
        }
        float acc[MAXOT][2][4];
        init_bias<MAXOT>(m.dense, lane, tiles, cnt, acc);
        run_tiles<MAXOT>(m.dense, Ahi, Alo, kp, lane, tiles, cnt, acc);
#pragma unroll
        for (int i = 0; i < MAXOT; i++)
            if (i < cnt)
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        const int row = mt * 16 + g + 8 * hf, o = tiles[i] * 8 + 2 * t;
                        float v[2];
#pragma unroll
                        for (int c = 0; c < 2; c++)
                            v[c] = (o + c < m.nd) ? activate(m.act_dense, WEIGHTS_SCALE * acc[i][mt][2 * hf + c], table) : 0.0f;
                        store_pair(Ahi, Alo, row * kp + m.c_dense + o, v[0], v[1]);
                    }
    }
    __syncthreads();

    // ---- vad_gru (src/rnn.rs:356-358) ----
    gru_layer(m.vad_zr, m.vad_h, m.act_vad, m.nv, m.c_vad, m.c_rh, so_v, Ahi, Alo, kp, Hf, hs, table);

    // ---- vad_output (src/rnn.rs:359): one neuron -> tile 0, warp 0 (no barrier needed: it only reads the vad state) ----
    if (warp == 0) {
        int tiles[1] = {0};
        float acc[1][2][4];
        init_bias<1>(m.vad_out, lane, tiles, 1, acc);
        run_tiles<1>(m.vad_out, Ahi, Alo, kp, lane, tiles, 1, acc);
        if (t == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const int row = mt * 16 + g + 8 * hf;
                    if (row < ns && !bb.silence[s0 + row])
                        bb.vad[s0 + row] = activate(m.act_vadout, WEIGHTS_SCALE * acc[0][mt][2 * hf], table);
                }
        }
    }

    // ---- noise_gru, denoise_gru (src/rnn.rs:361-377) ----
    gru_layer(m.noise_zr, m.noise_h, m.act_noise, m.nn, m.c_noise, m.c_rh, so_n, Ahi, Alo, kp, Hf, hs, table);
    gru_layer(m.den_zr, m.den_h, m.act_den, m.ndn, m.c_den, m.c_rh, so_d, Ahi, Alo, kp, Hf, hs, table);

    // ---- denoise_output (src/rnn.rs:378): 22 band gains ----
    {
        int tiles[1] = {warp};
        if (warp < (NB_BANDS + 7) / 8) {
            float acc[1][2][4];
            init_bias<1>(m.out, lane, tiles, 1, acc);
            run_tiles<1>(m.out, Ahi, Alo, kp, lane, tiles, 1, acc);
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const int row = mt * 16 + g + 8 * hf;
                    if (row < ns && !bb.silence[s0 + row]) {
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const int o = warp * 8 + 2 * t + c;
                            if (o < NB_BANDS)
                                bb.gains[(size_t)(s0 + row) * NB_BANDS + o] =
                                    activate(m.act_out, WEIGHTS_SCALE * acc[0][mt][2 * hf + c], table);
                        }
                    }
                }
        }
    }
    // ---- state write-back; silent frames leave the RNN state untouched (src/denoise.rs:102) ----
    for (int i = tid; i < TS * SS; i += NT) {
        const int s = i / SS, j = i - s * SS;
        if (s < ns && !bb.silence[s0 + s]) {
            int ho;
            if (j < m.nv) ho = so_v + j;
            else if (j < m.nv + m.nn) ho = so_n + (j - m.nv);
            else ho = so_d + (j - m.nv - m.nn);
            bb.gru_state[(size_t)(s0 + s) * SS + j] = Hf[s * hs + ho];
        }
    }
}

}  // namespace

cudaError_t launch_rnn_mma(const BatchBuffers& b, const DeviceModelMma& m, const DeviceTables* tab, cudaStream_t st) {
    const size_t smem = (size_t)TS * m.kp * 2 * 2 + (size_t)TS * m.hs * 4 + 208 * 4;
    static size_t attr_smem[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || smem > attr_smem[dev]) {
        e = cudaFuncSetAttribute(rnn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_smem[dev] = smem;
    }
    const int grid = (b.n_streams + TS - 1) / TS;
    rnn_mma_kernel<<<grid, NT, smem, st>>>(b, m, tab);
    return cudaGetLastError();
}

}  // namespace nnb
