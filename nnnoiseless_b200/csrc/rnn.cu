// rnn.cu -- the GRU network (src/rnn.rs:251-379), batched ACROSS streams: one block advances TS = 32 streams
// through input_dense -> vad_gru -> vad_output -> noise_gru -> denoise_gru -> denoise_output.  Every layer is a
// [TS x K] x [K x O] tile product held in registers: a thread owns 4 outputs x 4 streams (16 accumulators) and per
// input j issues one 128-bit weight load (L1/L2-resident f32-expanded int8 weights), one 128-bit shared-memory
// load of the 4 streams' activations and 16 FMAs -- so each weight fetched is reused for 4 streams from
// registers and for all 32 streams of the block from L1.
//
// Activations: src/util.rs:29-53 (table tanh, sigmoid = .5 + .5 tanh(x/2), relu), chosen per layer at run time
// from the model header.  GRU semantics: src/rnn.rs:292-327 (reset gate applied to the state BEFORE the
// recurrent product; gate order z | r | h).  f32 with FMA; compared with the oracle within tolerance.
#include "common.cuh"

namespace nnb {

namespace {

#ifndef RNN_TS
#define RNN_TS 32
#endif
#ifndef RNN_RT
#define RNN_RT 128
#endif
#ifndef RNN_UNROLL
#define RNN_UNROLL 4
#endif
constexpr int TS = RNN_TS;    // streams per block
constexpr int RT = RNN_RT;    // threads per block
constexpr int SG = TS / 4;  // stream groups of 4
constexpr int UNR = RNN_UNROLL;
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

// out[o][s] = bias[o] + sum_j W[j][o] * xin[j][s]   for o < OP (multiple of 4), s < TS.
// W: global [K][OP]; xin, out: shared [rows][TS].
__device__ __forceinline__ void tile_gemm(const float* __restrict__ W, int K, int OP, const float* __restrict__ bias,
                                          const float* xin, float* out) {
    const int nitems = (OP >> 2) * SG;
    for (int item = threadIdx.x; item < nitems; item += RT) {
        const int oq = item / SG, sg = item - oq * SG;
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + oq);
        float acc[4][4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            acc[0][s] = b.x;
            acc[1][s] = b.y;
            acc[2][s] = b.z;
            acc[3][s] = b.w;
        }
        const float4* wp = reinterpret_cast<const float4*>(W) + oq;
        const float4* xp = reinterpret_cast<const float4*>(xin) + sg;
        const int wstride = OP >> 2;
#pragma unroll UNR
        for (int j = 0; j < K; j++) {
            const float4 w = __ldg(wp + (size_t)j * wstride);
            const float4 x = xp[j * (TS / 4)];
            const float wv[4] = {w.x, w.y, w.z, w.w};
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int s = 0; s < 4; s++) acc[a][s] = fmaf(wv[a], xv[s], acc[a][s]);
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
            *reinterpret_cast<float4*>(out + (4 * oq + a) * TS + 4 * sg) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }
}

// One GRU layer on a tile.  xin rows [0, ni) hold the layer input; h: state [nn][TS]; zr: scratch [3*np][TS].
__device__ void gru_tile(const DeviceLayer& L, float* xin, float* h, float* zr, const float* __restrict__ table) {
    const int ni = L.ni, nn = L.nn, np = L.np, K = ni + nn;
    for (int i = threadIdx.x; i < nn * TS; i += RT) xin[ni * TS + i] = h[i];
    __syncthreads();
    tile_gemm(L.w, K, 2 * np, L.bias, xin, zr);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * TS; i += RT) {
        const float z = sigmoid_approx(WEIGHTS_SCALE * zr[i], table);
        const float r = sigmoid_approx(WEIGHTS_SCALE * zr[np * TS + i], table);
        zr[i] = z;
        xin[ni * TS + i] = h[i] * r;
    }
    __syncthreads();
    tile_gemm(L.wh, K, np, L.bias + 2 * np, xin, zr + 2 * np * TS);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * TS; i += RT) {
        const float z = zr[i];
        const float hh = activate(L.act, WEIGHTS_SCALE * zr[2 * np * TS + i], table);
        h[i] = z * h[i] + (1.0f - z) * hh;
    }
    __syncthreads();
}

struct SmemPlan {
    int feat, dense, hv, hn, hd, xin, zr, outb, total;  // offsets in floats
};

__host__ __device__ inline SmemPlan plan_smem(const DeviceModel& m) {
    SmemPlan p;
    int o = 208;  // tansig table
    p.feat = o;   o += 42 * TS;
    p.dense = o;  o += m.input_dense.np * TS;
    p.hv = o;     o += m.vad_gru.nn * TS;
    p.hn = o;     o += m.noise_gru.nn * TS;
    p.hd = o;     o += m.denoise_gru.nn * TS;
    int kmax = m.vad_gru.ni + m.vad_gru.nn;
    if (m.noise_gru.ni + m.noise_gru.nn > kmax) kmax = m.noise_gru.ni + m.noise_gru.nn;
    if (m.denoise_gru.ni + m.denoise_gru.nn > kmax) kmax = m.denoise_gru.ni + m.denoise_gru.nn;
    int npmax = m.vad_gru.np;
    if (m.noise_gru.np > npmax) npmax = m.noise_gru.np;
    if (m.denoise_gru.np > npmax) npmax = m.denoise_gru.np;
    p.xin = o;    o += kmax * TS;
    p.zr = o;     o += 3 * npmax * TS;
    p.outb = o;   o += 24 * TS;
    p.total = o;
    return p;
}

__global__ void __launch_bounds__(RT) rnn_kernel(BatchBuffers bb, DeviceModel m, const DeviceTables* __restrict__ tab) {
    extern __shared__ __align__(16) float sm[];
    const SmemPlan pl = plan_smem(m);
    const int nd = m.input_dense.nn, nv = m.vad_gru.nn, nn = m.noise_gru.nn, ndn = m.denoise_gru.nn;
    float* table = sm;
    float* feat = sm + pl.feat;    // [42][TS]
    float* dense = sm + pl.dense;  // [nd pad][TS]
    float* hv = sm + pl.hv;        // [nv][TS]   hv | hn | hd contiguous
    float* hn = sm + pl.hn;
    float* hd = sm + pl.hd;
    float* xin = sm + pl.xin;
    float* zr = sm + pl.zr;
    float* outb = sm + pl.outb;

    const int s0 = blockIdx.x * TS, tid = threadIdx.x;
    const int ns = min(TS, bb.n_streams - s0);
    const int SS = m.state_size;

    for (int i = tid; i < 201; i += RT) table[i] = tab->tansig[i];
    for (int i = tid; i < 42 * TS; i += RT) {
        int s = i / 42, j = i - s * 42;
        feat[j * TS + s] = (s < ns) ? bb.features[(size_t)(s0 + s) * NB_FEATURES + j] : 0.0f;
    }
    for (int i = tid; i < SS * TS; i += RT) {
        int s = i / SS, j = i - s * SS;
        hv[j * TS + s] = (s < ns) ? bb.gru_state[(size_t)(s0 + s) * SS + j] : 0.0f;
    }
    __syncthreads();

    // input_dense (src/rnn.rs:353-355)
    tile_gemm(m.input_dense.w, 42, m.input_dense.np, m.input_dense.bias, feat, dense);
    __syncthreads();
    for (int i = tid; i < nd * TS; i += RT) dense[i] = activate(m.input_dense.act, WEIGHTS_SCALE * dense[i], table);
    __syncthreads();

    // vad_gru (src/rnn.rs:356-358)
    for (int i = tid; i < nd * TS; i += RT) xin[i] = dense[i];
    gru_tile(m.vad_gru, xin, hv, zr, table);

    // vad_output (src/rnn.rs:359)
    tile_gemm(m.vad_output.w, nv, m.vad_output.np, m.vad_output.bias, hv, outb);
    __syncthreads();
    if (tid < ns && !bb.silence[s0 + tid]) bb.vad[s0 + tid] = activate(m.vad_output.act, WEIGHTS_SCALE * outb[tid], table);

    // noise_gru input = [dense | vad_state | features] (src/rnn.rs:361-366)
    for (int i = tid; i < nd * TS; i += RT) xin[i] = dense[i];
    for (int i = tid; i < nv * TS; i += RT) xin[nd * TS + i] = hv[i];
    for (int i = tid; i < 42 * TS; i += RT) xin[(nd + nv) * TS + i] = feat[i];
    gru_tile(m.noise_gru, xin, hn, zr, table);

    // denoise_gru input = [vad_state | noise_state | features] (src/rnn.rs:368-377)
    for (int i = tid; i < nv * TS; i += RT) xin[i] = hv[i];
    for (int i = tid; i < nn * TS; i += RT) xin[nv * TS + i] = hn[i];
    for (int i = tid; i < 42 * TS; i += RT) xin[(nv + nn) * TS + i] = feat[i];
    gru_tile(m.denoise_gru, xin, hd, zr, table);

    // denoise_output (src/rnn.rs:378)
    tile_gemm(m.denoise_output.w, ndn, m.denoise_output.np, m.denoise_output.bias, hd, outb);
    __syncthreads();
    for (int i = tid; i < NB_BANDS * TS; i += RT) {
        int s = i / NB_BANDS, j = i - s * NB_BANDS;
        if (s < ns && !bb.silence[s0 + s])
            bb.gains[(size_t)(s0 + s) * NB_BANDS + j] = activate(m.denoise_output.act, WEIGHTS_SCALE * outb[j * TS + s], table);
    }
    // state write-back; silent frames leave the RNN state untouched (src/denoise.rs:102)
    for (int i = tid; i < SS * TS; i += RT) {
        int s = i / SS, j = i - s * SS;
        if (s < ns && !bb.silence[s0 + s]) bb.gru_state[(size_t)(s0 + s) * SS + j] = hv[j * TS + s];
    }
}

}  // namespace

cudaError_t launch_rnn(const BatchBuffers& b, const DeviceModel& m, const DeviceTables* tab, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)plan_smem(m).total;
    static size_t attr_smem[64] = {0};  // per device: largest dynamic smem size already enabled
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || smem > attr_smem[dev]) {
        e = cudaFuncSetAttribute(rnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_smem[dev] = smem;
    }
    int grid = (b.n_streams + TS - 1) / TS;
    rnn_kernel<<<grid, RT, smem, st>>>(b, m, tab);
    return cudaGetLastError();
}

}  // namespace nnb
