// rnn.cu -- the GRU network (src/rnn.rs:251-379), batched ACROSS streams: one block advances
// TS streams through input_dense -> vad_gru -> vad_output -> noise_gru -> denoise_gru ->
// denoise_output, so every weight fetched from L2/L1 is reused TS times from registers
// (a [TS x K] x [K x O] tile product per layer instead of TS independent mat-vecs).
//
// Activations: src/util.rs:29-53 (table tanh, sigmoid = .5 + .5 tanh(x/2), relu), chosen per layer
// at run time from the model header.  GRU semantics: src/rnn.rs:292-327 (reset gate applied to the
// state BEFORE the recurrent product; gate order z | r | h).
#include "common.cuh"

namespace nnb {

constexpr int TS = 8;     // streams per block
constexpr int RT = 128;   // threads per block
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

// out[o][s] = bias[o] + sum_j W[j][o] * xin[j][s]   (xin, out: shared, [rows][TS])
__device__ __forceinline__ void tile_matvec(const float* __restrict__ W, int K, int O, const float* __restrict__ bias,
                                            const float* xin, float* out) {
    for (int o = threadIdx.x; o < O; o += RT) {
        float acc[TS];
        const float b = bias[o];
#pragma unroll
        for (int s = 0; s < TS; s++) acc[s] = b;
        const float* w = W + o;
#pragma unroll 4
        for (int j = 0; j < K; j++) {
            const float wv = __ldg(w + (size_t)j * O);
            const float4 x0 = *reinterpret_cast<const float4*>(xin + j * TS);
            const float4 x1 = *reinterpret_cast<const float4*>(xin + j * TS + 4);
            acc[0] += wv * x0.x; acc[1] += wv * x0.y; acc[2] += wv * x0.z; acc[3] += wv * x0.w;
            acc[4] += wv * x1.x; acc[5] += wv * x1.y; acc[6] += wv * x1.z; acc[7] += wv * x1.w;
        }
        float4* op = reinterpret_cast<float4*>(out + o * TS);
        op[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        op[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// One GRU layer on a tile.  xin rows [0, ni) must already hold the layer input; h: state [nn][TS].
__device__ void gru_tile(const DeviceLayer& L, float* xin, float* h, float* zr, const float* __restrict__ table) {
    const int ni = L.ni, nn = L.nn, K = ni + nn;
    for (int i = threadIdx.x; i < nn * TS; i += RT) xin[ni * TS + i] = h[i];
    __syncthreads();
    tile_matvec(L.w, K, 2 * nn, L.bias, xin, zr);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * TS; i += RT) {
        float z = sigmoid_approx(WEIGHTS_SCALE * zr[i], table);
        float r = sigmoid_approx(WEIGHTS_SCALE * zr[nn * TS + i], table);
        zr[i] = z;
        xin[ni * TS + i] = h[i] * r;
    }
    __syncthreads();
    tile_matvec(L.wh, K, nn, L.bias + 2 * nn, xin, zr + nn * TS);
    __syncthreads();
    for (int i = threadIdx.x; i < nn * TS; i += RT) {
        float z = zr[i];
        float hh = activate(L.act, WEIGHTS_SCALE * zr[nn * TS + i], table);
        h[i] = z * h[i] + (1.0f - z) * hh;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RT) rnn_kernel(BatchBuffers bb, DeviceModel m, const DeviceTables* __restrict__ tab) {
    extern __shared__ __align__(16) float sm[];
    const int nd = m.input_dense.nn, nv = m.vad_gru.nn, nn = m.noise_gru.nn, ndn = m.denoise_gru.nn;
    float* table = sm;                       // 208 floats
    float* feat = table + 208;               // [42][TS]
    float* dense = feat + 42 * TS;           // [nd][TS]
    float* hv = dense + nd * TS;             // [nv][TS]
    float* hn = hv + nv * TS;                // [nn][TS]
    float* hd = hn + nn * TS;                // [ndn][TS]
    float* xin = hd + ndn * TS;              // [256][TS]
    float* zr = xin + 256 * TS;              // [256][TS]
    float* outb = zr + 256 * TS;             // [32][TS]

    const int s0 = blockIdx.x * TS, tid = threadIdx.x;
    const int ns = min(TS, bb.n_streams - s0);
    const int SS = m.state_size;

    for (int i = tid; i < 201; i += RT) table[i] = tab->tansig[i];
    for (int i = tid; i < 42 * TS; i += RT) {
        int s = i / 42, j = i % 42;
        feat[j * TS + s] = (s < ns) ? bb.features[(size_t)(s0 + s) * NB_FEATURES + j] : 0.0f;
    }
    for (int i = tid; i < SS * TS; i += RT) {
        int s = i / SS, j = i % SS;
        float v = (s < ns) ? bb.gru_state[(size_t)(s0 + s) * SS + j] : 0.0f;
        hv[j * TS + s] = v;  // hv | hn | hd are contiguous
    }
    __syncthreads();

    // input_dense (src/rnn.rs:353-355)
    tile_matvec(m.input_dense.w, 42, nd, m.input_dense.bias, feat, dense);
    __syncthreads();
    for (int i = tid; i < nd * TS; i += RT) dense[i] = activate(m.input_dense.act, WEIGHTS_SCALE * dense[i], table);
    __syncthreads();

    // vad_gru (src/rnn.rs:356-358)
    for (int i = tid; i < nd * TS; i += RT) xin[i] = dense[i];
    gru_tile(m.vad_gru, xin, hv, zr, table);

    // vad_output (src/rnn.rs:359)
    tile_matvec(m.vad_output.w, nv, 1, m.vad_output.bias, hv, outb);
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
    tile_matvec(m.denoise_output.w, ndn, NB_BANDS, m.denoise_output.bias, hd, outb);
    __syncthreads();
    for (int i = tid; i < NB_BANDS * TS; i += RT) {
        int s = i / NB_BANDS, j = i % NB_BANDS;
        if (s < ns && !bb.silence[s0 + s])
            bb.gains[(size_t)(s0 + s) * NB_BANDS + j] = activate(m.denoise_output.act, WEIGHTS_SCALE * outb[j * TS + s], table);
    }
    // state write-back; silent frames leave the RNN state untouched (src/denoise.rs:102)
    for (int i = tid; i < SS * TS; i += RT) {
        int s = i / SS, j = i % SS;
        if (s < ns && !bb.silence[s0 + s]) bb.gru_state[(size_t)(s0 + s) * SS + j] = hv[j * TS + s];
    }
}

static size_t rnn_smem_bytes(const DeviceModel& m) {
    size_t floats = 208 + 42 * TS + (size_t)(m.input_dense.nn + m.state_size) * TS + 256 * TS + 256 * TS + 32 * TS;
    return floats * sizeof(float);
}

cudaError_t launch_rnn(const BatchBuffers& b, const DeviceModel& m, const DeviceTables* tab, cudaStream_t st) {
    const size_t smem = rnn_smem_bytes(m);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(rnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    int grid = (b.n_streams + TS - 1) / TS;
    rnn_kernel<<<grid, RT, smem, st>>>(b, m, tab);
    return cudaGetLastError();
}

}  // namespace nnb
