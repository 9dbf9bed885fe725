// exact.cu -- ORDER-EXACT high-pass biquad (the pitch analysis, also order-exact, is in pitch.cu).
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
    static unsigned long long attr_devs = 0;  // bit d: attribute set on device d
    const size_t smem = sizeof(float) * HP_STREAMS * HP_LD;
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (dev >= 64 || !((attr_devs >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(hp_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_devs |= 1ull << dev;
    }
    int vec_ok = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (stream_stride % 4 == 0);
    int grid = (b.n_streams + HP_STREAMS - 1) / HP_STREAMS;
    hp_filter_kernel<<<grid, HP_THREADS, smem, st>>>(in, stream_stride, b.hist, b.hp_mem, b.n_streams, slot, vec_ok);
    return cudaGetLastError();
}

}  // namespace nnb
