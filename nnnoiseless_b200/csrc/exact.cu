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
// K1: high-pass biquad, one lane per stream (src/util.rs:95-107: f64 arithmetic, f32 state, 480 dependent
// steps per frame).  The recurrence cannot be re-associated, so throughput comes from running many of them
// side by side: 64 streams per block (one per thread; 1024 blocks at 65,536 streams = 6.9 per SM -- with 128 the single
// wave left SMs with 3 or 4 blocks, a 15 % imbalance).  The [64][480] input is streamed
// in six 80-sample chunks staged through shared memory, so global traffic is coalesced 128-bit while each
// lane walks its own row (stride 81 words: conflict-free).
// ================================================================================================
constexpr int HP_STREAMS = 64;
constexpr int HP_THREADS = 64;
constexpr int HP_CHUNK = 80;
constexpr int HP_LD = HP_CHUNK + 1;
static_assert(FRAME_SIZE % HP_CHUNK == 0 && HP_CHUNK % 4 == 0, "chunking must tile the frame");

// TIn = float (the reference's f32-in-i16-range samples) or short (16-bit PCM, src/nnnoiseless.rs:147-177 front-end fused in)
template <typename TIn>
__global__ void __launch_bounds__(HP_THREADS) hp_filter_kernel(const TIn* __restrict__ in, long stream_stride, long sample_stride,
                                                               float* __restrict__ hist, float* __restrict__ hp_mem,
                                                               int n_streams, int slot, int vec_ok) {
    __shared__ float tile[HP_STREAMS * HP_LD];
    const int s0 = blockIdx.x * HP_STREAMS;
    const int tid = threadIdx.x;
    const int ns = min(HP_STREAMS, n_streams - s0);
    const double a0 = (double)-1.99599f, a1 = (double)0.99600f, b0 = (double)-2.0f, b1 = (double)1.0f;
    float m0 = 0.0f, m1 = 0.0f;
    if (tid < ns) {
        m0 = hp_mem[2 * (s0 + tid)];
        m1 = hp_mem[2 * (s0 + tid) + 1];
    }
    constexpr int Q = HP_CHUNK / 4;  // float4 per row per chunk
    for (int c = 0; c < FRAME_SIZE / HP_CHUNK; c++) {
        if (vec_ok && sizeof(TIn) == 4) {
            for (int idx = tid; idx < ns * Q; idx += HP_THREADS) {
                const int row = idx / Q, q = idx - row * Q;
                const float4 v = __ldg(reinterpret_cast<const float4*>(in + (long)(s0 + row) * stream_stride + c * HP_CHUNK) + q);
                float* t = tile + row * HP_LD + 4 * q;
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            }
        } else if (vec_ok) {  // 16-bit PCM: 8 samples per 128-bit load
            constexpr int Q8 = HP_CHUNK / 8;
            for (int idx = tid; idx < ns * Q8; idx += HP_THREADS) {
                const int row = idx / Q8, q = idx - row * Q8;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (long)(s0 + row) * stream_stride + c * HP_CHUNK) + q);
                float* t = tile + row * HP_LD + 8 * q;
                const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    t[2 * k] = (float)(short)(w[k] & 0xffffu);
                    t[2 * k + 1] = (float)(short)(w[k] >> 16);
                }
            }
        } else if (sample_stride == 1) {
            for (int idx = tid; idx < ns * HP_CHUNK; idx += HP_THREADS) {
                const int row = idx / HP_CHUNK, i = idx - row * HP_CHUNK;
                tile[row * HP_LD + i] = (float)in[(long)(s0 + row) * stream_stride + c * HP_CHUNK + i];
            }
        } else {  // interleaved channels: neighbouring threads take neighbouring streams of the same sample
            for (int idx = tid; idx < ns * HP_CHUNK; idx += HP_THREADS) {
                const int i = idx / ns, row = idx - i * ns;
                tile[row * HP_LD + i] = (float)in[(long)(s0 + row) * stream_stride + (long)(c * HP_CHUNK + i) * sample_stride];
            }
        }
        __syncthreads();
        if (tid < ns) {
            float* row = tile + tid * HP_LD;
#pragma unroll 4
            for (int i = 0; i < HP_CHUNK; i++) {
                const double x64 = (double)row[i];
                const double y64 = __dadd_rn(x64, (double)m0);
                const double t0 = __dsub_rn(__dmul_rn(b0, x64), __dmul_rn(a0, y64));
                const double t1 = __dsub_rn(__dmul_rn(b1, x64), __dmul_rn(a1, y64));
                m0 = __double2float_rn(__dadd_rn((double)m1, t0));
                m1 = __double2float_rn(t1);
                row[i] = __double2float_rn(y64);
            }
        }
        __syncthreads();
        // hist rows are 16-byte aligned (HIST_CAP*4, slot*480*4 and c*80*4 are multiples of 16)
        for (int idx = tid; idx < ns * Q; idx += HP_THREADS) {
            const int row = idx / Q, q = idx - row * Q;
            const float* t = tile + row * HP_LD + 4 * q;
            reinterpret_cast<float4*>(hist + (size_t)(s0 + row) * HIST_CAP + slot * FRAME_SIZE + c * HP_CHUNK)[q] =
                make_float4(t[0], t[1], t[2], t[3]);
        }
        __syncthreads();
    }
    if (tid < ns) {
        hp_mem[2 * (s0 + tid)] = m0;
        hp_mem[2 * (s0 + tid) + 1] = m1;
    }
}

cudaError_t launch_hp_filter(const BatchBuffers& b, const void* in, bool pcm16, long stream_stride, long sample_stride, int slot,
                             cudaStream_t st) {
    const int per16 = pcm16 ? 8 : 4;  // elements per 128-bit load
    int vec_ok = sample_stride == 1 && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (stream_stride % per16 == 0);
    int grid = (b.n_streams + HP_STREAMS - 1) / HP_STREAMS;
    if (pcm16)
        hp_filter_kernel<short><<<grid, HP_THREADS, 0, st>>>(static_cast<const short*>(in), stream_stride, sample_stride, b.hist, b.hp_mem,
                                                             b.n_streams, slot, vec_ok);
    else
        hp_filter_kernel<float><<<grid, HP_THREADS, 0, st>>>(static_cast<const float*>(in), stream_stride, sample_stride, b.hist, b.hp_mem,
                                                             b.n_streams, slot, vec_ok);
    return cudaGetLastError();
}

}  // namespace nnb
