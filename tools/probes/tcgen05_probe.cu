// tcgen05_probe.cu -- stand-alone check of the tcgen05 building blocks the GRU kernel uses, against the CPU:
//   (1) SS: D[128 x N] (TMEM, f32) = A[128 x K] (smem, f16, K-major, no swizzle) * B[N x K]^T (smem, f16, K-major, no swizzle)
//   (2) TS: the same with A read from TMEM (written there with tcgen05.st, two f16 per 32-bit column)
// Descriptor fields as in cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor); canonical K-major
// no-swizzle layout in 16-byte units ((8,n),2):((1,SBO),LBO).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tcgen05_probe tools/probes/tcgen05_probe.cu && timeout 60 /tmp/tcgen05_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);            // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;  // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // version = 1 (Blackwell)
    return d;                                          // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                     // c_format = F32
    // a_format = b_format = 0 (F16), no negate, a_major = b_major = 0 (K)
    d |= (uint32_t)(n >> 3) << 17;    // n_dim
    d |= (uint32_t)(m >> 4) << 24;    // m_dim
    return d;
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(128) probe(const __half* gA, const __half* gB, float* out_ss, float* out_ts) {
    __shared__ __align__(128) __half sA[M * K];
    __shared__ __align__(128) __half sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // canonical layout: element (r, k) at (k / 8) * (R * 16 B) + r * 16 B + (k % 8) * 2 B
    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        sA[(k / 8) * (M * 8) + r * 8 + (k % 8)] = gA[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        sB[(k / 8) * (N * 8) + r * 8 + (k % 8)] = gB[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t d_ss = tmem, d_ts = tmem + 32, a_tm = tmem + 64;  // columns: [0,32) D_ss | [32,64) D_ts | [64,96) A (K=64 halves = 32 columns)
    const uint32_t idesc = make_idesc(M, N);

    // ---- (1) SS ----
    if (tid == 0) {
        for (int kk = 0; kk < K / 16; kk++) {
            const uint64_t ad = make_desc(smem_u32(sA) + kk * 2 * (M * 16), M * 16, 128);
            const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * (N * 16), N * 16, 128);
            mma_ss(d_ss, ad, bd, idesc, kk > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
        uint32_t r[32];
        const uint32_t taddr = d_ss + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
            "%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int n = 0; n < 32; n++) out_ss[tid * N + n] = __uint_as_float(r[n]);
    }

    // ---- (2) TS: A row of this thread into TMEM, two halves per column ----
    {
        for (int c0 = 0; c0 < K / 2; c0 += 8) {  // 8 columns = 16 halves per store
            uint32_t v[8];
            for (int c = 0; c < 8; c++) {
                const __half lo = gA[tid * K + 2 * (c0 + c)], hi = gA[tid * K + 2 * (c0 + c) + 1];
                v[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            const uint32_t taddr = a_tm + c0 + ((uint32_t)(warp * 32) << 16);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                         "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0) {
        for (int kk = 0; kk < K / 16; kk++) {
            const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * (N * 16), N * 16, 128);
            mma_ts(d_ts, a_tm + kk * 8, bd, idesc, kk > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
        uint32_t r[32];
        const uint32_t taddr = d_ts + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
            "%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int n = 0; n < 32; n++) out_ts[tid * N + n] = __uint_as_float(r[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> A(M * K), B(N * K);
    std::vector<float> Af(M * K), Bf(N * K);
    srand(3);
    for (int i = 0; i < M * K; i++) {
        Af[i] = (float)(rand() % 2001 - 1000) / 512.0f;
        A[i] = __float2half(Af[i]);
        Af[i] = __half2float(A[i]);
    }
    for (int i = 0; i < N * K; i++) {
        Bf[i] = (float)(rand() % 257 - 128);
        B[i] = __float2half(Bf[i]);
    }
    __half *dA, *dB;
    float *dss, *dts;
    cudaMalloc(&dA, sizeof(__half) * M * K);
    cudaMalloc(&dB, sizeof(__half) * N * K);
    cudaMalloc(&dss, sizeof(float) * M * N);
    cudaMalloc(&dts, sizeof(float) * M * N);
    cudaMemset(dss, 0xff, sizeof(float) * M * N);
    cudaMemset(dts, 0xff, sizeof(float) * M * N);
    cudaMemcpy(dA, A.data(), sizeof(__half) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), sizeof(__half) * N * K, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dss, dts);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> ss(M * N), ts(M * N);
    cudaMemcpy(ss.data(), dss, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(ts.data(), dts, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, ref_max = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += (double)Af[m * K + k] * Bf[n * K + k];
            e1 = fmax(e1, fabs(ss[m * N + n] - s));
            e2 = fmax(e2, fabs(ts[m * N + n] - s));
            ref_max = fmax(ref_max, fabs(s));
        }
    printf("ref max %.1f  SS max err %.4g  TS max err %.4g\n", ref_max, e1, e2);
    printf("sample ss[0][0..3] %g %g %g %g  ts %g %g %g %g\n", ss[0], ss[1], ss[2], ss[3], ts[0], ts[1], ts[2], ts[3]);
    return (e == cudaSuccess && e1 < 0.05 && e2 < 0.05) ? 0 : 1;
}
