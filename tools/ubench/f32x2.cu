// micro-benchmark: packed f32x2 (FADD2 / FMUL2, Blackwell) vs scalar FADD / FMUL: dependent-chain latency and
// per-SM throughput, non-FMA (mul then add, each IEEE-rounded) as the order-exact pitch path needs.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// NOTE: ptxas 12.9 merges mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 (fused, despite .rn): "fused" below is that.
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// unfused packed multiply: fma(a, b, -0.0) with the -0.0 pair opaque to ptxas (constant memory)
__constant__ u64 c_nz = 0x8000000080000000ull;
__device__ __forceinline__ u64 mul2x(u64 a, u64 b) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c_nz)); return r; }

// CH independent chains of acc = acc + x*y  (scalar: CH floats;  packed: CH pairs = 2*CH streams)
template <int CH>
__global__ void scalar_k(float* out, long long* cyc, int n, const float* __restrict__ xs) {
    float acc[CH];
    for (int c = 0; c < CH; c++) acc[c] = 0;
    float x0 = xs[threadIdx.x], x1 = xs[threadIdx.x + 32];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = __fadd_rn(acc[c], __fmul_rn(x0, x1 + c));
        x0 += 1.0f;
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void packed_k(float* out, long long* cyc, int n, const float* __restrict__ xs) {
    u64 acc[CH], y[CH];
    for (int c = 0; c < CH; c++) { acc[c] = pk(0.f, 0.f); y[c] = pk(xs[threadIdx.x + 32] + c, xs[threadIdx.x + 64] + c); }
    u64 x = pk(xs[threadIdx.x], xs[threadIdx.x + 96]);
    const u64 one = pk(1.0f, 1.0f);
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = add2(acc[c], mul2(x, y[c]));
        x = add2(x, one);
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; c++) s += lo(acc[c]) + hi(acc[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void packed_unfused_k(float* out, long long* cyc, int n, const float* __restrict__ xs) {
    u64 acc[CH], y[CH];
    for (int c = 0; c < CH; c++) { acc[c] = pk(0.f, 0.f); y[c] = pk(xs[threadIdx.x + 32] + c, xs[threadIdx.x + 64] + c); }
    u64 x = pk(xs[threadIdx.x], xs[threadIdx.x + 96]);
    const u64 one = pk(1.0f, 1.0f);
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = add2(acc[c], mul2x(x, y[c]));
        x = add2(x, one);
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; c++) s += lo(acc[c]) + hi(acc[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// pure dependent add chain
__global__ void add_chain_scalar(float* out, long long* cyc, int n, const float* __restrict__ xs) {
    float acc = xs[threadIdx.x], a = xs[threadIdx.x + 32];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) acc = __fadd_rn(acc, a);
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void add_chain_packed(float* out, long long* cyc, int n, const float* __restrict__ xs) {
    u64 acc = pk(xs[threadIdx.x], xs[threadIdx.x + 64]), a = pk(xs[threadIdx.x + 32], xs[threadIdx.x + 96]);
    long long t0 = clock64();
    for (int i = 0; i < n; i++) acc = add2(acc, a);
    long long t1 = clock64();
    out[threadIdx.x] = lo(acc) + hi(acc);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *out, *xs; long long* cyc;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8192); cudaMalloc(&xs, 4096); cudaMemset(xs, 0, 4096);
    long long h[2];
    const int n = 4096;
    add_chain_scalar<<<1, 32>>>(out, cyc, n, xs); cudaDeviceSynchronize();
    add_chain_scalar<<<1, 32>>>(out, cyc, n, xs); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent FADD  chain: %.2f cycles/step\n", (double)h[0] / n);
    add_chain_packed<<<1, 32>>>(out, cyc, n, xs); cudaDeviceSynchronize();
    add_chain_packed<<<1, 32>>>(out, cyc, n, xs); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent FADD2 chain: %.2f cycles/step\n", (double)h[0] / n);
    auto run = [&](const char* name, auto kern, int threads, int ch, int per) {
        kern<<<1, threads>>>(out, cyc, n, xs); cudaDeviceSynchronize();
        kern<<<1, threads>>>(out, cyc, n, xs); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        // MACs per cycle per SM (scalar MAC = one mul + one add of one float)
        printf("%-22s warps %2d chains %d: %7.2f cycles/step, %6.1f lane-MACs/cycle/SM\n", name, threads / 32, ch, (double)h[0] / n,
               (double)n * ch * per * threads / h[0]);
    };
    for (int w : {1, 2, 4, 8, 16, 32}) {
        run("scalar fmul+fadd", scalar_k<1>, 32 * w, 1, 1);
        run("packed FUSED ffma2", packed_k<1>, 32 * w, 1, 2);
        run("scalar fmul+fadd", scalar_k<4>, 32 * w, 4, 1);
        run("packed FUSED ffma2", packed_k<2>, 32 * w, 2, 2);
        run("packed FUSED ffma2", packed_k<4>, 32 * w, 4, 2);
        run("packed ffma2(-0)+fadd2", packed_unfused_k<2>, 32 * w, 2, 2);
        run("packed ffma2(-0)+fadd2", packed_unfused_k<4>, 32 * w, 4, 2);
        run("scalar fmul+fadd", scalar_k<8>, 32 * w, 8, 1);
    }
    return 0;
}
