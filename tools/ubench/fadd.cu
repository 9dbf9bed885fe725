// micro-benchmark: dependent FADD/FMUL chain latency and per-SM issue throughput on the target GPU
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void chain(float* out, long long* cyc, int n, float a) {
    float acc[CH];
    for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = __fadd_rn(acc[c], a);
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void chain_mul_add(float* out, long long* cyc, int n, const float* __restrict__ xs) {
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
__global__ void lds_chain(float* out, long long* cyc, int n) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i & 63;
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) idx = (int)sm[idx] + (threadIdx.x & 31);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float *out, *xs; long long* cyc;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8192); cudaMalloc(&xs, 4096); cudaMemset(xs, 0, 4096);
    long long h[4];
    const int n = 4096;
    auto run = [&](const char* name, auto kern, int threads, int ch) {
        kern<<<1, threads>>>(out, cyc, n, 1.0f); cudaDeviceSynchronize();
        kern<<<1, threads>>>(out, cyc, n, 1.0f); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s threads %4d: %.2f cycles per FADD-per-chain step, %.2f warp-instr/cycle/SM\n", name, threads, (double)h[0] / n,
               (double)n * ch * (threads / 32) / h[0]);
    };
    run("fadd 1 chain", chain<1>, 32, 1);
    run("fadd 4 chains", chain<4>, 32, 4);
    run("fadd 8 chains", chain<8>, 32, 8);
    run("fadd 8 chains 4 warps", chain<8>, 128, 8);
    run("fadd 8 chains 8 warps", chain<8>, 256, 8);
    run("fadd 8 chains 16 warps", chain<8>, 512, 8);
    run("fadd 8 chains 32 warps", chain<8>, 1024, 8);
    auto run2 = [&](const char* name, auto kern, int threads, int ch) {
        kern<<<1, threads>>>(out, cyc, n, xs); cudaDeviceSynchronize();
        kern<<<1, threads>>>(out, cyc, n, xs); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s threads %4d: %.2f cycles per step, %.2f FP warp-instr/cycle/SM\n", name, threads, (double)h[0] / n,
               (double)n * ch * 2 * (threads / 32) / h[0]);
    };
    run2("fmul+fadd 1 chain", chain_mul_add<1>, 32, 1);
    run2("fmul+fadd 4 chains", chain_mul_add<4>, 32, 4);
    run2("fmul+fadd 4 chains 16 warps", chain_mul_add<4>, 512, 4);
    run2("fmul+fadd 8 chains 32 warps", chain_mul_add<8>, 1024, 8);
    lds_chain<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDS (+cvt+add) chain: %.2f cycles per step\n", (double)h[0] / n);
    return 0;
}
