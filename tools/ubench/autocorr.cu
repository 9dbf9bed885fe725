// micro-benchmark: the lane-per-stream autocorrelation chain of pitch.cu in isolation
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
constexpr int PB = 864, P_LD = 868;
template <int K>
__device__ __forceinline__ float autocorr_lag(const float4* __restrict__ row) {
    float c = 0.0f;
    float4 w0 = row[0];
#pragma unroll 5
    for (int m = 0; m < (PB - 4) / 4; m++) {
        const float4 w1 = row[m + 1];
        const float e[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int d = 0; d < 4; d++) c = fa(c, fm(e[d], e[d + K]));
        w0 = w1;
    }
    return c;
}
__global__ void k(float* out, long long* cyc, int nactive) {
    extern __shared__ __align__(16) float P[];
    for (int i = threadIdx.x; i < 16 * P_LD; i += blockDim.x) P[i] = (i * 37 % 101) * 0.01f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    float v = 0;
    if (warp < nactive) {
        const float4* row = reinterpret_cast<const float4*>(P + (lane % 16) * P_LD);
        switch (warp % 5) {
            case 0: v = autocorr_lag<0>(row); break;
            case 1: v = autocorr_lag<1>(row); break;
            case 2: v = autocorr_lag<2>(row); break;
            case 3: v = autocorr_lag<3>(row); break;
            default: v = autocorr_lag<4>(row); break;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; long long h[2];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * P_LD * 4);
    for (int blocks : {1, 296, 592})
        for (int na : {1, 5, 8}) {
            k<<<blocks, 256, 16 * P_LD * 4>>>(out, cyc, na); cudaDeviceSynchronize();
            k<<<blocks, 256, 16 * P_LD * 4>>>(out, cyc, na); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
            printf("blocks %d (one SM each), active warps %d of 8: %lld cycles for 215 iterations (%.1f per iteration)\n", blocks, na, h[0], h[0] / 215.0);
        }
    return 0;
}
