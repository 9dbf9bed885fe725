// micro-benchmark: inner_prod_480 / inner_prod_window as used by the pitch kernel, timed standalone
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
constexpr int HALF_N = 480;
__device__ __forceinline__ float inner_prod_480(const float4* __restrict__ xr, const float* __restrict__ y) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
    for (int m = 0; m < HALF_N / 4; m++) {
        const float4 x = xr[m];
        s0 = fa(s0, fm(x.x, y[4 * m]));
        s1 = fa(s1, fm(x.y, y[4 * m + 1]));
        s2 = fa(s2, fm(x.z, y[4 * m + 2]));
        s3 = fa(s3, fm(x.w, y[4 * m + 3]));
    }
    return fa(fa(fa(s0, s1), s2), s3);
}
template <int NLAG>
__device__ __forceinline__ void inner_prod_window(const float4* __restrict__ xr, const float* __restrict__ y, float* out) {
    float acc[NLAG][4];
    for (int c = 0; c < NLAG; c++) for (int u = 0; u < 4; u++) acc[c][u] = 0.0f;
    float w[8];
    for (int u = 0; u < 4; u++) w[u] = y[u];
#pragma unroll 2
    for (int m = 0; m < HALF_N / 4; m++) {
        const float4 x = xr[m];
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; u++) w[4 + u] = y[4 * m + 4 + u];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NLAG; c++) acc[c][u] = fa(acc[c][u], fm(xv[u], w[u + c]));
#pragma unroll
        for (int u = 0; u < 4; u++) w[u] = w[4 + u];
    }
    for (int c = 0; c < NLAG; c++) out[c] = fa(fa(fa(acc[c][0], acc[c][1]), acc[c][2]), acc[c][3]);
}
__global__ void k_ip(float* out, long long* cyc, int mode) {
    extern __shared__ __align__(16) float P[];  // [16][868]
    for (int i = threadIdx.x; i < 16 * 868; i += blockDim.x) P[i] = (i * 37 % 101) * 0.01f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int s = lane & 15, lag = 30 + 7 * (lane >> 0);
    const float* prow = P + s * 868;
    long long t0 = clock64();
    float r = 0;
    if (mode == 0) r = inner_prod_480(reinterpret_cast<const float4*>(prow + 384), prow + 384 - (lag % 350));
    else if (mode == 1) { float o[3]; inner_prod_window<3>(reinterpret_cast<const float4*>(prow + 384), prow + 384 - (lag % 350), o); r = o[0] + o[1] + o[2]; }
    else { float o[5]; inner_prod_window<5>(reinterpret_cast<const float4*>(prow + 384), prow + (lag % 280), o); r = o[0] + o[1] + o[2] + o[3] + o[4]; }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    cudaFuncSetAttribute(k_ip, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 868 * 4);
    const char* names[3] = {"inner_prod_480 (1 lag/lane)", "window<3>", "window<5>"};
    for (int mode = 0; mode < 3; mode++)
        for (int threads : {32, 128, 256}) {
            k_ip<<<1, threads, 16 * 868 * 4>>>(out, cyc, mode); cudaDeviceSynchronize();
            k_ip<<<1, threads, 16 * 868 * 4>>>(out, cyc, mode); cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s threads %3d: %lld cycles\n", names[mode], threads, h);
        }
    return 0;
}
