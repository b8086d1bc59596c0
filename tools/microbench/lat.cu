// Measured B200 instruction costs behind the kernels' ceilings:
//  * dependent-issue latency of the instruction kinds on the prefix kernel's chain (one warp, then one warp per SM sub-partition);
//  * issue throughput per SM of MUFU.EX2 / MUFU.RCP / FFMA with every sub-partition saturated (the attention energy
//    kernel's ceiling: 4 MUFU per hypothesis-frame-channel).
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lat tools/microbench/lat.cu && build/lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_ffma_imm(float *out, long long *cyc, float e) {
    float p = e;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = fmaf(p, e, 0.0783616676926612854f);
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma_reg(float *out, long long *cyc, float e, float c) {
    float p = e;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = fmaf(p, e, c);
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fadd(float *out, long long *cyc, float e) {
    float p = e;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = __fadd_rn(p, e);
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fmnmx(float *out, long long *cyc, float e) {
    float p = e;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = fmaxf(p * 1.0001f, e);      // FMUL + FMNMX
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_mufu(float *out, long long *cyc, float e) {
    float p = e;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(p)); p = y; }
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(float *out, long long *cyc, int stride) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + stride) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = s[p];
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the polynomial log-add-exp chain itself (Horner), as the kernel runs it
__device__ __forceinline__ float lae(float a, float b) {
    const float m = fmaxf(a, b), ad = fabsf(a - b);
    float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ad * -1.4426950408889634f));
    float p = fmaf(-0.029588507f, e, 0.078361668f);
    p = fmaf(p, e, -0.1367477f); p = fmaf(p, e, 0.19111431f); p = fmaf(p, e, -0.24844369f); p = fmaf(p, e, 0.3331927f);
    p = fmaf(p, e, -0.49999502f); p = fmaf(p, e, 1.0f);
    return __fadd_rn(m, e * p);
}
__global__ void k_lae(float *out, long long *cyc, const float *a) {
    float psi = a[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) psi = lae(psi, a[(i + threadIdx.x) & 1023]);
    long long t1 = clock64();
    out[threadIdx.x] = psi; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lae_smem(float *out, long long *cyc, const float *a) {
    __shared__ float s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = a[i];
    __syncthreads();
    float psi = s[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) psi = lae(psi, s[(i + threadIdx.x) & 1023]);
    long long t1 = clock64();
    out[threadIdx.x] = psi; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---- throughput: 8 independent chains per thread, 1024 threads per CTA, one CTA per SM ----------------------------------
template <int KIND>
__global__ void __launch_bounds__(1024) k_thr(float *out, long long *cyc, float e)
{
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = e + q * 0.01f + threadIdx.x * 1e-4f;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 512; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float y;
            if (KIND == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[q]));
            else if (KIND == 1) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[q])); y += 0.75f; }     // (+ FADD: rcp(rcp(x)) alone is folded away)
            else y = fmaf(v[q], e, 0.25f);
            v[q] = y;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += v[q];
    out[blockIdx.x * 1024 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *out, *a; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&a, 4096);
    float ha[1024]; for (int i = 0; i < 1024; ++i) ha[i] = -3.0f - (i % 7) * 0.37f;
    cudaMemcpy(a, ha, 4096, cudaMemcpyHostToDevice);
    for (int threads = 32; threads <= 128; threads *= 4) {
#define RUN(name, ...) name<<<1, threads>>>(__VA_ARGS__); name<<<1, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-12s threads %3d: %.2f cycles per dependent op\n", #name, threads, (double)h / N);
        RUN(k_ffma_imm, out, cyc, 0.3f)
        RUN(k_ffma_reg, out, cyc, 0.3f, 0.0783f)
        RUN(k_fadd, out, cyc, 1e-3f)
        RUN(k_fmnmx, out, cyc, 0.3f)
        RUN(k_mufu, out, cyc, -0.5f)
        RUN(k_lds, out, cyc, 33)
        RUN(k_lae, out, cyc, a)
        RUN(k_lae_smem, out, cyc, a)
    }
    float *big; cudaMalloc(&big, 148 * 1024 * 4);
#define THR(kind, name) k_thr<kind><<<148, 1024>>>(big, cyc, 0.37f); k_thr<kind><<<148, 1024>>>(big, cyc, 0.37f); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-12s throughput: %.2f lanes per clock per SM (%.2f cycles per warp instruction per sub-partition)\n", name, 512.0 * 8 * 1024 / (double)h, (double)h * 4 / (512.0 * 8 * 32));
    THR(0, "MUFU.EX2") THR(1, "MUFU.RCP") THR(2, "FFMA")
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
