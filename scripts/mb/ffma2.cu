#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0001f, c = 0.5f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k_ffma2(float* out, int iters) {
    unsigned long long a[4];
    for (int k = 0; k < 4; ++k) { float2 v = make_float2(threadIdx.x + 2 * k, threadIdx.x + 2 * k + 1); a[k] = *reinterpret_cast<unsigned long long*>(&v); }
    float2 mv = make_float2(1.0001f, 1.0001f), cv = make_float2(0.5f, 0.5f);
    unsigned long long m = *reinterpret_cast<unsigned long long*>(&mv), c = *reinterpret_cast<unsigned long long*>(&cv);
    for (int i = 0; i < iters; ++i) {
        fma2(a[0], m, c); fma2(a[1], m, c); fma2(a[2], m, c); fma2(a[3], m, c);
    }
    float s = 0;
    for (int k = 0; k < 4; ++k) { float2 v = *reinterpret_cast<float2*>(&a[k]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_ffma<<<148 * 2, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms, 148.0 * 2 * 1024 * iters * 8 * 2 / ms / 1e9);
        cudaEventRecord(e0); k_ffma2<<<148 * 2, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, 148.0 * 2 * 1024 * iters * 8 * 2 / ms / 1e9);
    }
    return 0;
}
