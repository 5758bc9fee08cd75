// Microbenchmark: exp2 throughput per SM on sm_100a for the variants the attention softmax could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp2_tput exp2_tput.cu && ./exp2_tput
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// Cody-Waite + degree-3 polynomial on the FMA pipe (x <= 0, clamped at -126)
__device__ __forceinline__ float ex2poly(float x) {
    x = fmaxf(x, -126.0f);
    float fl = floorf(x);            // FRND
    float f = x - fl;
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    int e = static_cast<int>(fl);
    return __int_as_float(__float_as_int(p) + (e << 23));
}
// magic-number variant: no FRND / F2I
__device__ __forceinline__ float ex2poly2(float x) {
    x = fmaxf(x, -126.0f);
    float t = x + 12582912.0f;                   // 1.5 * 2^23: integer part lands in the low mantissa bits
    float fl = t - 12582912.0f;
    float f = x - fl;                            // in [-0.5, 0.5]
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + i) * 1e-6f - 1.0f;
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0xB800B800u + threadIdx.x + i;          // (-0.5, -0.5) f16
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2f(v[i]) - 1.5f;
            if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
            if (MODE == 2) h[i] = ex2b2(h[i]) ^ 0x80008000u;
            if (MODE == 3) v[i] = ex2poly(v[i]) - 1.5f;
            if (MODE == 4) v[i] = ex2poly2(v[i]) - 1.5f;
            if (MODE == 5) { if (i & 1) v[i] = ex2poly2(v[i]) - 1.5f; else v[i] = ex2f(v[i]) - 1.5f; }
            if (MODE == 6) { if ((i & 3) == 3) v[i] = ex2poly2(v[i]) - 1.5f; else v[i] = ex2f(v[i]) - 1.5f; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(h[i]);
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
    float* d; cudaMalloc(&d, 4);
    int sms = 148, iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms * 2, 512>>>(d, 16, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<sms * 2, 512>>>(d, iters, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double elems = double(sms) * 2 * 512 * iters * 8 * per_op;
    printf("%-28s %8.3f ms  %7.2f Gelem/s  %6.2f elem/clk/SM @1.9GHz\n", name, ms, elems / ms / 1e6, elems / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.ftz.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("poly3 (floor/f2i)", 1);
    run<4>("poly3 (magic)", 1);
    run<5>("1:1 mufu:poly", 1);
    run<6>("3:1 mufu:poly", 1);
    // accuracy of the polynomial
    return 0;
}
