// Microbenchmark: the softmax inner loop's instruction mix (FFMA, MUFU.EX2, FADD, F2FP per pair) as a function of
// resident warps per SM sub-partition, with a fraction of the exponentials moved to an FMA-pipe polynomial.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_mix softmax_mix.cu && ./softmax_mix
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2poly(float x) {     // x <= 0; Cody-Waite with the round-to-nearest magic constant, degree 3
    x = fmaxf(x, -126.0f);
    float t = x + 12582912.0f;
    float fl = t - 12582912.0f;
    float f = x - fl;
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}

// POLY = number of polynomial exponentials per 8 elements
template <int POLY>
__global__ void k(float* out, int iters, float scale, float negm) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = -1.0f - 0.01f * ((threadIdx.x + i) & 63);
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float a = fmaf(s[2 * j], scale, negm), b = fmaf(s[2 * j + 1], scale, negm);
            const float p0 = ((2 * j) % 8 < POLY) ? ex2poly(a) : ex2f(a);
            const float p1 = ((2 * j + 1) % 8 < POLY) ? ex2poly(b) : ex2f(b);
            sum0 += p0; sum1 += p1;
            acc ^= pack(p0, p1);
        }
        negm += 1e-7f;
    }
    if (sum0 + sum1 == 123.456f || acc == 0x12345678u) out[0] = sum0 + sum1 + acc;
}

template <int POLY>
void run(int warps_per_smsp) {
    float* d; cudaMalloc(&d, 4);
    const int sms = 148, iters = 2000;
    const int threads = warps_per_smsp * 4 * 32;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<POLY><<<sms, threads>>>(d, 10, 0.18f, 0.3f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<POLY><<<sms, threads>>>(d, iters, 0.18f, 0.3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double elems = double(sms) * threads * iters * 64;
    printf("poly %d/8  warps/SMSP %d : %7.3f ms  %6.2f elem/clk/SM @1.9GHz\n", POLY, warps_per_smsp, ms, elems / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    for (int w : {1, 2, 4, 8}) run<0>(w);
    for (int w : {1, 2, 4}) run<1>(w);
    for (int w : {1, 2, 4}) run<2>(w);
    for (int w : {1, 2, 4}) run<3>(w);
    for (int w : {1, 2, 4}) run<4>(w);
    return 0;
}
