// FFMA2 / FFMA issue-rate probe: independent accumulators, all SMs, many warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float a)
{
    unsigned long long acc[16];
    float facc[32];
    for (int i = 0; i < 16; ++i) { float2 v = make_float2(threadIdx.x + i, i); acc[i] = *reinterpret_cast<unsigned long long *>(&v); }
    for (int i = 0; i < 32; ++i) facc[i] = threadIdx.x + i;
    float2 w = make_float2(a, a), x = make_float2(a * 0.5f, a * 0.25f);
    unsigned long long W = *reinterpret_cast<unsigned long long *>(&w), X = *reinterpret_cast<unsigned long long *>(&x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(W), "l"(X));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(a), "f"(x.x));
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) { float2 v = *reinterpret_cast<float2 *>(&acc[i]); s += v.x + v.y; }
    for (int i = 0; i < 32; ++i) s += facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 32; warps *= 2) {
            const int iters = 20000;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.0001f); else k<1><<<148, warps * 32>>>(out, iters, 1.0001f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double inst = (double)148 * warps * iters * (mode == 0 ? 16 : 32);
            double per_smsp_clk = inst / (148.0 * 4) / (ms * 1e-3 * clk * 1e3);
            printf("%s warps/SM %2d: %.3f ms, %.3f warp-inst/clk/SMSP (nominal clock %d kHz), %.1f TFLOP/s\n", mode == 0 ? "FFMA2" : "FFMA ", warps, ms,
                   per_smsp_clk, clk, inst * 32 * (mode == 0 ? 4 : 2) / (ms * 1e-3) / 1e12);
        }
    return 0;
}
