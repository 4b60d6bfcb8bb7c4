// Throughput of the packed radix-16 butterfly itself (registers only), tuning tool.
#include <cstdio>
#include <cuda_runtime.h>
#include "b2s_duo_kernel.cuh"
using namespace b2s;
#define ITERS 512
template <int KIND>
__global__ void __launch_bounds__(384, 1) k(float* out, float a, long long* cyc) {
    cpx2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i].re = cmk(threadIdx.x * 0.001f + i, 1.f + i * a);
        v[i].im = cmk(threadIdx.x * 0.002f - i, 2.f - i * a);
    }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (KIND == 0) c2radix16(v);
        if (KIND == 1) {
#pragma unroll
            for (int tt = 1; tt < 16; ++tt) v[tt] = c2mul(v[tt], cmk(a, 1.f - a));
            c2radix16(v);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i].re = pk_muls(v[i].re, 0.25f); }   // keep values bounded (16 FMUL2)
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += v[i].re.x + v[i].re.y + v[i].im.x + v[i].im.y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int ctas_per_sm) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 8 * 128 * 4); cudaMalloc(&cyc, 8);
    k<KIND><<<148, 128 * ctas_per_sm>>>(out, 0.7f, cyc);
    k<KIND><<<148, 128 * ctas_per_sm>>>(out, 0.7f, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-32s %d warps/SMSP: %.1f cycles per call per warp, %.1f per SMSP\n", name, ctas_per_sm, (double)h / ITERS, (double)h / ITERS / ctas_per_sm);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("c2radix16 (+16 FMUL2)", 1); run<0>("c2radix16 (+16 FMUL2)", 2); run<0>("c2radix16 (+16 FMUL2)", 3);
    run<1>("15 c2mul + c2radix16 (+16)", 1); run<1>("15 c2mul + c2radix16 (+16)", 3);
    return 0;
}
