// FP32x2 latency / throughput vs ILP and warps per SM (tuning tool).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int ILP, int KIND>
__global__ void k(float* out, float a, long long* cyc) {
    unsigned long long v[ILP];
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(threadIdx.x * 0.001f + i), "f"(1.0f + i));
    float s1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s1[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(aa));
            else if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[i]) : "l"(aa));
            else s1[i] = fmaf(s1[i], a, a);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[i])); s += x + y + s1[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int KIND>
void run(int warps_per_sm) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<ILP, KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, cyc);
    k<ILP, KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (ITERS * ILP);           // cycles per instruction per warp
    double smsp_rate = (warps_per_sm / 4.0) / per;    // instr / cycle / SMSP
    printf("%s ILP %d warps/SM %2d: %.2f cyc/instr/warp, %.3f instr/cyc/SMSP\n", KIND == 0 ? "FADD2" : (KIND == 1 ? "FFMA2" : "FFMA "), ILP, warps_per_sm, per, smsp_rate);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>(4); run<2, 0>(4); run<4, 0>(4); run<8, 0>(4);
    run<1, 0>(12); run<2, 0>(12); run<4, 0>(12); run<8, 0>(12);
    run<1, 0>(16); run<2, 0>(16); run<4, 0>(16);
    run<1, 1>(4); run<2, 1>(4); run<4, 1>(4); run<8, 1>(4);
    run<1, 1>(12); run<2, 1>(12); run<4, 1>(12);
    run<1, 2>(4); run<2, 2>(4); run<4, 2>(4); run<8, 2>(4);
    run<1, 2>(12); run<2, 2>(12); run<4, 2>(12);
    return 0;
}
