// Does operand fetch limit FP32x2 throughput?  FADD2 / FFMA2 with distinct 64-bit register
// operands (tuning tool).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024
template <int KIND>
__global__ void k(float* out, float a, long long* cyc) {
    unsigned long long v[16];
    float sc[4] = {a, a * 1.5f, a * 0.5f, a * 0.25f};
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(threadIdx.x * 0.001f + i), "f"(1.0f + i * a));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (KIND == 0) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(v[i]) : "l"(v[(i + 5) & 15]), "l"(v[(i + 9) & 15]));
            else if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v[i]) : "l"(v[(i + 5) & 15]), "l"(v[(i + 9) & 15]), "l"(v[(i + 13) & 15]));
            else if (KIND == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(v[(i + 8) & 15]));
            else if (KIND == 3) {   // FFMA2 with a broadcast scalar multiplier
                unsigned long long bb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(sc[i & 3]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v[i]) : "l"(v[(i + 5) & 15]), "l"(bb), "l"(v[(i + 13) & 15]));
            } else if (KIND == 4) { // FMUL2 by a broadcast scalar
                unsigned long long bb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(sc[i & 3]));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(v[i]) : "l"(v[(i + 5) & 15]), "l"(bb));
            } else if (KIND == 5) { // FFMA2 d = a * s + d (accumulate in place)
                unsigned long long bb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(sc[i & 3]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v[i]) : "l"(v[(i + 5) & 15]), "l"(bb));
            } else if (KIND == 6) { // scalar FFMA 3 distinct regs x2 (same flops as one FFMA2)
                float x0, y0, x1, y1, x2, y2;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(y0) : "l"(v[(i + 5) & 15]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x1), "=f"(y1) : "l"(v[(i + 9) & 15]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(y2) : "l"(v[(i + 13) & 15]));
                float r0 = fmaf(x0, x1, x2), r1 = fmaf(y0, y1, y2);
                asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(r0), "f"(r1));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int warps_per_sm) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, cyc);
    k<KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (ITERS * 16);
    printf("%-28s warps/SM %2d: %.2f cyc/instr/warp, %.3f instr/cyc/SMSP\n", name, warps_per_sm, per, (warps_per_sm / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FADD2 d = a + b (3 regs)", 4); run<0>("FADD2 d = a + b (3 regs)", 12);
    run<1>("FFMA2 d = a*b + c (4 regs)", 4); run<1>("FFMA2 d = a*b + c (4 regs)", 12);
    run<2>("FADD2 d += b", 4); run<2>("FADD2 d += b", 12);
    run<3>("FFMA2 d = a*s + c (s scalar)", 4); run<3>("FFMA2 d = a*s + c (s scalar)", 12);
    run<4>("FMUL2 d = a*s (s scalar)", 4); run<4>("FMUL2 d = a*s (s scalar)", 12);
    run<5>("FFMA2 d = a*s + d", 4); run<5>("FFMA2 d = a*s + d", 12);
    run<6>("2x FFMA scalar (3 regs each)", 4); run<6>("2x FFMA scalar (3 regs each)", 12);
    return 0;
}
