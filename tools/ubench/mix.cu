// Does an FP32x2 instruction block the issue port for its second cycle?  FADD2 interleaved
// with integer ALU / shuffle / LDS instructions (tuning tool).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024
template <int KIND>
__global__ void k(float* out, float a, int ia, long long* cyc) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = a;
    __syncthreads();
    unsigned long long v[8], aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    int n[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(threadIdx.x * 0.001f + i), "f"(1.0f + i * a));
        n[i] = threadIdx.x + i;
        s[i] = threadIdx.x * 0.5f + i;
    }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND != 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(aa));
            if (KIND == 1 || KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[i]) : "r"(ia), "r"(it));
            if (KIND == 3) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+f"(s[i]));
            if (KIND == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s[i]) : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + i * 32) & 1023])));
            if (KIND == 5) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(a));
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[i])); r += x + y + n[i] + s[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND>
void run(const char* name, int warps_per_sm) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, 3, cyc);
    k<KIND><<<148, warps_per_sm * 32>>>(out, 1.0001f, 3, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM %2d: %.2f cycles per group (per SMSP, %d warps)\n", name, warps_per_sm, (double)h / (ITERS * 8) / (warps_per_sm / 4.0), warps_per_sm / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FADD2 alone", 12);
    run<1>("LOP3 alone", 12);
    run<2>("FADD2 + LOP3", 12);
    run<3>("FADD2 + SHFL", 12);
    run<4>("FADD2 + LDS.32", 12);
    run<5>("FADD2 + FADD", 12);
    return 0;
}
