// Micro-benchmarks of B200 issue/pipe rates that decide the STFT kernel design:
// scalar FFMA vs packed fma.rn.f32x2, FADD vs add.f32x2, SHFL and LDS.64 throughput.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0) {            // scalar FFMA, 16 independent chains
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {     // packed fma.rn.f32x2, 8 independent chains (same flops)
        unsigned long long aa, bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        unsigned long long v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v[i]));
    } else if (MODE == 2) {     // scalar FADD
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + a;
    } else if (MODE == 3) {     // packed add.f32x2
        unsigned long long aa;
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        unsigned long long v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(aa));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v[i]));
    } else if (MODE == 4) {     // SHFL.BFLY
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1);
    } else if (MODE == 5) {     // LDS.64 conflict-free
        __shared__ float2 s[256 * 4];
        s[threadIdx.x] = make_float2(a, b);
        __syncthreads();
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float tx, ty;
                unsigned addr = (unsigned)__cvta_generic_to_shared(&s[(threadIdx.x + i * 32 + (int)x[2 * i + 1]) & 1023]);
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(tx), "=f"(ty) : "r"(addr));
                x[2 * i] += tx + ty * 1e-9f;
            }
    } else if (MODE == 6) {     // mixed: 8 FFMA2 + 8 FADD (does packing free issue slots?)
        unsigned long long aa, bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        unsigned long long v[4];
        float y[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = x[8 + i];
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) x[8 + i] = y[i];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_thread_iter) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double thread_ops = 148.0 * 8 * 256 * ITERS * ops_per_thread_iter;
    // lane-ops per clock per SM assuming 1.9 GHz
    printf("%-28s %8.3f ms  %8.1f Gop/s  %6.1f lane-ops/clk/SM@1.9GHz\n", name, ms, thread_ops / ms / 1e6,
           thread_ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main() {
    run<0>("FFMA scalar (16/iter)", 16);
    run<1>("FFMA2 packed (8 instr=16 fma)", 16);
    run<2>("FADD scalar", 16);
    run<3>("FADD2 packed", 16);
    run<4>("SHFL.BFLY", 16);
    run<5>("LDS.64", 8);
    run<6>("4 FFMA2 + 8 FFMA (16 fma)", 16);
    return 0;
}
