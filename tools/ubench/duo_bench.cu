// Stand-alone timing harness for the frame-duo kernel variants (tuning tool, not product):
// C2 shape (1000 x 40000, nperseg 512, hop 128), CUDA events, prints ms per variant and a
// checksum so variants can be compared.   Build: see tools/ubench/build.sh
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "b2s_dispatch.hpp"

using namespace b2s;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Ctx {
    StftArgs a;
    float2* tw;
    int sms;
};

template <typename K>
void run(const char* name, K kern, int nt, size_t smem, int fpc, Ctx& c, int iters = 20) {
    if (const char* ex = getenv("EXTRA_SMEM")) smem += (size_t)atoi(ex);      // lowers the occupancy
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem));
    StftParams p{};
    std::string err;
    const long long resident = (long long)c.sms * occ;
    const bool dynamic = getenv("DYN") && atoi(getenv("DYN"));
    if (plan_stft(c.a, fpc, resident * fpc, p, err, dynamic) < 0) { printf("plan: %s\n", err.c_str()); exit(1); }
    if (dynamic) {
        static int* work = nullptr;
        if (!work) { CK(cudaMalloc(&work, 8)); CK(cudaMemset(work, 0, 8)); }
        p.work = work;
    }
    if (const char* cf = getenv("CHUNK")) {
        p.chunk_frames = atoi(cf);
        p.units_per_signal = (c.a.nframes + p.chunk_frames - 1) / p.chunk_frames;
        p.n_units = p.units_per_signal * c.a.batch;
    }
    p.tw = c.tw;
    const long long need = (p.n_units + fpc - 1) / fpc;
    const unsigned grid = (unsigned)(need < resident ? need : resident);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaMemset(c.a.out, 0, sizeof(float) * c.a.batch * c.a.out_batch_stride));
    for (int i = 0; i < 3; ++i) kern<<<grid, nt, smem>>>(p);
    CK(cudaDeviceSynchronize());
    float best = 1e9f, tot = 0.f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(e0));
        kern<<<grid, nt, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
        tot += ms;
    }
    CK(cudaGetLastError());
    // checksum over a sample of the output
    std::vector<float> h(1 << 20);
    CK(cudaMemcpy(h.data(), c.a.out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double cs = 0;
    for (float v : h) cs += v;
    const double bytes = 4.0 * c.a.batch * c.a.n + 4.0 * c.a.batch * c.a.nframes * (c.a.nperseg / 2 + 1);
    printf("%-34s occ %d grid %4u chunk %2d  mean %.4f ms  min %.4f ms  %.1f GB/s  frac %.3f  checksum %.9g\n", name, occ, grid,
           p.chunk_frames, tot / iters, best, bytes / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9 / 6537.0, cs);
}

template <int LOG2N>
int main_cta(int B, int N, int HOP) {
    const int NP = 1 << LOG2N;
    const int F = (N - NP) / HOP + 1, K = NP / 2 + 1;
    Ctx c;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    c.sms = prop.multiProcessorCount;
    std::vector<float> hx((size_t)B * N), hw(NP), htw;
    unsigned s = 12345u;
    for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f + 0.25f; }
    for (int i = 0; i < NP; ++i) hw[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265358979f * i / NP);
    make_tables(NP, htw);
    float *dx, *dw, *dout;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMalloc(&dw, NP * 4));
    CK(cudaMalloc(&dout, (size_t)B * F * K * 4));
    CK(cudaMalloc(&c.tw, htw.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), NP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.tw, htw.data(), htw.size() * 4, cudaMemcpyHostToDevice));
    c.a = StftArgs{dx, 0, B, N, N, NP, HOP, dw, 1, 1.0 / (20000.0 * 192.0), 0, 0.f, 0, NP / 2, 0, F, dout, (long long)F * K, 0};
    printf("nperseg %d hop %d batch %d n %d\n", NP, HOP, B, N);
    using DP = DuoCtaPlan<LOG2N>;
    if constexpr (LOG2N == 10) {
        using WP = WarpPlan<10>;
        if (HOP == 256) run("warp<10,float,4,0>", stft_psd_warp_kernel<10, float, 4, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        else run("warp<10,float,0,0>", stft_psd_warp_kernel<10, float, 0, 0>, WP::NT, WP::SMEM, WP::FPC, c);
    } else {
        using PL = Plan<LOG2N>;
        run("cta<LOG2N,float,2,0>", stft_psd_kernel<LOG2N, float, 2, 0>, PL::NT, PL::SMEM, PL::FPC, c);
    }
    run("duo_cta<LOG2N,float,0>", stft_psd_duo_cta_kernel<LOG2N, float, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    using D4 = Duo4Plan<LOG2N>;
    if (HOP * 16 == NP * 4) run("duo4<LOG2N,float,4,0>", stft_psd_duo4_kernel<LOG2N, float, 4, 0>, D4::NT, D4::SMEM, D4::FPC, c);
    if (HOP * 16 == NP * 2) run("duo4<LOG2N,float,2,0>", stft_psd_duo4_kernel<LOG2N, float, 2, 0>, D4::NT, D4::SMEM, D4::FPC, c);
    if (HOP * 16 == NP * 8) run("duo4<LOG2N,float,8,0>", stft_psd_duo4_kernel<LOG2N, float, 8, 0>, D4::NT, D4::SMEM, D4::FPC, c);
    if (HOP * 16 == NP * 14) run("duo4<LOG2N,float,14,0>", stft_psd_duo4_kernel<LOG2N, float, 14, 0>, D4::NT, D4::SMEM, D4::FPC, c);
    if (HOP * 16 == NP * 16) run("duo4<LOG2N,float,16,0>", stft_psd_duo4_kernel<LOG2N, float, 16, 0>, D4::NT, D4::SMEM, D4::FPC, c);
    return 0;
}

int main_256(int B, int N, int HOP) {
    const int NP = 256;
    const int F = (N - NP) / HOP + 1, K = NP / 2 + 1;
    Ctx c;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    c.sms = prop.multiProcessorCount;
    std::vector<float> hx((size_t)B * N), hw(NP), htw;
    unsigned s = 12345u;
    for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f + 0.25f; }
    for (int i = 0; i < NP; ++i) hw[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265358979f * i / NP);
    make_tables(NP, htw);
    float *dx, *dw, *dout;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMalloc(&dw, NP * 4));
    CK(cudaMalloc(&dout, (size_t)B * F * K * 4));
    CK(cudaMalloc(&c.tw, htw.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), NP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.tw, htw.data(), htw.size() * 4, cudaMemcpyHostToDevice));
    c.a = StftArgs{dx, 0, B, N, N, NP, HOP, dw, 1, 1.0 / (20000.0 * 96.0), 0, 0.f, 0, NP / 2, 0, F, dout, (long long)F * K, 0};
    printf("nperseg %d hop %d batch %d n %d\n", NP, HOP, B, N);
    using WP = WarpPlan<8>;
    using DP = Duo256Plan;
    if (HOP == 64) {
        run("warp<8,float,4,0>", stft_psd_warp_kernel<8, float, 4, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo256<float,4,0>", stft_psd_duo256_kernel<float, 4, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 224) {
        run("warp<8,float,14,0>", stft_psd_warp_kernel<8, float, 14, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo256<float,14,0>", stft_psd_duo256_kernel<float, 14, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 256) {
        run("warp<8,float,0,0>", stft_psd_warp_kernel<8, float, 0, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo256<float,16,0>", stft_psd_duo256_kernel<float, 16, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 32) {
        run("warp<8,float,2,0>", stft_psd_warp_kernel<8, float, 2, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo256<float,2,0>", stft_psd_duo256_kernel<float, 2, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else {
        run("warp<8,float,8,0>", stft_psd_warp_kernel<8, float, 8, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo256<float,8,0>", stft_psd_duo256_kernel<float, 8, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 4 && atoi(argv[1]) == 256) return main_256(atoi(argv[3]), atoi(argv[4]), atoi(argv[2]));
    if (argc > 4) {
        const int np = atoi(argv[1]), hop = atoi(argv[2]), b = atoi(argv[3]), n = atoi(argv[4]);
        if (np == 1024) return main_cta<10>(b, n, hop);
        if (np == 2048) return main_cta<11>(b, n, hop);
        if (np == 4096) return main_cta<12>(b, n, hop);
        return 1;
    }
    const int B = 1000, N = (argc > 2) ? atoi(argv[2]) : 40000, NP = 512, HOP = (argc > 1) ? atoi(argv[1]) : 128;
    const int F = (N - NP) / HOP + 1, K = NP / 2 + 1;
    Ctx c;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    c.sms = prop.multiProcessorCount;
    std::vector<float> hx((size_t)B * N), hw(NP), htw;
    unsigned s = 12345u;
    for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f + 0.25f; }
    for (int i = 0; i < NP; ++i) hw[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265358979f * i / NP);
    make_tables(NP, htw);
    float *dx, *dw, *dout;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMalloc(&dw, NP * 4));
    CK(cudaMalloc(&dout, (size_t)B * F * K * 4));
    CK(cudaMalloc(&c.tw, htw.size() * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), NP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.tw, htw.data(), htw.size() * 4, cudaMemcpyHostToDevice));
    c.a = StftArgs{dx, 0, B, N, N, NP, HOP, dw, 1, 1.0 / (20000.0 * 192.0), 0, 0.f, 0, NP / 2, 0, F, dout, (long long)F * K, 0};
    using WP = WarpPlan<9>;
    using DP = DuoPlan;
    if (HOP == 128) {
        run("warp<9,float,4,0>", stft_psd_warp_kernel<9, float, 4, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo<float,4,0>", stft_psd_duo_kernel<float, 4, 0, 3, 0>, DP::NT, DP::SMEM, DP::FPC, c);
        if (getenv("MINB2")) run("duo<float,4,0,MINB=2>", stft_psd_duo_kernel<float, 4, 0, 2, 0>, DP::NT, DP::SMEM, DP::FPC, c);
        run("duo OPT=8  no STG", stft_psd_duo_kernel<float, 4, 0, 3, 8>, DP::NT, DP::SMEM, DP::FPC, c);
        run("duo OPT=16 no exchange", stft_psd_duo_kernel<float, 4, 0, 3, 16>, DP::NT, DP::SMEM, DP::FPC, c);
        run("duo OPT=32 no mirror shfl", stft_psd_duo_kernel<float, 4, 0, 3, 32>, DP::NT, DP::SMEM, DP::FPC, c);
        run("duo OPT=56 none of the three", stft_psd_duo_kernel<float, 4, 0, 3, 56>, DP::NT, DP::SMEM, DP::FPC, c);
        c.a.detrend = 0;
        run("duo OPT=0 detrend off", stft_psd_duo_kernel<float, 4, 0, 3, 0>, DP::NT, DP::SMEM, DP::FPC, c);
        run("duo OPT=56 detrend off", stft_psd_duo_kernel<float, 4, 0, 3, 56>, DP::NT, DP::SMEM, DP::FPC, c);
        c.a.detrend = 1;
    } else if (HOP == 64) {
        run("warp<9,float,2,0>", stft_psd_warp_kernel<9, float, 2, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo<float,2,0,MINB=3>", stft_psd_duo_kernel<float, 2, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 448) {
        run("warp<9,float,14,0>", stft_psd_warp_kernel<9, float, 14, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo<float,14,0,MINB=3>", stft_psd_duo_kernel<float, 14, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 512) {
        run("warp<9,float,0,0>", stft_psd_warp_kernel<9, float, 0, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo<float,16,0,MINB=3>", stft_psd_duo_kernel<float, 16, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    } else if (HOP == 256) {
        run("warp<9,float,8,0>", stft_psd_warp_kernel<9, float, 8, 0>, WP::NT, WP::SMEM, WP::FPC, c);
        run("duo<float,8,0,MINB=3>", stft_psd_duo_kernel<float, 8, 0>, DP::NT, DP::SMEM, DP::FPC, c);
    }
    return 0;
}
