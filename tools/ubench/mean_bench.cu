// Stand-alone A/B for the cross-sweep mean of config 2 (tuning tool, not product):
//   (a) stft_psd_duo_kernel + batch_sum_kernel (two slab passes), as the library runs it
//   (b) stft_psd_duo_sum_kernel (sum fused) + one fold over the sweep blocks, for several block sizes,
//       static round-robin and dynamic draws (BS=<rows> picks one block size)
// Prints ms per part and compares the per-sweep rows (must be bit-identical) and the sums.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "b2s_dispatch.hpp"
#include "b2s_duo_sum_kernel.cuh"
#include "b2s_aux_kernels.cuh"

using namespace b2s;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <typename F>
float time_ms(F&& f, int iters = 20) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

struct V { int B, F, sms; long long E; float *dout, *dout2, *dsum, *dsum2, *dacc; int* work; StftParams pa; };
template <int TM>
void variant(V& v, const std::function<void()>& stft_a, const std::function<void()>& sum_a) {
    using DP = DuoPlan;
    const int B = v.B, F = v.F, sms = v.sms; const long long E = v.E;
    float *dout = v.dout, *dout2 = v.dout2, *dsum = v.dsum, *dsum2 = v.dsum2, *dacc = v.dacc; int* work = v.work;
    const StftParams& pa = v.pa;
    const unsigned gx = (unsigned)((E + 255) / 256);
    const size_t SMB = DuoSumPlan::SMEM;
    auto kb = stft_psd_duo_sum_kernel<float, 4, TM>;
    CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMB));
    int occb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, kb, DP::NT, SMB));
    if (getenv("OCC")) occb = atoi(getenv("OCC"));        // the occupancy query answers 1 for kernels that allocate tensor memory
    const long long resb = (long long)sms * occb;
    const int nduos = (F + 1) / 2;
    std::vector<int> bss;
    if (getenv("BS")) bss.push_back(atoi(getenv("BS")));
    else {
        const int fit = (int)(resb * DP::FPC / ((nduos + 1) / 2 * 2));          // blocks that fill the machine once
        for (int w = 1; w <= 4; ++w) bss.push_back((B + fit * w - 1) / (fit * w));
        bss.push_back(8); bss.push_back(16);
    }
    for (int dynamic = 0; dynamic < 2; ++dynamic)
    for (int bs : bss) {
        if (bs < 1) continue;
        const int nblk = (B + bs - 1) / bs;
        if (nblk > 512) continue;
        StftParams pb = pa;
        pb.out = dout2;
        pb.acc = dacc;
        pb.acc_rows = bs;
        pb.acc_batch = B;
        pb.units_per_signal = (nduos + 1) / 2 * 2;
        pb.n_units = (long long)nblk * pb.units_per_signal;
        pb.chunk_frames = 2;
        pb.work = dynamic ? work + 2 : nullptr;
        const unsigned grid_b = (unsigned)std::min<long long>((pb.n_units + DP::FPC - 1) / DP::FPC, resb);
        auto stft_b = [&] { kb<<<grid_b, DP::NT, SMB>>>(pb); };
        auto fold_b = [&] { batch_sum_kernel<<<dim3(gx, 1), 256>>>(dacc, E, nblk, nblk, E, dsum2, 1.0f); };
        CK(cudaMemset(dout2, 0xff, (size_t)B * E * 4));
        const float t_b1 = time_ms(stft_b), t_b = time_ms([&] { stft_b(); fold_b(); });
        // compare
        stft_a(); sum_a();
        CK(cudaDeviceSynchronize());
        std::vector<float> h1((size_t)B * E), h2((size_t)B * E), s1(E), s2(E);
        CK(cudaMemcpy(h1.data(), dout, h1.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h2.data(), dout2, h2.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s1.data(), dsum, E * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s2.data(), dsum2, E * 4, cudaMemcpyDeviceToHost));
        const bool same = memcmp(h1.data(), h2.data(), h1.size() * 4) == 0;
        double worst = 0;
        for (long long i = 0; i < E; ++i) worst = std::max(worst, std::fabs((double)s1[i] - s2[i]) / std::fabs((double)s1[i]));
        printf("(b) tmem-acc %d %s bs %3d nblk %3d units %6lld grid %4u occ %d: fused %.4f ms  fused + fold %.4f ms  rows %s  sum rel diff %.2e\n",
               TM, dynamic ? "dyn " : "stat", bs, nblk, pb.n_units, grid_b, occb, t_b1, t_b, same ? "identical" : "DIFFER", worst);
    }
}

int main(int argc, char** argv) {
    const int B = (argc > 1) ? atoi(argv[1]) : 1000, N = 40000, NP = 512, HOP = 128;
    const int F = (N - NP) / HOP + 1, K = NP / 2 + 1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::vector<float> hx((size_t)B * N), hw(NP), htw;
    unsigned s = 12345u;
    for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f + 0.25f; }
    for (int i = 0; i < NP; ++i) hw[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265358979f * i / NP);
    make_tables(NP, htw);
    float *dx, *dw, *dout, *dout2, *dsum, *dsum2, *dscr, *dacc;
    float2* dtw;
    int* work;
    const long long E = (long long)F * K;
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMalloc(&dw, NP * 4));
    CK(cudaMalloc(&dout, (size_t)B * E * 4));
    CK(cudaMalloc(&dout2, (size_t)B * E * 4));
    CK(cudaMalloc(&dsum, E * 4));
    CK(cudaMalloc(&dsum2, E * 4));
    CK(cudaMalloc(&dscr, 64 * E * 4));
    CK(cudaMalloc(&dacc, 512 * E * 4));
    CK(cudaMalloc(&dtw, htw.size() * 4));
    CK(cudaMalloc(&work, 16));
    CK(cudaMemset(work, 0, 16));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), NP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtw, htw.data(), htw.size() * 4, cudaMemcpyHostToDevice));
    StftArgs a{dx, 0, B, N, N, NP, HOP, dw, 1, 1.0 / (20000.0 * 192.0), 0, 0.f, 0, NP / 2, 0, F, dout, E, 0};
    using DP = DuoPlan;
    std::string err;

    // ---- (a) ----
    auto ka = stft_psd_duo_kernel<float, 4, 0, 3, 0>;
    CK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP::SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka, DP::NT, DP::SMEM));
    const long long resident = (long long)sms * occ;
    StftParams pa{};
    if (plan_stft(a, DP::FPC, resident * DP::FPC, pa, err, true) < 0) { printf("plan: %s\n", err.c_str()); return 1; }
    pa.tw = dtw;
    pa.work = work;
    const unsigned grid_a = (unsigned)std::min<long long>((pa.n_units + DP::FPC - 1) / DP::FPC, resident);
    const unsigned gx = (unsigned)((E + 255) / 256);
    const int slabs = (B + 127) / 128;
    auto stft_a = [&] { ka<<<grid_a, DP::NT, DP::SMEM>>>(pa); };
    auto sum_a = [&] {
        if (slabs == 1) { batch_sum_kernel<<<dim3(gx, 1), 256>>>(dout, E, B, B, E, dsum, 1.0f); return; }
        batch_sum_kernel<<<dim3(gx, slabs), 256>>>(dout, E, B, 128, E, dscr, 1.0f);
        batch_sum_kernel<<<dim3(gx, 1), 256>>>(dscr, E, slabs, slabs, E, dsum, 1.0f);
    };
    const float t_a1 = time_ms(stft_a), t_a = time_ms([&] { stft_a(); sum_a(); });
    printf("(a) duo %.4f ms   duo + batch_sum %.4f ms   occ %d grid %u chunk %d\n", t_a1, t_a, occ, grid_a, pa.chunk_frames);

    V v{B, F, sms, E, dout, dout2, dsum, dsum2, dacc, work, pa};
    if (!getenv("ONLY_TM")) variant<0>(v, stft_a, sum_a);
    variant<1>(v, stft_a, sum_a);
    return 0;
}
