// CPU SIMT-emulator harness for csrc/b2s_kernels.cuh -- TEST TOOLING ONLY.
// Build: g++ -std=c++20 -O1 -DB2S_EMU -shared -fPIC -pthread -I tests/emu \
//        -I spectrogram_generator_b200/csrc tests/emu/emu_stft.cpp -o tests/emu/libb2s_emu.so
#include "emu_cuda.hpp"
#include <cstdlib>

thread_local uint3_ threadIdx;
thread_local uint3_ blockIdx;
dim3_ blockDim;
dim3_ gridDim;
namespace emu { Cta* g_cta = nullptr; }

#include "b2s_aux_kernels.cuh"
#include "b2s_dispatch.hpp"
#include "b2s_duo_sum_kernel.cuh"

using namespace b2s;

// which kernel family the last emu_stft_psd call ran (tests assert that a shape reaches the kernel it is meant for)
static int g_last_family = 0;
extern "C" int emu_last_family() { return g_last_family; }
enum { FAM_MIXED = 11, FAM_PAIRQ = 10, FAM_PAIR = 9, FAM_DUO256 = 1, FAM_DUO4 = 2, FAM_DUO_CTA = 3, FAM_DUO = 4, FAM_WARP = 5, FAM_BIG = 6, FAM_CTA = 7, FAM_DFT = 8 };

struct EmuLauncher {
    StftParams p;
    unsigned grid;
    bool allow_duo = true;
    bool duo1024 = true;
    bool allow_duo4 = true;
    bool allow_big = true;
    bool allow_pair = true;
    bool dynamic_units = true;
    int work[2] = {0, 0};
    const StftArgs* args = nullptr;
    bool allow_pairq = false;        // opt-in (B2S_PAIRQ=1), as in the library
    template <int LOG2N, typename Tin, int MODE>
    int pairq(const StftArgs& a) {
        g_last_family = FAM_PAIRQ;
        using PP = PairQPlan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        q.ring = PP::N;
        const int nt = (PP::G <= 128) ? (PP::NT_MAX / PP::G >= 2 ? 2 * PP::G : PP::G) : PP::G;     // two groups per CTA where they fit
        if (pair_units >= 0) plan_pair_units(a, (long long)grid * (nt / PP::G), pair_units, dynamic_units, q);
        PairQConst qc;
        make_pairq_const<PP::HW>(qc);
        emu::launch(grid, nt, PP::smem_bytes((int)sizeof(Tin), nt), [&] { stft_psd_pairq_kernel<LOG2N, Tin, MODE>(q, qc); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <int LOG2N, typename Tin, int MODE>
    int pair(const StftArgs& a) {
        g_last_family = FAM_PAIR;
        using PP = PairPlan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        q.ring = PP::ring_samples(a.hop);
        if (pair_units >= 0) {
            plan_pair_units(a, (long long)grid * (PP::NT / 32), pair_units, dynamic_units, q);
        }
        int nt = PP::NT;
        if (const char* v = getenv("B2S_PAIR_NT")) nt = atoi(v) / 32 * 32;
        if (nt > PP::NT_WIDE) nt = PP::NT_WIDE;
        if (nt < 32) nt = 32;
        if (pair_units >= 0) plan_pair_units(a, (long long)grid * (nt / 32), pair_units, dynamic_units, q);
        const size_t smem = PP::smem_bytes(a.hop, (int)sizeof(Tin), nt);
        if (nt > PP::NT_MID) emu::launch(grid, nt, smem, [&] { stft_psd_pair_wide_kernel<LOG2N, Tin, MODE>(q); });
        else if (nt > PP::NT) emu::launch(grid, nt, smem, [&] { stft_psd_pair_mid_kernel<LOG2N, Tin, MODE>(q); });
        else emu::launch(grid, nt, smem, [&] { stft_psd_pair_kernel<LOG2N, Tin, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    int pair_units = -1;         // >= 0: re-plan the runs with plan_pair_units (0: its default)
    template <typename Tin, int S, int MODE>
    int duo256(const StftArgs&) {
        g_last_family = FAM_DUO256;
        using DP = Duo256Plan;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, DP::NT, DP::SMEM, [&] { stft_psd_duo256_kernel<Tin, S, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;      // the last CTA re-armed the counters
    }
    template <int LOG2N, typename Tin, int S, int MODE>
    int duo4(const StftArgs&) {
        g_last_family = FAM_DUO4;
        using DP = Duo4Plan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, DP::NT, DP::SMEM, [&] { stft_psd_duo4_kernel<LOG2N, Tin, S, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <int LOG2N, typename Tin, int MODE>
    int duo_cta(const StftArgs&) {
        g_last_family = FAM_DUO_CTA;
        using DP = DuoCtaPlan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, DP::NT, DP::SMEM, [&] { stft_psd_duo_cta_kernel<LOG2N, Tin, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <typename Tin, int S, int MODE>
    int duo(const StftArgs&) {
        g_last_family = FAM_DUO;
        using DP = DuoPlan;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, DP::NT, DP::SMEM, [&] { stft_psd_duo_kernel<Tin, S, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <int LOG2N, typename Tin, int SHIFT, int MODE>
    int warp(const StftArgs&) {
        g_last_family = FAM_WARP;
        using WP = WarpPlan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, WP::NT, WP::SMEM, [&] { stft_psd_warp_kernel<LOG2N, Tin, SHIFT, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <int LOG2N, typename Tin, int MODE>
    int big(const StftArgs&) {
        g_last_family = FAM_BIG;
        using BP = BigPlan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, BP::NT, BP::SMEM, [&] { stft_psd_big_kernel<LOG2N, Tin, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
    template <int LOG2N, typename Tin, int MODE>
    int cta(const StftArgs&) {
        g_last_family = FAM_CTA;
        using PL = Plan<LOG2N>;
        StftParams q = p;
        if (dynamic_units) q.work = work;
        emu::launch(grid, PL::NT, PL::SMEM, [&] { stft_psd_kernel<LOG2N, Tin, 1, MODE>(q); });
        return (work[0] == 0 && work[1] == 0) ? 0 : -100;
    }
};

// force_shift: -1 = the library's own choice, 0 = disable the sliding variant
extern "C" int emu_stft_psd(const void* x, int x_is_f64, long long batch, long long n, long long x_batch_stride,
                            int nperseg, int hop, const float* window, int detrend, double scale, int out_mode,
                            float db_floor, int kmin, int kmax, long long frame0, long long nframes, float* out,
                            long long out_batch_stride, int grid, int force_chunk, int band_mode) {
    StftArgs a{x, x_is_f64, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, out_mode,
               db_floor, kmin, kmax, frame0, nframes, out, out_batch_stride, band_mode};
    std::string err;
    if (validate_args(a, err) < 0) return validate_args(a, err);
    if (nperseg_support(nperseg) == 3 && !(getenv("B2S_NO_MIXED") && atoi(getenv("B2S_NO_MIXED")))) {
        g_last_family = FAM_MIXED;
        MixedParams mp{};
        if (!fill_mixed_params(a, mp)) return -2;
        std::vector<float> tw;
        make_mixed_table(nperseg, tw);
        mp.d.tw = reinterpret_cast<const float2*>(tw.data());
        if (mp.d.total_frames == 0) return 0;
        const unsigned g = (unsigned)(mp.d.total_frames < grid ? mp.d.total_frames : grid);
        if (x_is_f64) emu::launch(g, kMixedThreads, mixed_smem_bytes(nperseg), [&] { mixed_psd_kernel<double>(mp); });
        else emu::launch(g, kMixedThreads, mixed_smem_bytes(nperseg), [&] { mixed_psd_kernel<float>(mp); });
        return 0;
    }
    if (nperseg_support(nperseg) >= 2) {
        g_last_family = FAM_DFT;
        DftParams dp{};
        fill_dft_params(a, dp);
        std::vector<float> tw;
        make_dft_table(nperseg, tw);
        dp.tw = reinterpret_cast<const float2*>(tw.data());
        if (dp.total_frames == 0) return 0;
        const unsigned g = (unsigned)(dp.total_frames < grid ? dp.total_frames : grid);
        if (x_is_f64) emu::launch(g, kDftThreads, dft_smem_bytes(nperseg), [&] { dft_psd_kernel<double>(dp); });
        else emu::launch(g, kDftThreads, dft_smem_bytes(nperseg), [&] { dft_psd_kernel<float>(dp); });
        return 0;
    }
    EmuLauncher L;
    L.grid = (unsigned)grid;
    const int log2n = plan_stft(a, 1, (long long)grid * 4, L.p, err);
    if (log2n < 0) return log2n;
    if (force_chunk > 0) {
        L.p.chunk_frames = force_chunk;
        L.p.units_per_signal = (nframes + force_chunk - 1) / force_chunk;
        L.p.n_units = L.p.units_per_signal * batch;
    }
    std::vector<float> tw;
    make_tables(nperseg, tw);
    L.p.tw = reinterpret_cast<const float2*>(tw.data());
    if (L.p.n_units == 0) return 0;
    if (const char* v = getenv("B2S_NO_DUO")) L.allow_duo = (atoi(v) == 0);
    if (const char* v = getenv("B2S_DUO1024")) L.duo1024 = (atoi(v) != 0);
    if (const char* v = getenv("B2S_NO_DUO4")) L.allow_duo4 = (atoi(v) == 0);
    if (const char* v = getenv("B2S_NO_BIG")) L.allow_big = (atoi(v) == 0);
    if (const char* v = getenv("B2S_NO_PAIR")) L.allow_pair = (atoi(v) == 0);
    if (const char* v = getenv("B2S_PAIRQ")) L.allow_pairq = (atoi(v) != 0);
    if (const char* v = getenv("B2S_PAIR_UNITS")) L.pair_units = atoi(v);
    if (const char* v = getenv("B2S_STATIC_UNITS")) L.dynamic_units = (atoi(v) == 0);
    return dispatch_stft(a, L);
}

extern "C" int emu_batch_sum(const float* in, long long in_stride, int batch, int rows_per_slab, long long elems,
                             float* out, float post_scale) {
    const unsigned block = 64;
    const unsigned gx = (unsigned)((elems + block - 1) / block);
    const unsigned slabs = (unsigned)((batch + rows_per_slab - 1) / rows_per_slab);
    for (unsigned s = 0; s < slabs; ++s) {
        // blockIdx.y is not modelled by the emulator: fold it by offsetting pointers
        emu::launch(gx, block, 0, [&] {
            blockIdx.y = 0;
            const int r0 = (int)s * rows_per_slab;
            const int nb = (batch - r0 < rows_per_slab) ? batch - r0 : rows_per_slab;
            batch_sum_kernel(in + (long long)r0 * in_stride, in_stride, nb, rows_per_slab, elems,
                             out + (long long)s * elems, post_scale);
        });
    }
    return 0;
}

// the sum-fused frame-duo kernel followed by the fold over its sweep blocks, planned by the library's
// own plan_stft_sum with `grid` resident CTAs; returns the number of blocks (> 0) or an error (< 0)
extern "C" int emu_stft_psd_sum(const void* x, int x_is_f64, long long batch, long long n, long long x_batch_stride,
                                int nperseg, int hop, const float* window, int detrend, double scale,
                                long long frame0, long long nframes, float* out, long long out_batch_stride,
                                float* sum_out, float post_scale, int grid, int max_blocks) {
    StftArgs a{x, x_is_f64, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, 0,
               0.f, 0, nperseg / 2, frame0, nframes, out, out_batch_stride, 0};
    std::string err;
    if (validate_args(a, err) < 0) return validate_args(a, err);
    const long long elems = nframes * (nperseg / 2 + 1);
    if (nperseg == 1024) {
        // the SUM mode of the staged-sample pair kernel (sums in shared memory: the twin of the product's
        // tensor-memory kernel), `grid` resident CTAs of four warps
        using PP = PairPlan<10>;
        {
            EmuLauncher Lsel{};
            if (!pair_preferred(a, Lsel)) return -200;      // (as the library: float64 at hop 128 / 256 / 512 has no fused kernel)
        }
        StftParams p{};
        const int blocks = plan_stft_sum(a, (long long)grid * (PP::NT / 32), max_blocks, p, err, 1);
        if (blocks < 0) return blocks;
        std::vector<float> tw;
        make_tables(nperseg, tw);
        p.tw = reinterpret_cast<const float2*>(tw.data());
        std::vector<float> part((size_t)blocks * elems, std::nanf(""));
        p.acc = part.data();
        p.ring = PP::ring_samples(hop);
        const long long need = (p.n_units + PP::NT / 32 - 1) / (PP::NT / 32);
        const unsigned g = (unsigned)(need < grid ? need : grid);
        const size_t smem = PP::sum_smem_bytes(hop, x_is_f64 ? 8 : 4);
        if (x_is_f64) emu::launch(g, PP::NT, smem, [&] { stft_psd_pair_sum_kernel<10, double, 1>(p); });
        else emu::launch(g, PP::NT, smem, [&] { stft_psd_pair_sum_kernel<10, float, 1>(p); });
        emu_batch_sum(part.data(), elems, blocks, blocks, elems, sum_out, post_scale);
        return blocks;
    }
    if (nperseg == 2048 || nperseg == 4096) {      // (4096: emulator only -- the library does not ship it, b2s_inst_sum4.cu)
        // the SUM mode of the four-step frame-duo kernel (sums in shared memory: the twin of the product's kernel)
        const int s4 = duo4_slots(a);
        if (s4 != 2 && s4 != 4 && s4 != 8) return -200;
        StftParams p{};
        const int fpc = nperseg == 2048 ? Duo4Plan<11>::FPC : Duo4Plan<12>::FPC;
        const int blocks = plan_stft_sum(a, (long long)grid * fpc, max_blocks, p, err, 1);
        if (blocks < 0) return blocks;
        std::vector<float> tw;
        make_tables(nperseg, tw);
        p.tw = reinterpret_cast<const float2*>(tw.data());
        std::vector<float> part((size_t)blocks * elems, std::nanf(""));
        p.acc = part.data();
        const long long need = (p.n_units + fpc - 1) / fpc;
        const unsigned g = (unsigned)(need < grid ? need : grid);
        auto run = [&](auto kern, int nt, size_t smem) { emu::launch(g, nt, smem, [&] { kern(p); }); };
#define B2S_EMU_SUM4(L, T)                                                                                                  \
        do {                                                                                                                 \
            using DP = Duo4Plan<L>;                                                                                          \
            if (s4 == 2) run(stft_psd_duo4_sum_kernel<L, T, 2, 1>, DP::NT, DP::SUM_SMEM);                                   \
            else if (s4 == 4) run(stft_psd_duo4_sum_kernel<L, T, 4, 1>, DP::NT, DP::SUM_SMEM);                              \
            else run(stft_psd_duo4_sum_kernel<L, T, 8, 1>, DP::NT, DP::SUM_SMEM);                                           \
        } while (0)
        if (nperseg == 2048) {
            if (x_is_f64) B2S_EMU_SUM4(11, double); else B2S_EMU_SUM4(11, float);
        } else {
            if (x_is_f64) B2S_EMU_SUM4(12, double); else B2S_EMU_SUM4(12, float);
        }
#undef B2S_EMU_SUM4
        emu_batch_sum(part.data(), elems, blocks, blocks, elems, sum_out, post_scale);
        return blocks;
    }
    if (nperseg == 256) {
        // the SUM mode of the 256-point frame-duo kernel (sums in shared memory: the twin of the product's kernel)
        using DP = Duo256Plan;
        const int s256 = duo256_sum_slots(a, 8);
        if (!s256) return -200;
        StftParams p{};
        const int blocks = plan_stft_sum(a, (long long)grid * DP::FPC, max_blocks, p, err, 4);
        if (blocks < 0) return blocks;
        std::vector<float> tw;
        make_tables(nperseg, tw);
        p.tw = reinterpret_cast<const float2*>(tw.data());
        std::vector<float> part((size_t)blocks * elems, std::nanf(""));
        p.acc = part.data();
        const long long need = (p.n_units + DP::FPC - 1) / DP::FPC;
        const unsigned g = (unsigned)(need < grid ? need : grid);
        auto run = [&](auto kern) { emu::launch(g, DP::NT, DP::SUM_SMEM, [&] { kern(p); }); };
        if (x_is_f64) {
            if (s256 == 2) run(stft_psd_duo256_sum_kernel<double, 2, 1>);
            else if (s256 == 4) run(stft_psd_duo256_sum_kernel<double, 4, 1>);
            else if (s256 == 8) run(stft_psd_duo256_sum_kernel<double, 8, 1>);
            else run(stft_psd_duo256_sum_kernel<double, 16, 1>);
        } else {
            if (s256 == 2) run(stft_psd_duo256_sum_kernel<float, 2, 1>);
            else if (s256 == 4) run(stft_psd_duo256_sum_kernel<float, 4, 1>);
            else if (s256 == 8) run(stft_psd_duo256_sum_kernel<float, 8, 1>);
            else run(stft_psd_duo256_sum_kernel<float, 16, 1>);
        }
        emu_batch_sum(part.data(), elems, blocks, blocks, elems, sum_out, post_scale);
        return blocks;
    }
    const int slots = duo_slots(a, ilog2_exact(nperseg));
    if (!slots || slots > 8) return -200;
    StftParams p{};
    const int blocks = plan_stft_sum(a, (long long)grid * DuoPlan::FPC, max_blocks, p, err);
    if (blocks < 0) return blocks;
    std::vector<float> tw;
    make_tables(nperseg, tw);
    p.tw = reinterpret_cast<const float2*>(tw.data());
    std::vector<float> part((size_t)blocks * elems, std::nanf(""));
    p.acc = part.data();
    const long long need = (p.n_units + DuoPlan::FPC - 1) / DuoPlan::FPC;
    const unsigned g = (unsigned)(need < grid ? need : grid);
    auto run = [&](auto kern) { emu::launch(g, DuoPlan::NT, DuoSumPlan::SMEM, [&] { kern(p); }); };
    if (x_is_f64) {
        if (slots == 2) run(stft_psd_duo_sum_kernel<double, 2>);
        else if (slots == 4) run(stft_psd_duo_sum_kernel<double, 4>);
        else run(stft_psd_duo_sum_kernel<double, 8>);
    } else {
        if (slots == 2) run(stft_psd_duo_sum_kernel<float, 2>);
        else if (slots == 4) run(stft_psd_duo_sum_kernel<float, 4>);
        else run(stft_psd_duo_sum_kernel<float, 8>);
    }
    emu_batch_sum(part.data(), elems, blocks, blocks, elems, sum_out, post_scale);
    return blocks;
}

// plan_stft_sum alone: out = {sweeps per block, units per block, units}; returns the blocks
extern "C" int emu_plan_sum(long long batch, long long nframes, long long resident_groups, int max_blocks, long long* out) {
    static float dummy[4];
    StftArgs a{dummy, 0, batch, 512 + 128 * (nframes - 1), 512 + 128 * (nframes - 1), 512, 128, dummy, 1, 1.0, 0,
               0.f, 0, 256, 0, nframes, dummy, nframes * 257, 0};
    StftParams p{};
    std::string err;
    const int blocks = plan_stft_sum(a, resident_groups, max_blocks, p, err);
    out[0] = p.acc_rows;
    out[1] = p.units_per_signal;
    out[2] = p.n_units;
    return blocks;
}
