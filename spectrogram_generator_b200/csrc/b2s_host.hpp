// Host-side launch planning shared by the CUDA library (b2s_api.cu) and the
// CPU emulator harness (tests/emu).  Plain C++: no CUDA runtime calls here.
#pragma once

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b2s.h"
#include "b2s_kernels.cuh"
#include "b2s_dft_kernel.cuh"
#include "b2s_mixed_kernel.cuh"

namespace b2s {

// The arguments of the C-ABI entry b2s_stft_psd_* (include/b2s.h).
struct StftArgs {
    const void* x;
    int x_is_f64;
    long long batch, n, x_batch_stride;
    int nperseg, hop;
    const float* window;
    int detrend;
    double scale;
    int out_mode;
    float db_floor;
    int kmin, kmax;
    long long frame0, nframes;
    float* out;
    long long out_batch_stride;
    int band_mode = 0;           // 1: out is [batch][nframes], the sum of bins kmin..kmax per frame
};

inline int ilog2_exact(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// exp(-2 pi i num / den), computed in double and rounded once
inline void put_w(std::vector<float>& out, size_t idx, long long num, long long den) {
    const double a = 2.0 * M_PI * (double)(num % den) / (double)den;
    out[2 * idx] = (float)std::cos(a);
    out[2 * idx + 1] = (float)(-std::sin(a));
}

// The constant tables of Plan<LOG2N> (see b2s_kernels.cuh), as interleaved floats.
template <int LOG2N>
inline void make_tables_t(std::vector<float>& out) {
    using PL = Plan<LOG2N>;
    out.assign(2 * (size_t)PL::TABLE, 0.f);
    if (PL::P >= 2)
        for (int r = 1; r < 16; ++r)
            for (int jm = 0; jm < 16; ++jm) put_w(out, PL::OFF_P1 + (r - 1) * 16 + jm, (long long)r * jm, 256);
    if (PL::P == 3)
        for (int r = 1; r < 16; ++r)
            for (int jm = 0; jm < 256; ++jm) put_w(out, PL::OFF_P2 + (r - 1) * 256 + jm, (long long)r * jm, 4096);
    for (int r = 1; r < PL::GF; ++r)
        for (int k = 0; k < PL::NS; ++k) put_w(out, PL::OFF_FIN + (r - 1) * PL::NS + k, (long long)r * k, PL::M);
    for (int k = 0; k <= PL::M; ++k) put_w(out, PL::OFF_POST + k, k, PL::N);
    if constexpr (LOG2N >= 11) {
        // twiddle pairs of the staged-sample kernel (b2s_pairq_kernel.cuh), [2 sp + v][kap] float4 with
        // kap = 0 .. 128: (wr[r], wr[r+1], wi[r], wi[r+1]) for r = 4 sp + 2 v, w[r] = W_N^(r kap)
        constexpr int HW = PL::N / 1024;
        constexpr int OFF_PQ = PL::TABLE + (PL::TABLE & 1);
        out.resize(2 * (size_t)(OFF_PQ + 129 * 4 * HW), 0.f);
        for (int kap = 0; kap <= 128; ++kap)
            for (int sp = 0; sp < HW; ++sp)
                for (int v = 0; v < 2; ++v) {
                    float* q = out.data() + 2 * (size_t)OFF_PQ + 4 * ((size_t)(2 * sp + v) * 129 + kap);
                    for (int h = 0; h < 2; ++h) {
                        const long long r = 4 * sp + 2 * v + h;
                        const double a = 2.0 * M_PI * (double)((r * kap) % PL::N) / (double)PL::N;
                        q[h] = (float)std::cos(a);
                        q[2 + h] = (float)(-std::sin(a));
                    }
                }
    }
}

inline long long frames_available(long long n, int nperseg, int hop) {
    return (n < nperseg) ? 0 : (n - nperseg) / hop + 1;
}

// Validates the arguments and fills the kernel parameter block (everything
// except the device twiddle pointer).  `resident_groups` is how many frame
// groups the launch can keep resident (grid * groups per CTA); it sizes the
// work units so that every group gets several runs of consecutive frames.
constexpr int kMaxNperseg = 16384;

// 1: radix-16 FFT kernels (powers of two 32..16384); 3: mixed-radix kernel (other lengths >= 256 whose
// prime factors are all <= 13); 2: direct-DFT kernel (everything else up to 16384); 0: unsupported.
inline int nperseg_support(int nperseg) {
    if (nperseg < 1 || nperseg > kMaxNperseg) return 0;
    const int l = ilog2_exact(nperseg);
    if (l >= 5 && l <= 14) return 1;
    return mixed_supported(nperseg) ? 3 : 2;
}

// Argument checks common to every kernel family (mirrors the shapes SciPy accepts).
inline int validate_args(const StftArgs& a, std::string& err) {
    if (a.nperseg < 1 || a.hop < 1 || a.batch < 0 || a.nframes < 0 || a.n < 0) {
        err = "b2s: nperseg, hop must be >= 1 and sizes non-negative";
        return B2S_ERR_BAD_ARG;
    }
    if (nperseg_support(a.nperseg) == 0) {
        err = "b2s: nperseg must be in [1, 16384]";
        return B2S_ERR_UNSUPPORTED;
    }
    if (!a.x || !a.window || !a.out) {
        err = "b2s: null pointer";
        return B2S_ERR_BAD_ARG;
    }
    const int K = a.nperseg / 2 + 1;
    if (a.kmin < 0 || a.kmax >= K || a.kmin > a.kmax) {
        err = "b2s: bin crop [kmin,kmax] outside [0, nperseg/2]";
        return B2S_ERR_BAD_ARG;
    }
    if (a.frame0 < 0 || a.frame0 + a.nframes > frames_available(a.n, a.nperseg, a.hop)) {
        err = "b2s: frame range exceeds (n - nperseg)//hop + 1";
        return B2S_ERR_BAD_ARG;
    }
    if (a.nframes > 0x7fffffffLL || a.batch > 0x7fffffffLL) {
        err = "b2s: nframes/batch too large for one launch";
        return B2S_ERR_BAD_ARG;
    }
    const int kout = a.kmax - a.kmin + 1;
    if (a.band_mode && a.out_mode != B2S_OUT_LINEAR) {
        err = "b2s: band power is linear";
        return B2S_ERR_BAD_ARG;
    }
    if (a.batch > 1 && (a.x_batch_stride < a.n || a.out_batch_stride < a.nframes * (long long)(a.band_mode ? 1 : kout))) {
        err = "b2s: batch strides overlap";
        return B2S_ERR_BAD_ARG;
    }
    return B2S_OK;
}

inline int plan_stft(const StftArgs& a, int groups_per_cta, long long resident_groups, StftParams& p,
                     std::string& err, bool dynamic = false) {
    const int log2n = ilog2_exact(a.nperseg);
    const int vrc = validate_args(a, err);
    if (vrc != B2S_OK) return vrc;
    if (log2n < 5 || log2n > 14) {
        err = "b2s: the FFT kernels take power-of-two nperseg in [32, 16384]";
        return B2S_ERR_UNSUPPORTED;
    }
    const int kout = a.kmax - a.kmin + 1;
    (void)kout;
    p.x = a.x;
    p.x_batch_stride = a.x_batch_stride;
    p.frame0 = a.frame0;
    p.out_batch_stride = a.out_batch_stride;
    p.window = a.window;
    p.tw = nullptr;
    p.work = nullptr;
    p.acc = nullptr;
    p.acc_rows = 0;
    p.acc_batch = 0;
    p.out = a.out;
    p.nframes = (int)a.nframes;
    p.hop = a.hop;
    p.detrend = a.detrend ? 1 : 0;
    p.out_mode = a.out_mode ? 1 : 0;
    p.kmin = a.kmin;
    p.kmax = a.kmax;
    p.scale = (float)a.scale;
    p.db_floor = a.db_floor;
    // vector loads need every frame start on a 2-element boundary
    const size_t esz = a.x_is_f64 ? 8 : 4;
    const bool base_ok = (reinterpret_cast<uintptr_t>(a.x) % (2 * esz)) == 0;
    p.vec_ok = (base_ok && (a.hop % 2 == 0) && (a.x_batch_stride % 2 == 0 || a.batch <= 1)) ? 1 : 0;
    // work units: runs of consecutive frames of one signal
    long long total = a.batch * a.nframes;
    // static round-robin: ~4 units per resident group; dynamic (atomic counter): ~6 smaller ones
    long long want_units = resident_groups * (dynamic ? 6 : 4);
    long long cf = (want_units > 0) ? (total + want_units - 1) / want_units : a.nframes;
    if (cf < 1) cf = 1;
    if (cf > 64) cf = 64;
    if (cf > 1 && (cf & 1)) ++cf;          // the frame-pair kernel walks a run two frames at a time
    if (cf > a.nframes) cf = a.nframes > 0 ? a.nframes : 1;
    p.chunk_frames = (int)cf;
    p.units_per_signal = (a.nframes + cf - 1) / cf;
    p.n_units = p.units_per_signal * a.batch;
    (void)groups_per_cta;
    return log2n;
}

// Work units of the staged-sample pair kernel (b2s_pair_kernel.cuh): one run per warp at a time.  A run
// starts by staging nperseg + hop samples, so runs are longer than the other families' (about
// `units_per_warp` runs per resident warp) and a signal is cut into runs of equal, even length.
inline void plan_pair_units(const StftArgs& a, long long resident_warps, int units_per_warp, bool dynamic,
                            StftParams& p) {
    if (units_per_warp <= 0) units_per_warp = dynamic ? 4 : 2;
    if (resident_warps < 1) resident_warps = 1;
    const long long total = a.batch * a.nframes;
    const long long want = resident_warps * units_per_warp;
    long long cf = (total + want - 1) / want;
    if (cf < 4) cf = 4;
    if (cf > 512) cf = 512;
    if (cf > a.nframes) cf = a.nframes > 0 ? a.nframes : 1;
    long long ups = (a.nframes + cf - 1) / cf;
    if (ups < 1) ups = 1;
    cf = (a.nframes + ups - 1) / ups;
    if (cf & 1) ++cf;
    p.chunk_frames = (int)cf;
    p.units_per_signal = (a.nframes + cf - 1) / cf;
    p.n_units = p.units_per_signal * a.batch;
}

// Work units of the sum-fused frame-duo kernel (b2s_duo_sum_kernel.cuh): a unit is one frame duo
// over a block of `rows` consecutive sweeps, so every unit costs the same and the static
// round-robin is balanced when the units fill the resident lane groups a whole number of times.
// Fewest rounds x (rows + 1 for the start-up of a unit) wins, fewer blocks (less partial-sum
// traffic) break ties.  Returns the number of blocks (<= max_blocks).
// `duos_per_warp`: the lane groups of a warp walk the same sweep block, so the units per block are rounded up
// to a multiple of it -- 2 for the 512-point frame-duo kernel, 4 for the 256-point one (b2s_duo256_kernel.cuh),
// 1 for the SUM mode of the staged-sample pair kernel (b2s_pair_kernel.cuh: one pair of frames per warp).
inline int plan_stft_sum(const StftArgs& a, long long resident_groups, int max_blocks, StftParams& p,
                         std::string& err, int duos_per_warp = 2) {
    const int log2n = plan_stft(a, 1, resident_groups, p, err, false);
    if (log2n < 0) return log2n;
    const long long nduos = (a.nframes + 1) / 2, ups = (nduos + duos_per_warp - 1) / duos_per_warp * duos_per_warp;
    if (resident_groups < 1) resident_groups = 1;
    long long best_cost = -1, best_rows = a.batch > 0 ? a.batch : 1;
    for (long long nb = 1; nb <= max_blocks && nb <= a.batch; ++nb) {
        const long long rows = (a.batch + nb - 1) / nb, blocks = (a.batch + rows - 1) / rows;
        const long long rounds = (blocks * ups + resident_groups - 1) / resident_groups;
        const long long cost = rounds * (rows + 1) * 4096 + blocks;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best_rows = rows;
        }
    }
    const long long blocks = a.batch > 0 ? (a.batch + best_rows - 1) / best_rows : 0;
    p.acc = nullptr;
    p.acc_rows = (int)best_rows;
    p.acc_batch = (int)a.batch;
    p.chunk_frames = 2;
    p.units_per_signal = ups;
    p.n_units = blocks * ups;
    return (int)blocks;
}

#define B2S_DISPATCH_LOG2N(log2n, F)   \
    switch (log2n) {                   \
        case 5: F(5); break;           \
        case 6: F(6); break;           \
        case 7: F(7); break;           \
        case 8: F(8); break;           \
        case 9: F(9); break;           \
        case 10: F(10); break;         \
        case 11: F(11); break;         \
        case 12: F(12); break;         \
        case 13: F(13); break;         \
        case 14: F(14); break;         \
        default: break;                \
    }

}  // namespace b2s

namespace b2s {
inline void make_tables(int nperseg, std::vector<float>& out) {
    const int log2n = ilog2_exact(nperseg);
#define B2S_TBL(L) make_tables_t<L>(out)
    B2S_DISPATCH_LOG2N(log2n, B2S_TBL)
#undef B2S_TBL
}
}  // namespace b2s

namespace b2s {
// W_N^j for the direct-DFT kernel
inline void make_dft_table(int nperseg, std::vector<float>& out) {
    out.assign(2 * (size_t)nperseg, 0.f);
    for (int j = 0; j < nperseg; ++j) put_w(out, j, j, nperseg);
}

// [Mc] W_Mc^j, then (even nperseg) [Mc + 1] W_N^k, for the mixed-radix kernel
inline void make_mixed_table(int nperseg, std::vector<float>& out) {
    const bool even = (nperseg % 2) == 0;
    const int mc = even ? nperseg / 2 : nperseg;
    out.assign(2 * (size_t)(mc + (even ? mc + 1 : 0)), 0.f);
    for (int j = 0; j < mc; ++j) put_w(out, j, j, mc);
    if (even)
        for (int k = 0; k <= mc; ++k) put_w(out, mc + k, k, nperseg);
}

inline void fill_dft_params(const StftArgs& a, DftParams& p);
inline bool fill_mixed_params(const StftArgs& a, MixedParams& mp) {
    fill_dft_params(a, mp.d);
    mp.mc = (a.nperseg % 2 == 0) ? a.nperseg / 2 : a.nperseg;
    mp.npass = mixed_radices(mp.mc, mp.radix);
    return mp.npass > 0;
}

inline void fill_dft_params(const StftArgs& a, DftParams& p) {
    p.x = a.x;
    p.x_batch_stride = a.x_batch_stride;
    p.frame0 = a.frame0;
    p.out_batch_stride = a.out_batch_stride;
    p.total_frames = a.batch * a.nframes;
    p.window = a.window;
    p.tw = nullptr;
    p.out = a.out;
    p.nframes = (int)a.nframes;
    p.nperseg = a.nperseg;
    p.hop = a.hop;
    p.detrend = a.detrend ? 1 : 0;
    p.out_mode = a.out_mode ? 1 : 0;
    p.kmin = a.kmin;
    p.kmax = a.kmax;
    p.scale = (float)a.scale;
    p.db_floor = a.db_floor;
    p.band = a.band_mode;
}
}  // namespace b2s
