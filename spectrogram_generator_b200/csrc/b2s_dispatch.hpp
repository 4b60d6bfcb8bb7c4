// Kernel selection shared by the CUDA library and the emulator harness.
//
//   nperseg <= 1024 : stft_psd_warp_kernel<LOG2N, Tin, SHIFT, MODE>
//       SHIFT = hop*16/nperseg in {2,4,8,14} (float samples, 2-element aligned frames,
//       nperseg >= 256) selects the sliding-register-window variant, 0 otherwise;
//   nperseg >= 2048 : stft_psd_kernel<LOG2N, Tin, MINB, MODE>  (multi-warp groups).
//
// A Launcher provides
//   template <int LOG2N, typename Tin, int SHIFT, int MODE> int warp(const StftArgs&);
//   template <int LOG2N, typename Tin, int MODE>            int cta(const StftArgs&);
//   template <typename Tin, int S, int MODE>                int duo(const StftArgs&);
//   template <int LOG2N, typename Tin, int MODE>            int duo_cta(const StftArgs&);
//   template <int LOG2N, typename Tin, int S, int MODE>     int duo4(const StftArgs&);
//   template <typename Tin, int S, int MODE>                int duo256(const StftArgs&);
//   template <int LOG2N, typename Tin, int MODE>            int big(const StftArgs&);   (8192, 16384)
//   bool allow_big;
//   template <int LOG2N, typename Tin, int MODE>            int pair(const StftArgs&);  (1024, staged samples)
//   bool allow_pair;
//   bool allow_duo, duo1024, allow_duo4;
//
//   nperseg == 512 with hop in {64, 128, 256, 448, 512} (2-element aligned frames) takes the packed
//   two-frames-per-lane stft_psd_duo_kernel<Tin, S, MODE> instead (b2s_duo_kernel.cuh), and
//   nperseg == 256 with hop in {32, 64, 128, 224, 256} stft_psd_duo256_kernel (b2s_duo256_kernel.cuh);
//   nperseg 1024 / 2048 / 4096 with hop = S * nperseg/16, S in {2, 4, 8} -- 1024 also S = 14, 16 -- (2-element aligned
//   frames) take the four-step stft_psd_duo4_kernel (b2s_duo4_kernel.cuh); with any other hop
//   stft_psd_duo_cta_kernel (b2s_duo_cta_kernel.cuh), measured 7-22 % faster than the one-frame
//   kernels on B200 (tools/ubench/duo_bench).
#pragma once

#include "b2s_host.hpp"
#include "b2s_big_kernel.cuh"
#include "b2s_duo256_kernel.cuh"
#include "b2s_duo4_kernel.cuh"
#include "b2s_duo_cta_kernel.cuh"
#include "b2s_duo_kernel.cuh"
#include "b2s_pair_kernel.cuh"
#include "b2s_pairq_kernel.cuh"
#include "b2s_warp_kernel.cuh"

namespace b2s {

inline bool frames_vec_aligned(const StftArgs& a) {
    const size_t esz = a.x_is_f64 ? 8 : 4;
    const bool base_ok = (reinterpret_cast<uintptr_t>(a.x) % (2 * esz)) == 0;
    return base_ok && (a.hop % 2 == 0) && (a.x_batch_stride % 2 == 0 || a.batch <= 1);
}

inline int sliding_shift(const StftArgs& a, int log2n) {
    if (a.x_is_f64 || log2n < 8 || log2n > 10 || !frames_vec_aligned(a)) return 0;
    const long long n16 = a.nperseg / 16;
    if (a.hop % n16) return 0;
    const long long s = a.hop / n16;
    return (s == 2 || s == 4 || s == 8 || s == 14) ? (int)s : 0;
}

// frame-duo kernel: nperseg 512, hop = S * 32 samples, S in {2, 4, 8}
inline int duo_slots(const StftArgs& a, int log2n) {
    if (log2n != 9 || !frames_vec_aligned(a)) return 0;
    // 14, 16: hop = 7/8 nperseg (the reference's default overlap) and hop = nperseg -- the register window
    // holds both frames whole; measured 71 -> 84 % and 76 -> 88 % of the HBM peak against the warp kernel.
    // Any other (even) hop takes the S = 16 variant too: it loads both frames afresh wherever B starts.
    const long long s = (a.hop % 32) ? 16 : a.hop / 32;
    return (s == 2 || s == 4 || s == 8 || s == 14) ? (int)s : 16;
}

// sum-fused 256-point frame-duo kernel (SUM mode of b2s_duo256_kernel.cuh): S = hop / 16 for hop 32 / 64 / 128
// (the same register layout and detrend tree as the per-sweep kernel of that hop: bit-identical rows); any other
// even hop takes S = 16, which holds both frames whole -- the per-sweep kernels of those hops (S = 14, 16) use
// the same per-frame arithmetic.
inline int duo256_sum_slots(const StftArgs& a, int log2n) {
    if (log2n != 8 || !frames_vec_aligned(a)) return 0;
    return (a.hop == 32) ? 2 : (a.hop == 64) ? 4 : (a.hop == 128) ? 8 : 16;
}

template <typename Tin, int MODE, class Launcher>
int dispatch_duo(const StftArgs& a, Launcher& L, int s) {
    switch (s) {
        case 2: return L.template duo<Tin, 2, MODE>(a);
        case 4: return L.template duo<Tin, 4, MODE>(a);
        case 14: return L.template duo<Tin, 14, MODE>(a);
        case 16: return L.template duo<Tin, 16, MODE>(a);
        default: return L.template duo<Tin, 8, MODE>(a);
    }
}

template <int LOG2N, typename Tin, int MODE, class Launcher>
int dispatch_warp_shift(const StftArgs& a, Launcher& L, int shift) {
    if constexpr (LOG2N == 9) {
        const int s = L.allow_duo ? duo_slots(a, LOG2N) : 0;
        if (s) return dispatch_duo<Tin, MODE>(a, L, s);
    }
    if constexpr (LOG2N == 8) {
        if (L.allow_duo && frames_vec_aligned(a)) {
            switch (a.hop) {
                case 32: return L.template duo256<Tin, 2, MODE>(a);
                case 64: return L.template duo256<Tin, 4, MODE>(a);
                case 128: return L.template duo256<Tin, 8, MODE>(a);
                // 7/8 nperseg (the reference's default overlap) and no overlap: 51 -> 74 %, 62 -> 77 % of the HBM peak
                case 224: return L.template duo256<Tin, 14, MODE>(a);
                default: return L.template duo256<Tin, 16, MODE>(a);        // no overlap, and any other even hop
            }
        }
    }
    if constexpr (LOG2N >= 8 && sizeof(Tin) == 4) {
        switch (shift) {
            case 2: return L.template warp<LOG2N, Tin, 2, MODE>(a);
            case 4: return L.template warp<LOG2N, Tin, 4, MODE>(a);
            case 8: return L.template warp<LOG2N, Tin, 8, MODE>(a);
            case 14: return L.template warp<LOG2N, Tin, 14, MODE>(a);
            default: break;
        }
    }
    return L.template warp<LOG2N, Tin, 0, MODE>(a);
}

// four-step duo kernel: nperseg 1024..4096, hop = S * nperseg/16, S in {2, 4, 8}
inline int duo4_slots(const StftArgs& a) {
    if (!frames_vec_aligned(a)) return 0;
    const long long n16 = a.nperseg / 16;
    if (a.hop % n16) return 0;
    const long long s = a.hop / n16;
    return (s == 2 || s == 4 || s == 8 || s == 14 || s == 16) ? (int)s : 0;
}

template <int LOG2N, typename Tin, int MODE, class Launcher>
int dispatch_duo_big(const StftArgs& a, Launcher& L) {
    if (L.allow_duo4) {
        switch (duo4_slots(a)) {
            case 2: return L.template duo4<LOG2N, Tin, 2, MODE>(a);
            case 4: return L.template duo4<LOG2N, Tin, 4, MODE>(a);
            case 8: return L.template duo4<LOG2N, Tin, 8, MODE>(a);
            // hop = 7/8 nperseg (the reference's default overlap) and hop = nperseg: the register window
            // holds both frames whole.  Measured on B200: nperseg 1024 gains (63.6 -> 66.9 % and
            // 69.7 -> 70.6 % of the HBM peak); 2048 / 4096 lose to the CTA kernel (55 -> 52 %, 50 -> 37 %).
            case 14:
                if constexpr (LOG2N == 10) return L.template duo4<LOG2N, Tin, 14, MODE>(a);
                break;
            case 16:
                if constexpr (LOG2N == 10) return L.template duo4<LOG2N, Tin, 16, MODE>(a);
                break;
            default: break;
        }
    }
    return L.template duo_cta<LOG2N, Tin, MODE>(a);
}

// nperseg 1024: does this call take the staged-sample pair kernel (b2s_pair_kernel.cuh)?  Any hop that keeps
// the frames 16-byte aligned -- except float64 samples at the four-step kernel's overlapping hops (128, 256,
// 512): there every frame re-reads and re-converts its staged doubles (two trips, 64 LDS.128 and 128 F2F per
// lane and frame, and the double ring leaves room for two CTAs per SM instead of three), while the frame-duo
// kernel converts each sample once on its way into the sliding register window.  Measured on B200
// (tools/microbench.py --set n1024x --dtype f64, 1000 x 40 000): hop 256 0.324 -> 0.201 ms, hop 512 0.160 ->
// 0.125, hop 128 (1024 x 100 000) 1.350 -> 0.853; from hop 896 up the pair kernel is level or ahead and stays.
// (Also used by the sum-fused entry: its rows must come from the same per-frame code as b2s_stft_psd_*'s.)
template <class Launcher>
bool pair_preferred(const StftArgs& a, const Launcher& L) {
    if (!L.allow_pair || !pair_kernel_ok(a.x, a.x_is_f64, a.batch, a.x_batch_stride, a.nperseg, a.hop, a.frame0) ||
        reinterpret_cast<uintptr_t>(a.window) % 16 != 0)
        return false;
    if (a.x_is_f64 && L.allow_duo && L.duo1024 && L.allow_duo4 && a.hop <= 512) {
        const int s4 = duo4_slots(a);
        if (s4 == 2 || s4 == 4 || s4 == 8) return false;
    }
    return true;
}

// staged-sample kernel for nperseg 2048 .. 16384 (b2s_pairq_kernel.cuh): 16-byte aligned frames
template <class Launcher>
bool pairq_ok(const StftArgs& a, const Launcher& L) {
    if (a.x_is_f64 && a.nperseg > 8192) return false;      // a float64 ring of 16384 samples does not fit beside the buffers
    return L.allow_pairq && pairq_kernel_ok(a.x, a.batch, a.x_batch_stride, a.nperseg, a.hop) &&
           reinterpret_cast<uintptr_t>(a.window) % 16 == 0;
}

template <typename Tin, int MODE, class Launcher>
int dispatch_tg(const StftArgs& a, Launcher& L) {
    const int log2n = ilog2_exact(a.nperseg);
    const int shift = sliding_shift(a, log2n);
    switch (log2n) {
        case 5: return dispatch_warp_shift<5, Tin, MODE>(a, L, shift);
        case 6: return dispatch_warp_shift<6, Tin, MODE>(a, L, shift);
        case 7: return dispatch_warp_shift<7, Tin, MODE>(a, L, shift);
        case 8: return dispatch_warp_shift<8, Tin, MODE>(a, L, shift);
        case 9: return dispatch_warp_shift<9, Tin, MODE>(a, L, shift);
        case 10:
            // staged-sample pair kernel: any hop that keeps the frames 16-byte aligned
            if (pair_preferred(a, L)) return L.template pair<10, Tin, MODE>(a);
            if (L.allow_duo && L.duo1024) return dispatch_duo_big<10, Tin, MODE>(a, L);
            return dispatch_warp_shift<10, Tin, MODE>(a, L, shift);
        case 11:
            if (pairq_ok(a, L)) return L.template pairq<11, Tin, MODE>(a);
            return L.allow_duo ? dispatch_duo_big<11, Tin, MODE>(a, L) : L.template cta<11, Tin, MODE>(a);
        case 12:
            if (pairq_ok(a, L)) return L.template pairq<12, Tin, MODE>(a);
            return L.allow_duo ? dispatch_duo_big<12, Tin, MODE>(a, L) : L.template cta<12, Tin, MODE>(a);
        case 13:
            if (pairq_ok(a, L)) return L.template pairq<13, Tin, MODE>(a);
            return L.allow_big ? L.template big<13, Tin, MODE>(a) : L.template cta<13, Tin, MODE>(a);
        case 14:
            if (pairq_ok(a, L)) return L.template pairq<14, Tin, MODE>(a);
            return L.allow_big ? L.template big<14, Tin, MODE>(a) : L.template cta<14, Tin, MODE>(a);
        default: return B2S_ERR_UNSUPPORTED;
    }
}

template <class Launcher>
int dispatch_stft(const StftArgs& a, Launcher& L) {
    // the reference's call (linear power, every bin) takes the branch-free epilogue
    const bool general = (a.out_mode != B2S_OUT_LINEAR) || a.kmin != 0 || a.kmax != a.nperseg / 2;
    const int mode = a.band_mode ? EPI_BAND : (general ? EPI_GENERAL : EPI_PLAIN);
    if (a.x_is_f64) {
        if (mode == EPI_BAND) return dispatch_tg<double, EPI_BAND>(a, L);
        return mode == EPI_GENERAL ? dispatch_tg<double, EPI_GENERAL>(a, L) : dispatch_tg<double, EPI_PLAIN>(a, L);
    }
    if (mode == EPI_BAND) return dispatch_tg<float, EPI_BAND>(a, L);
    return mode == EPI_GENERAL ? dispatch_tg<float, EPI_GENERAL>(a, L) : dispatch_tg<float, EPI_PLAIN>(a, L);
}

}  // namespace b2s
