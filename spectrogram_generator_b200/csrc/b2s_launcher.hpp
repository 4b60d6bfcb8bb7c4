// The launcher the C-ABI entries hand to b2s::dispatch_tg, shared by the translation units the
// kernel instantiations are spread over (b2s_inst_*.cu: one (sample type, epilogue mode) each,
// compiled in parallel; b2s_api.cu holds the ABI, the caches and b2s_launch_any).
#pragma once

#include <cuda_runtime.h>

#include "b2s_dispatch.hpp"

// One launch of an STFT kernel of any family: persistent grid sized from the occupancy, work
// units sized from the grid (b2s::plan_stft).  Defined in b2s_api.cu.
// `dynamic`: the kernel draws its work units from an atomic counter (StftParams::work).
int b2s_launch_any(const void* kern, int nt, size_t smem, int fpc, const b2s::StftArgs& a, cudaStream_t stream,
                   bool dynamic = false);

// One launch of the staged-sample pair kernel (b2s_pair_kernel.cuh).  Defined in b2s_api.cu.
int b2s_launch_pair(const void* kern, const void* kern_mid, const void* kern_wide, int esz, const b2s::StftArgs& a,
                    cudaStream_t stream, bool dynamic);

// Records which kernel the calling thread launched last (b2s_last_kernel() in include/b2s.h).
void b2s_note_kernel(const char* family, const b2s::StftArgs& a);

// One launch of the staged-sample kernel for nperseg >= 2048 (b2s_pairq_kernel.cuh).  Defined in b2s_api.cu.
int b2s_launch_pairq(const void* kern, int group_threads, int nt_max, size_t (*smem_fn)(int, int), int esz,
                     const b2s::PairQConst& qc, const b2s::StftArgs& a, cudaStream_t stream, bool dynamic);

namespace b2s {

struct CudaLauncher {
    cudaStream_t stream;
    bool allow_duo = true;
    bool duo1024 = true;
    bool allow_duo4 = true;
    bool allow_big = true;
    bool allow_pair = true;
    bool allow_pairq = true;
    template <int LOG2N, typename Tin, int MODE>
    int pairq(const StftArgs& a) {
        using PP = PairQPlan<LOG2N>;
        b2s_note_kernel("stft_psd_pairq_kernel (staged samples: TMA ring, sub-sequence pairs in fp32x2, fused untangle + radix-Q)", a);
        PairQConst qc;
        make_pairq_const<PP::HW>(qc);
        return b2s_launch_pairq((const void*)stft_psd_pairq_kernel<LOG2N, Tin, MODE>, PP::G, PP::NT_MAX, &PP::smem_bytes,
                                (int)sizeof(Tin), qc, a, stream, dynamic_units);
    }
    template <int LOG2N, typename Tin, int MODE>
    int pair(const StftArgs& a) {
        b2s_note_kernel("stft_psd_pair_kernel (staged samples: TMA ring, sub-sequence pairs in fp32x2)", a);
        return b2s_launch_pair((const void*)stft_psd_pair_kernel<LOG2N, Tin, MODE>,
                               (const void*)stft_psd_pair_mid_kernel<LOG2N, Tin, MODE>,
                               (const void*)stft_psd_pair_wide_kernel<LOG2N, Tin, MODE>,
                               (int)sizeof(Tin), a, stream, dynamic_units);
    }
    template <int LOG2N, typename Tin, int MODE>
    int big(const StftArgs& a) {
        b2s_note_kernel("stft_psd_big_kernel (three passes, one frame per CTA)", a);
        using BP = BigPlan<LOG2N>;
        return b2s_launch_any((const void*)stft_psd_big_kernel<LOG2N, Tin, MODE>, BP::NT, BP::SMEM, BP::FPC, a, stream,
                              dynamic_units);
    }
    bool dynamic_units = true;       // all FFT kernel families: atomic work counter instead of static round-robin (large launches)
    template <typename Tin, int S, int MODE>
    int duo256(const StftArgs& a) {
        b2s_note_kernel("stft_psd_duo256_kernel (frame duo, 8 lanes)", a);
        using DP = Duo256Plan;
        return b2s_launch_any((const void*)stft_psd_duo256_kernel<Tin, S, MODE>, DP::NT, DP::SMEM, DP::FPC, a, stream,
                              dynamic_units);
    }
    template <int LOG2N, typename Tin, int S, int MODE>
    int duo4(const StftArgs& a) {
        b2s_note_kernel("stft_psd_duo4_kernel (four-step frame duo)", a);
        using DP = Duo4Plan<LOG2N>;
        return b2s_launch_any((const void*)stft_psd_duo4_kernel<LOG2N, Tin, S, MODE>, DP::NT, DP::SMEM, DP::FPC, a,
                              stream, dynamic_units);
    }
    template <int LOG2N, typename Tin, int MODE>
    int duo_cta(const StftArgs& a) {
        b2s_note_kernel("stft_psd_duo_cta_kernel (frame duo per CTA group)", a);
        using DP = DuoCtaPlan<LOG2N>;
        return b2s_launch_any((const void*)stft_psd_duo_cta_kernel<LOG2N, Tin, MODE>, DP::NT, DP::SMEM, DP::FPC, a,
                              stream, dynamic_units);
    }
    template <typename Tin, int S, int MODE>
    int duo(const StftArgs& a) {
        b2s_note_kernel("stft_psd_duo_kernel (frame duo, 16 lanes)", a);
        using DP = DuoPlan;
        return b2s_launch_any((const void*)stft_psd_duo_kernel<Tin, S, MODE>, DP::NT, DP::SMEM, DP::FPC, a, stream,
                              dynamic_units);
    }
    template <int LOG2N, typename Tin, int SHIFT, int MODE>
    int warp(const StftArgs& a) {
        b2s_note_kernel("stft_psd_warp_kernel (one frame per lane group)", a);
        using WP = WarpPlan<LOG2N>;
        return b2s_launch_any((const void*)stft_psd_warp_kernel<LOG2N, Tin, SHIFT, MODE>, WP::NT, WP::SMEM,
                              WP::FPC, a, stream, dynamic_units);
    }
    template <int LOG2N, typename Tin, int MODE>
    int cta(const StftArgs& a) {
        b2s_note_kernel("stft_psd_kernel (one frame per thread group)", a);
        using PL = Plan<LOG2N>;
        constexpr int MINB = (PL::NT <= 256) ? 2 : 1;
        return b2s_launch_any((const void*)stft_psd_kernel<LOG2N, Tin, MINB, MODE>, PL::NT, PL::SMEM, PL::FPC,
                              a, stream, dynamic_units);
    }
};


// the sum-fused frame-duo kernel for hop = slots * 32 (b2s_inst_sum.cu); nullptr if there is none.
// acc_tmem: running sums in tensor memory (else shared memory)
const void* duo_sum_kernel_for(int x_is_f64, int slots, int acc_tmem);

// the SUM mode of the 256-point frame-duo kernel, slots = 2 / 4 / 8 / 16 (b2s_inst_sum.cu); nullptr if there is none
const void* duo256_sum_kernel_for(int x_is_f64, int slots, int acc_tmem);
// the SUM mode of the four-step frame-duo kernel, nperseg 2048 / 4096, slots = 2 / 4 / 8 (b2s_inst_sum4.cu)
const void* duo4_sum_kernel_for(int log2n, int x_is_f64, int slots, int acc_tmem);
// the SUM mode of the staged-sample pair kernel, nperseg 1024 (b2s_inst_sum.cu)
const void* pair_sum_kernel_for(int x_is_f64, int acc_tmem);

// dispatch_tg<Tin, MODE> instantiated in b2s_inst_*.cu
int dispatch_f32_plain(const StftArgs& a, CudaLauncher& L);
int dispatch_f32_general(const StftArgs& a, CudaLauncher& L);
int dispatch_f32_band(const StftArgs& a, CudaLauncher& L);
int dispatch_f64_plain(const StftArgs& a, CudaLauncher& L);
int dispatch_f64_general(const StftArgs& a, CudaLauncher& L);
int dispatch_f64_band(const StftArgs& a, CudaLauncher& L);

}  // namespace b2s
