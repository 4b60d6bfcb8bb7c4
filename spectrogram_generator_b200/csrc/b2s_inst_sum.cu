// kernel instantiations: the sum-fused frame-duo kernel (b2s_duo_sum_kernel.cuh), float and double samples,
// running sums in tensor memory (the product path) or in shared memory (B2S_SUM_ACC_SMEM=1; also the twin
// the occupancy is taken from: the occupancy query answers 1 for a kernel that allocates tensor memory,
// although three of these CTAs -- 64 of the 512 columns each -- do share an SM)
#include "b2s_launcher.hpp"
#include "b2s_duo_sum_kernel.cuh"

namespace b2s {
template <typename Tin, int TM>
static const void* pick(int slots) {
    switch (slots) {
        case 2: return (const void*)stft_psd_duo_sum_kernel<Tin, 2, TM>;
        case 4: return (const void*)stft_psd_duo_sum_kernel<Tin, 4, TM>;
        case 8: return (const void*)stft_psd_duo_sum_kernel<Tin, 8, TM>;
        default: return nullptr;
    }
}

// the SUM mode of the staged-sample pair kernel (nperseg 1024): sums in tensor memory, or the shared-memory twin
const void* pair_sum_kernel_for(int x_is_f64, int acc_tmem) {
    if (x_is_f64)
        return acc_tmem ? (const void*)stft_psd_pair_sum_kernel<10, double, 2> : (const void*)stft_psd_pair_sum_kernel<10, double, 1>;
    return acc_tmem ? (const void*)stft_psd_pair_sum_kernel<10, float, 2> : (const void*)stft_psd_pair_sum_kernel<10, float, 1>;
}

// the SUM mode of the 256-point frame-duo kernel (b2s_duo256_kernel.cuh)
template <typename Tin, int TM>
static const void* pick256(int slots) {
    switch (slots) {
        case 2: return (const void*)stft_psd_duo256_sum_kernel<Tin, 2, TM>;
        case 4: return (const void*)stft_psd_duo256_sum_kernel<Tin, 4, TM>;
        case 8: return (const void*)stft_psd_duo256_sum_kernel<Tin, 8, TM>;
        case 16: return (const void*)stft_psd_duo256_sum_kernel<Tin, 16, TM>;
        default: return nullptr;
    }
}

const void* duo256_sum_kernel_for(int x_is_f64, int slots, int acc_tmem) {
    if (x_is_f64) return acc_tmem ? pick256<double, 2>(slots) : pick256<double, 1>(slots);
    return acc_tmem ? pick256<float, 2>(slots) : pick256<float, 1>(slots);
}

const void* duo_sum_kernel_for(int x_is_f64, int slots, int acc_tmem) {
    if (x_is_f64) return acc_tmem ? pick<double, 1>(slots) : pick<double, 0>(slots);
    return acc_tmem ? pick<float, 1>(slots) : pick<float, 0>(slots);
}
}  // namespace b2s
