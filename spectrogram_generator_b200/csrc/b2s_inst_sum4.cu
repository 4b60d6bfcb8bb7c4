// kernel instantiations: the SUM mode of the four-step frame-duo kernel (b2s_duo4_kernel.cuh), nperseg 2048 at
// hop 256 / 512 / 1024 -- running sums in tensor memory (the product path) or in shared memory
// (B2S_SUM_ACC_SMEM=1; also the twin the occupancy is taken from).
// nperseg 4096 is NOT instantiated: measured on B200 (1000 x 100 000, tools/microbench.py --set mean) the fused
// form loses to kernel + two-pass sum there (hop 1024: 0.931 vs 0.786 ms, hop 2048: 0.535 vs 0.463) -- without
// the sliding register window every sweep re-requests all of a frame's strided samples through L1, and the 17 + 17
// packed sums and records of a task spill (ptxas: 136-168 bytes of stack).
#include "b2s_launcher.hpp"

namespace b2s {
template <int LOG2N, typename Tin, int TM>
static const void* pick4(int slots) {
    switch (slots) {
        case 2: return (const void*)stft_psd_duo4_sum_kernel<LOG2N, Tin, 2, TM>;
        case 4: return (const void*)stft_psd_duo4_sum_kernel<LOG2N, Tin, 4, TM>;
        case 8: return (const void*)stft_psd_duo4_sum_kernel<LOG2N, Tin, 8, TM>;
        default: return nullptr;
    }
}

const void* duo4_sum_kernel_for(int log2n, int x_is_f64, int slots, int acc_tmem) {
    if (log2n == 11) {
        if (x_is_f64) return acc_tmem ? pick4<11, double, 2>(slots) : pick4<11, double, 1>(slots);
        return acc_tmem ? pick4<11, float, 2>(slots) : pick4<11, float, 1>(slots);
    }
    return nullptr;
}
}  // namespace b2s
