// The staged-sample pair kernel (b2s_pair_kernel.cuh) is instantiated through dispatch_tg in the
// b2s_inst_f32_* / b2s_inst_f64_* units; this unit only exports its two plain-epilogue entry points so
// that tools (tools/ubench) can look them up.
#include "b2s_launcher.hpp"

namespace b2s {
const void* pair_kernel_probe(int x_is_f64, int wide) {
    if (wide)
        return x_is_f64 ? (const void*)stft_psd_pair_wide_kernel<10, double, EPI_PLAIN>
                        : (const void*)stft_psd_pair_wide_kernel<10, float, EPI_PLAIN>;
    return x_is_f64 ? (const void*)stft_psd_pair_kernel<10, double, EPI_PLAIN>
                    : (const void*)stft_psd_pair_kernel<10, float, EPI_PLAIN>;
}
}  // namespace b2s
