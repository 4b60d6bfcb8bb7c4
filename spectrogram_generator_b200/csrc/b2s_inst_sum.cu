// kernel instantiations: the sum-fused frame-duo kernel (b2s_duo_sum_kernel.cuh), float and double samples
#include "b2s_launcher.hpp"
#include "b2s_duo_sum_kernel.cuh"

namespace b2s {
const void* duo_sum_kernel_for(int x_is_f64, int slots) {
    if (x_is_f64) {
        switch (slots) {
            case 2: return (const void*)stft_psd_duo_sum_kernel<double, 2>;
            case 4: return (const void*)stft_psd_duo_sum_kernel<double, 4>;
            case 8: return (const void*)stft_psd_duo_sum_kernel<double, 8>;
            default: return nullptr;
        }
    }
    switch (slots) {
        case 2: return (const void*)stft_psd_duo_sum_kernel<float, 2>;
        case 4: return (const void*)stft_psd_duo_sum_kernel<float, 4>;
        case 8: return (const void*)stft_psd_duo_sum_kernel<float, 8>;
        default: return nullptr;
    }
}
}  // namespace b2s
