// kernel instantiations: float samples, EPI_GENERAL epilogue (see b2s_launcher.hpp)
#include "b2s_launcher.hpp"

namespace b2s {
int dispatch_f32_general(const StftArgs& a, CudaLauncher& L) { return dispatch_tg<float, EPI_GENERAL>(a, L); }
}  // namespace b2s
