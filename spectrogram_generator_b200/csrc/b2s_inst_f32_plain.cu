// kernel instantiations: float samples, EPI_PLAIN epilogue (see b2s_launcher.hpp)
#include "b2s_launcher.hpp"

namespace b2s {
int dispatch_f32_plain(const StftArgs& a, CudaLauncher& L) { return dispatch_tg<float, EPI_PLAIN>(a, L); }
}  // namespace b2s
