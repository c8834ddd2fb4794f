// modl_tiles_b.cu -- tile instantiations of the MoDL kernels for n_mix 32, 40, 64 (x-conditioned classes), compiled in their
// own translation unit so that the build stays parallel.  See modl_launch.cuh (extra_tile_ppt) and modl_kernels.cuh.
#include "modl_kernels.cuh"

namespace vaemdl {
int launch_tiled_extra_b(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan) {
  switch (a.M) {
    case 32:
      return bwd ? launch_tiled<8, 4, true, 0>(a, st, plan) : launch_tiled<8, 4, false, 0>(a, st, plan);
    case 40:
      return bwd ? launch_tiled<10, 4, true, 0>(a, st, plan) : launch_tiled<10, 4, false, 0>(a, st, plan);
    case 64:
      return bwd ? launch_tiled<8, 8, true, 0>(a, st, plan) : launch_tiled<8, 8, false, 0>(a, st, plan);
    default:
      return VAEMDL_EUNSUPPORTED;
  }
}
}  // namespace vaemdl
