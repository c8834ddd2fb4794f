// modl_tiles_c.cu -- tile instantiations of the MoDL kernels for n_mix 6, 14, 18, 28 (x-conditioned classes), compiled in
// their own translation unit so that the build stays parallel.  See modl_launch.cuh (extra_tile_ppt) and modl_kernels.cuh.
#include "modl_kernels.cuh"

namespace vaemdl {
int launch_tiled_extra_c(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan) {
  switch (a.M) {
    case 6:
      return bwd ? launch_tiled<6, 1, true, 0>(a, st, plan) : launch_tiled<6, 1, false, 0>(a, st, plan);
    case 14:
      return bwd ? launch_tiled<14, 1, true, 0>(a, st, plan) : launch_tiled<14, 1, false, 0>(a, st, plan);
    case 18:
      return bwd ? launch_tiled<6, 3, true, 0>(a, st, plan) : launch_tiled<6, 3, false, 0>(a, st, plan);
    case 28:
      return bwd ? launch_tiled<14, 2, true, 0>(a, st, plan) : launch_tiled<14, 2, false, 0>(a, st, plan);
    default:
      return VAEMDL_EUNSUPPORTED;
  }
}
}  // namespace vaemdl
