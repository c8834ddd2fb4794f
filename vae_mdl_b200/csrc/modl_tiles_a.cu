// modl_tiles_a.cu -- tile instantiations of the MoDL kernels for n_mix 8, 12, 16, 24 (x-conditioned classes), compiled in their
// own translation unit so that the build stays parallel.  See modl_launch.cuh (extra_tile_ppt) and modl_kernels.cuh.
#include "modl_kernels.cuh"

namespace vaemdl {
int launch_tiled_extra_a(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan) {
  switch (a.M) {
    case 8:
      return bwd ? launch_tiled<8, 1, true, 0>(a, st, plan) : launch_tiled<8, 1, false, 0>(a, st, plan);
    case 12:
      return bwd ? launch_tiled<6, 2, true, 0>(a, st, plan) : launch_tiled<6, 2, false, 0>(a, st, plan);
    case 16:
      return bwd ? launch_tiled<8, 2, true, 0>(a, st, plan) : launch_tiled<8, 2, false, 0>(a, st, plan);
    case 24:
      return bwd ? launch_tiled<6, 4, true, 0>(a, st, plan) : launch_tiled<6, 4, false, 0>(a, st, plan);
    default:
      return VAEMDL_EUNSUPPORTED;
  }
}
}  // namespace vaemdl
