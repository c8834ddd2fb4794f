// modl_tiles_d.cu -- tile instantiations of the MoDL kernels for n_mix 36, 48, 50, 56, 60 (x-conditioned classes), compiled in
// their own translation unit so that the build stays parallel.  See modl_launch.cuh (extra_tile_ppt) and modl_kernels.cuh.
#include "modl_kernels.cuh"

namespace vaemdl {
int launch_tiled_extra_d(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan) {
  switch (a.M) {
    case 36:
      return bwd ? launch_tiled<12, 3, true, 0>(a, st, plan) : launch_tiled<12, 3, false, 0>(a, st, plan);
    case 48:
      return bwd ? launch_tiled<12, 4, true, 0>(a, st, plan) : launch_tiled<12, 4, false, 0>(a, st, plan);
    case 50:
      return bwd ? launch_tiled<10, 5, true, 0>(a, st, plan) : launch_tiled<10, 5, false, 0>(a, st, plan);
    case 56:
      return bwd ? launch_tiled<14, 4, true, 0>(a, st, plan) : launch_tiled<14, 4, false, 0>(a, st, plan);
    case 60:
      return bwd ? launch_tiled<12, 5, true, 0>(a, st, plan) : launch_tiled<12, 5, false, 0>(a, st, plan);
    default:
      return VAEMDL_EUNSUPPORTED;
  }
}
}  // namespace vaemdl
