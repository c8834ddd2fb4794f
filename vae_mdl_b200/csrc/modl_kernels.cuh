#pragma once
// modl_kernels.cuh -- mixture-of-discretized-logistics log-likelihood (forward) and parameter gradient (backward):
// kernel templates, launchers and the host-side implementation shared by modl_kernels.cu (means chained on the observed
// x, utils/mdl.py / utils/mdl_openai*.py) and modl_plain.cu (means chained on the means, utils/mdl_plain.py).
//
// Replaces utils/mdl.py:56-207, utils/mdl_openai.py:83-157, utils/mdl_openai_iwae.py:33-67 and the autodiff of
// models/model05.py:141-145 (see include/vaemdl.h).
//
// Data movement.  The parameter tensor is a flat stream of rows (one 40*M-byte row per pixel-sample).  Every warp
// owns a run of consecutive tiles and runs its own pipeline: a 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) brings a
// tile of PPT rows into the warp's shared-memory slot and signals an mbarrier; each lane reads its own row (or its
// MC-mixture chunk of the row) with 64-bit shared loads inside a rolled loop over component pairs.  The backward
// kernel overwrites the row in place with the unscaled gradients, rescales them once the pixel's mixture sum is
// known and hands the tile back with a bulk store.  Per-image sums stay in registers (float64) along the run and leave
// the warp once per image (finish.cu adds them up).  There is no CTA-wide synchronisation.
//
// Work split.  n_mix 10 / 20 / 30: M = MC * LPP, LPP lanes share a pixel, each owning MC components, two components
// per packed register (modl_tile_kernel: 10x1, 10x2, 10x3).  n_mix 1..9: one lane owns two pixels, the same component
// of both per packed register (modl_pp_kernel).  Any other M runs on modl_rt_kernel: the tiled pipeline with the split of
// a pixel over lanes chosen at run time.  The two-pass gradient of every tile shape with an even MC <= 14 keeps the tile it
// works on in TENSOR MEMORY (modl_tile_tm_kernel, modl_tm.cuh: tcgen05.ld / tcgen05.st, one TMEM lane per thread = per pixel
// row): the shared-memory slot is then only the TMA staging area, the next tile lands while this one goes through its first
// pass, and the second pass swaps gradient out / next parameters in block by block.  The cooperative one-launch step
// (modl_step_kernel) is opt-in.  bfloat16 parameters are widened / narrowed in place in the shared-memory slot, or stay bf16.
// Template parameter AR selects what the green / blue means are chained on (pair_eval).
//
// Files: modl_core.cuh (arguments, helpers, pair_eval), modl_tile.cuh (tiled kernel + one-launch step), modl_tm.cuh (gradient
// kernel on tensor memory), modl_pp.cuh (pixel pairs), modl_rt.cuh (run-time tile, generic), modl_launch.cuh (launchers + host
// implementation).
#include "modl_launch.cuh"
