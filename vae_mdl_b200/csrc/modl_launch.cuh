#pragma once
// modl_launch.cuh -- launchers (grid / shared-memory sizing, run-time tile geometry, kernel choice per n_mix) and the host-side
// implementation of the MoDL entry points, shared by modl_kernels.cu and modl_plain.cu.
#include "modl_pp.cuh"
#include "modl_rt.cuh"
#include "modl_tile.cuh"
#include "modl_tm.cuh"

namespace vaemdl {

// ---- launchers -------------------------------------------------------------------------------------------------------------
// Two shapes per kernel: (2 slots, <=8 warps, 255 registers) and (1 slot, <=16 warps, 128 registers).  VAEMDL_TUNE
// ("fwd=S:W,bwd=S:W", S slots, W warps per CTA) overrides the built-in choice; used by the tuning sweeps under tools/.
struct Shape {
  int slots, warps;
};
static Shape tune_shape(bool bwd, Shape dflt) {
  const char* env = getenv("VAEMDL_TUNE");
  if (!env) return dflt;
  const char* key = bwd ? "bwd=" : "fwd=";
  const char* p = strstr(env, key);
  if (!p) return dflt;
  int s = 0, w = 0;
  if (sscanf(p + 4, "%d:%d", &s, &w) == 2 && (s == 1 || s == 2) && w >= 1 && w <= 16) return Shape{s, w};
  return dflt;
}

// Small problems: with only a few tiles per warp the rounding of tiles/warp up to an integer costs more than a little
// occupancy does, so pick the warp count (>= 10) whose runs come out most even.  Large problems keep `max_warps`.
static int pick_warps(long long num_tiles, int sm_count, int max_warps) {
  if (getenv("VAEMDL_TUNE")) return max_warps;
  if (num_tiles >= static_cast<long long>(sm_count) * max_warps * 8) return max_warps;
  int best = max_warps;
  double best_score = -1.0;
  for (int w = max_warps; w >= 10 && w >= max_warps - 6; --w) {
    const double per = static_cast<double>(num_tiles) / (static_cast<double>(sm_count) * w);
    if (per <= 1.0) break;  // fewer tiles than warps: the grid shrinks instead
    const double longest = static_cast<double>((num_tiles + static_cast<long long>(sm_count) * w - 1) / (static_cast<long long>(sm_count) * w));
    const double score = per / longest * (0.8 + 0.2 * w / max_warps);
    if (score > best_score) {
      best_score = score;
      best = w;
    }
  }
  return best;
}

// forward kernels (VAEMDL_FWD_WARPS=<n>: exactly n warps per CTA, A/B)
static int pick_warps_fwd(long long num_tiles, int sm_count, int max_warps) {
  static const int forced = [] { const char* e = getenv("VAEMDL_FWD_WARPS"); return e ? atoi(e) : 0; }();
  if (forced > 0) return forced < max_warps ? forced : max_warps;
  return pick_warps(num_tiles, sm_count, max_warps);
}

struct L2Opt {  // VAEMDL_L2="rev=0|1,keep=<MB>,hint=0|1": L2 reuse between the forward and the backward kernel of a step
  int rev = 1, keep_mb = 48, hint = 0;  // measured on B200: profiles/r01_l2_reuse.txt
  L2Opt() {
    const char* e = getenv("VAEMDL_L2");
    if (!e) return;
    const char* q;
    if ((q = strstr(e, "rev="))) rev = atoi(q + 4);
    if ((q = strstr(e, "keep="))) keep_mb = atoi(q + 5);
    if ((q = strstr(e, "hint="))) hint = atoi(q + 5);
  }
};
static void apply_l2_opt(ModlArgs& a, long long total_warps, long long tile_bytes) {
  static const L2Opt opt;
  a.reverse = opt.rev;
  a.bwd_hint = opt.hint;
  a.keep_tiles = static_cast<int>((static_cast<long long>(opt.keep_mb) << 20) / (total_warps * tile_bytes));
  if (opt.keep_mb > 0 && a.keep_tiles < 1) a.keep_tiles = 1;
}

// n_mix 5, x-conditioned class, float32: a 32-row tile is 6.4 KB, so TWO slots per warp fit next to 16 warps (as for bfloat16
// tiles of n_mix 10): the next tile lands while this one is processed.  Measured on B200 (tools/ab_m5.sh, profiles/r02n):
// at 16 x 64 x 64 x 64 the one-pass gradient on two slots takes 287 us (pixel-pair kernel: 312 us, one slot: 307 us) while the
// forward pass stays on the pixel-pair kernel (153 vs 177 us: the 32-row tile evaluates 6 component slots for 5 components);
// at 5 x 128 x 32 x 32 three launches on two slots (93.5 us) beat the one-launch step (99.1 us).  VAEMDL_M5_SLOTS=1: one slot.
static bool m5_two_slots() {
  static const bool two = [] { const char* e = getenv("VAEMDL_M5_SLOTS"); return !(e && e[0] == '1'); }();
  return two;
}

struct TilePlan {  // how the forward grid split the tile range: what the per-image reduction needs to know
  long long total_warps = 0, tw_base = 0, tw_rem = 0;
  int K = 0, PPT = 0;
};

template <int MC, int LPP, bool BWD, int NSLOT, int MAXT, int AR, int PD = 0, bool ST = false>
static int launch_tiled_shape(ModlArgs a, int warps, cudaStream_t st, TilePlan* plan) {
  using T = Tile<MC, LPP>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  // (PD = 2: the slots hold bfloat16; ST with two slots: one (S, SW) strip per slot -- the layout of tile_body)
  const size_t per_warp = (static_cast<size_t>(NSLOT) * (PD == 2 ? T::TILE_F / 2 : T::TILE_F) + ((BWD && !ST) ? T::AUX_F : 0) +
                           (ST ? (NSLOT > 1 ? NSLOT : 1) * 2 * T::PPT : 0)) * 4 + NSLOT * 8;
  if (warps > MAXT / 32) warps = MAXT / 32;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = BWD ? pick_warps(a.num_tiles, di.sm_count, warps) : pick_warps_fwd(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  if (smem > static_cast<size_t>(di.max_smem_optin)) return VAEMDL_EUNSUPPORTED;
  auto kern = modl_tile_kernel<MC, LPP, BWD, NSLOT, MAXT, AR, PD, ST>;
  // the function attribute and the occupancy query cost several microseconds of host time: once per (device, shape)
  static std::mutex mu;
  static int c_dev = -1, c_warps = -1, c_ctas = 1;
  int ctas_per_sm;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || c_warps != warps) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      int n = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem);
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_warps = warps;
      c_ctas = n < 1 ? 1 : n;
    }
    ctas_per_sm = c_ctas;
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = static_cast<long long>(di.sm_count) * ctas_per_sm;  // persistent: every CTA resident
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  if (a.partial && !partials_fit(a.n_px / a.HW, a.K)) return VAEMDL_EWORKSPACE;  // (before anything is enqueued)
  a.small = a.n_px < (1ll << 31) - 64;  // image / pixel indices fit 32 bits
  apply_l2_opt(a, total_warps, T::TILE_B / (PD ? 2 : 1));
  if (PD == 2 && NSLOT > 1 && a.keep_tiles > 0 && a.keep_tiles < NSLOT) a.keep_tiles = NSLOT;
  if (BWD && NSLOT > 1) {
    // two-slot gradient kernels: before which component pair of a tile the other slot is refilled (VAEMDL_REFILL2=<pair>, A/B).
    // Measured (tools/ab_refill2.sh, profiles/r02z_refill2.txt): n_mix 5 gains 1 % (291 -> 288 us) and 2.4 % at 5 x 128 x 32 x 32 with
    // pair 1 instead of 0; the bfloat16 tiles are indifferent.
    static const int refill2 = [] { const char* e = getenv("VAEMDL_REFILL2"); return e ? atoi(e) : 1; }();
    a.tm_refill = refill2 < T::NPAIR ? refill2 : T::NPAIR - 1;
  }
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = T::PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

// the two-pass gradient kernel on tensor memory (modl_tm.cuh): float32 parameters, one slot per warp, twelve warps per CTA.
// Measured on B200 (tools/ab_tm.sh, profiles/r02y_*): headline gradient kernel 280 -> 271 us, n_mix 20 at the same size 613 -> 551 us,
// BASELINE configs[0] 56 -> 48 us.  VAEMDL_TM=0 keeps the shared-memory kernel (A/B).
static bool tm_enabled() {  // (re-read on every call: the tests compare both kernels inside one process)
  const char* e = getenv("VAEMDL_TM");
  return !(e && e[0] == '0');
}
template <int MC, int LPP, int AR>
static int launch_tiled_tm(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  using T = Tile<MC, LPP>;
  constexpr int MAXT = kTmWarps * 32;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = static_cast<size_t>(T::TILE_F) * 4 + 8;
  int warps = kTmWarps;
  while (warps > 1 && warps * per_warp + 16 > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp + 16;
  auto kern = modl_tile_tm_kernel<MC, LPP, MAXT, AR>;
  static std::mutex mu;
  static int c_dev = -1;
  static size_t c_smem = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || smem > c_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = di.sm_count;  // persistent, ONE CTA per SM: the CTA owns all 512 TMEM columns
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  a.small = a.n_px < (1ll << 31) - 64;
  apply_l2_opt(a, total_warps, T::TILE_B);
  {
    // where in the first pass the staging slot is refilled: as late as possible -- before the LAST component pair -- measured
    // best everywhere (VAEMDL_TM_REFILL=<pair> / NPAIR = after the pass: headline gradient kernel 309 / 305 / 302 / 289 / 271 / 285 us
    // for 0 ... 5): the gradient tile the slot still holds drains slowly while the memory system is saturated with writes
    // (a problem of a few tiles per warp does not saturate the write path: there one pair earlier is 1-3 % ahead, tools/ab_tm.sh)
    static const int refill = [] { const char* e = getenv("VAEMDL_TM_REFILL"); return e ? atoi(e) : -1; }();
    const bool few_tiles = a.num_tiles < total_warps * 12 && T::NPAIR >= 3;
    a.tm_refill = refill < 0 ? T::NPAIR - (few_tiles ? 2 : 1) : (refill < T::NPAIR ? refill : T::NPAIR);
  }
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = T::PPT;
  }
  return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
}

template <int MC, int LPP, bool BWD, int AR>
static int launch_tiled(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  if (a.bf16) {  // bfloat16 parameters: x-conditioned class only
    if constexpr (AR == 0) {
      // Aligned component pairs (n_mix 10 / 20 / 30): the tile stays bfloat16 in shared memory, two slots per warp, pairs
      // widened as they are read (PD = 2).  The backward pass takes that route when the forward pass left its per-pixel
      // sums (one-pass gradient, rounded once); without them -- and for n_mix 5 -- the tile is widened in place (PD = 1).
      static const bool direct_off = getenv("VAEMDL_BF16_WIDEN") != nullptr;  // A/B: always widen in place
      if constexpr (Tile<MC, LPP>::ALIGNED) {
        if (!direct_off) {
          if constexpr (!BWD) return launch_tiled_shape<MC, LPP, false, 2, 512, 0, 2>(a, 16, st, plan);
          if constexpr (BWD) {
            if (a.pix_stats) return launch_tiled_shape<MC, LPP, true, 2, 512, 0, 2, true>(a, 16, st, plan);
          }
        }
      }
      a.pix_stats = nullptr;
      return launch_tiled_shape<MC, LPP, BWD, 1, 512, 0, 1>(a, tune_shape(BWD, Shape{1, 16}).warps, st, plan);
    } else {
      return VAEMDL_EUNSUPPORTED;
    }
  }
  // 1 slot x 16 warps: measured best on B200 for every M (profiles/r01_tune_shapes.txt); latency is hidden by the 4
  // warps per scheduler rather than by a second slot per warp
  const Shape sh = tune_shape(BWD, Shape{1, 16});
  // (2 slots x <= 8 warps was 20-30 % behind at every n_mix, profiles/r01_tune_shapes.txt: that instantiation is gone)
  if constexpr (MC == 5 && LPP == 1 && AR == 0) {
    // n_mix 5: two 6.4 KB slots per warp (m5_two_slots)
    if (m5_two_slots()) {
      if constexpr (BWD) {
        if (a.pix_stats) return launch_tiled_shape<5, 1, true, 2, 512, 0, 0, true>(a, 16, st, plan);
        return launch_tiled_shape<5, 1, true, 2, 512, 0>(a, 16, st, plan);
      } else {
        return launch_tiled_shape<5, 1, false, 2, 512, 0>(a, 16, st, plan);
      }
    }
  }
  if constexpr (BWD) {
    if (a.pix_stats) return launch_tiled_shape<MC, LPP, true, 1, 512, AR, 0, true>(a, sh.warps, st, plan);  // one-pass gradient
    if constexpr (tm_supported<MC, LPP>()) {
      // two-pass gradient with the tile in tensor memory
      if (tm_enabled() && !getenv("VAEMDL_TUNE")) return launch_tiled_tm<MC, LPP, AR>(a, st, plan);
    }
  }
  return launch_tiled_shape<MC, LPP, BWD, 1, 512, AR>(a, sh.warps, st, plan);
}

// the fused one-launch step (modl_step_kernel): backward shared-memory footprint, cooperative launch
template <int MC, int LPP, int AR>
static int launch_step(ModlArgs a, StepFinish f, long long n_img, cudaStream_t st) {
  using T = Tile<MC, LPP>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = (static_cast<size_t>(T::TILE_F) + 2 * T::PPT) * 4 + 8;  // (ST: the tile's (S, SW) pairs, no aux strip)
  int warps = 16;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_step_kernel<MC, LPP, AR>;
  static std::mutex mu;
  static int c_dev = -1;
  static size_t c_smem = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || smem > c_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = di.sm_count;  // one CTA per SM: every CTA is resident, as the grid barrier requires
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  if (!partials_fit(n_img, a.K)) return VAEMDL_EWORKSPACE;
  a.small = a.n_px < (1ll << 31) - 64;
  a.reverse = 1;
  {  // VAEMDL_L2S="keep=<MB>,hint=<bits>" (A/B, tools/ab_l2s.sh): L2 policy inside the one-launch step -- forward loads of the
     // last <MB> of every run evict_last, the rest evict_first; hint bit 0 / 1: evict_first on the backward pass's loads / stores.
     // Measured: 94.1-97.7 us at BASELINE configs[0] whatever the setting, so the default is no hint at all.
    static const int keep_mb = [] { const char* e = getenv("VAEMDL_L2S"); const char* q = e ? strstr(e, "keep=") : nullptr; return q ? atoi(q + 5) : 0; }();
    static const int hint = [] { const char* e = getenv("VAEMDL_L2S"); const char* q = e ? strstr(e, "hint=") : nullptr; return q ? atoi(q + 5) : 0; }();
    a.keep_tiles = keep_mb > 0 ? static_cast<int>((static_cast<long long>(keep_mb) << 20) / (total_warps * T::TILE_B)) : 0;
    if (keep_mb > 0 && a.keep_tiles < 1) a.keep_tiles = 1;
    a.bwd_hint = hint;
  }
  f.geom = PartialGeom{a.partial, a.tw_base, a.tw_rem, a.K, T::PPT, a.HW};
  StepArgs sa{a, f};
  void* args[] = {&sa};
  return cuda_rc(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(static_cast<unsigned>(grid)),
                                             dim3(static_cast<unsigned>(warps * 32)), args, smem, st));
}

// pixel-pair kernel, n_mix = M (1 <= M <= 9): one slot per warp, as many warps as shared memory allows (<= 16)
template <int M, bool BWD, int AR>
static int launch_pp(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  using T = TilePP<M>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = (static_cast<size_t>(T::TILE_F) + (BWD ? T::AUX_F : 0)) * 4 + 8;
  int warps = tune_shape(BWD, Shape{1, 16}).warps;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = BWD ? pick_warps(a.num_tiles, di.sm_count, warps) : pick_warps_fwd(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_pp_kernel<M, BWD, 512, AR>;
  static std::mutex mu;
  static int c_dev = -1, c_warps = -1, c_ctas = 1;
  int ctas_per_sm;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || c_warps != warps) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      int n = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem);
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_warps = warps;
      c_ctas = n < 1 ? 1 : n;
    }
    ctas_per_sm = c_ctas;
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = static_cast<long long>(di.sm_count) * ctas_per_sm;
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  if (a.partial && !partials_fit(a.n_px / a.HW, a.K)) return VAEMDL_EWORKSPACE;  // (before anything is enqueued)
  a.small = a.n_px < (1ll << 31) - 64;
  apply_l2_opt(a, total_warps, T::TILE_B);
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = T::PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

// ---- run-time tile geometry for modl_rt_kernel -------------------------------------------------------------------------
// For a given n_mix pick (LPP, rot): score = lane efficiency / sqrt(mean bank-conflict degree of the scalar parameter loads)
// * sqrt(min(1, warps that fit / 16)).  Evaluated once per n_mix (a few thousand integer operations) and cached.
struct RtPlan {
  int MC = 0, LPP = 0, PPT = 0, rot = 0;
};
// score = lane efficiency * MC / (MC + 3) * sqrt(min(1, warps that fit / 16)) / (mean bank-conflict degree)^(1/4):
// measured on B200 (tools/rt_sweep.py), long component chunks per lane win over many lanes per pixel (the per-tile work of
// a lane -- logit max, group reductions, log, index bookkeeping -- is amortised over MC components), as long as 16 warps
// still fit in shared memory.  Evaluated once per (n_mix, direction) and cached.
static RtPlan rt_plan_compute(int M, bool bwd, bool bf16) {
  RtPlan best;
  double best_score = -1.0;
  for (int LPP = 1; LPP <= 16; ++LPP) {
    const int MC = (M + LPP - 1) / LPP;
    if (MC > 13) continue;
    int PPT = 32 / LPP;
    // tile bytes = PPT * 40 * M (float32) or PPT * 20 * M (bfloat16) must be a multiple of 16
    while (PPT >= 1 && ((PPT * M) & (bf16 ? 3 : 1))) --PPT;
    if (PPT < 1) continue;
    const int NP = (MC + 1) / 2;
    const double eff = static_cast<double>(M) / (2.0 * NP * LPP) * (static_cast<double>(PPT) * LPP / 32.0);
    const double per_warp = (PPT * 10.0 * M + (bwd ? PPT * M : 0)) * 4.0 + 24.0;
    double occ = 227.0 * 1024.0 / per_warp / 16.0;
    if (occ > 1.0) occ = 1.0;
    for (int rot = 0; rot < (NP > 1 ? 4 : 1); ++rot) {
      long long tot = 0, cnt = 0;
      for (int pr = 0; pr < NP; ++pr) {
        for (int half = 0; half < 2; ++half) {
          // the ten parameter planes of a row are k*M apart: evaluate the first one
          int per_bank_addr[32][32];
          int per_bank_n[32] = {0};
          for (int lane = 0; lane < 32; ++lane) {
            const int pq = lane / LPP, sub = lane % LPP;
            if (pq >= PPT) continue;
            const int m0 = sub * MC, m_end = m0 + MC < M ? m0 + MC : M;
            const int prr = (pr + rot * pq) % NP;
            int m = m0 + 2 * prr + half;
            if (m >= m_end) m = m0 < M ? m0 : 0;
            const int addr = pq * 10 * M + m;
            const int bnk = addr & 31;
            bool seen = false;
            for (int q = 0; q < per_bank_n[bnk]; ++q) seen = seen || per_bank_addr[bnk][q] == addr;
            if (!seen) per_bank_addr[bnk][per_bank_n[bnk]++] = addr;
          }
          int deg = 1;
          for (int bnk = 0; bnk < 32; ++bnk) deg = per_bank_n[bnk] > deg ? per_bank_n[bnk] : deg;
          tot += deg;
          ++cnt;
        }
      }
      const double conflict = static_cast<double>(tot) / static_cast<double>(cnt);
      const double score = eff * MC / (MC + 3.0) * sqrt(occ) / sqrt(sqrt(conflict));
      if (score > best_score) {
        best_score = score;
        best = RtPlan{MC, LPP, PPT, rot};
      }
    }
  }
  return best;
}
static RtPlan rt_plan(int M, bool bwd, bool bf16 = false) {
  const char* env = getenv("VAEMDL_RT");  // "LPP:rot" overrides the choice (tuning sweeps; re-read on every call)
  int l = 0, r = 0;
  if (env && sscanf(env, "%d:%d", &l, &r) == 2 && l >= 1 && l <= 32 && (M + l - 1) / l <= 16) {
    int ppt = 32 / l;
    while (ppt >= 1 && ((ppt * M) & (bf16 ? 3 : 1))) --ppt;
    if (ppt >= 1) return RtPlan{(M + l - 1) / l, l, ppt, r};
  }
  static std::mutex mu;
  static RtPlan cache[4][VAEMDL_MAX_MIX + 1];
  std::lock_guard<std::mutex> lock(mu);
  RtPlan& c = cache[(bwd ? 1 : 0) + (bf16 ? 2 : 0)][M];
  if (c.LPP == 0) c = rt_plan_compute(M, bwd, bf16);
  return c;
}

template <bool BWD, int AR, bool AL, int PD>
static int launch_rt_al(ModlArgs a, const RtPlan& rp, cudaStream_t st, TilePlan* plan);

template <bool BWD, int AR>
static int launch_rt(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  const RtPlan rp = rt_plan(a.M, BWD, a.bf16 != 0);
  if (rp.LPP == 0) return VAEMDL_EUNSUPPORTED;
  static const bool no_al = getenv("VAEMDL_RT_NOAL") != nullptr;  // A/B: scalar shared accesses everywhere
  const bool al = (a.M % 2 == 0) && (rp.MC % 2 == 0 || rp.LPP == 1) && !no_al;
  if (a.bf16) {
    if constexpr (AR == 0)
      return al ? launch_rt_al<BWD, 0, true, 1>(a, rp, st, plan) : launch_rt_al<BWD, 0, false, 1>(a, rp, st, plan);
    else
      return VAEMDL_EUNSUPPORTED;
  }
  return al ? launch_rt_al<BWD, AR, true, 0>(a, rp, st, plan) : launch_rt_al<BWD, AR, false, 0>(a, rp, st, plan);
}

template <bool BWD, int AR, bool AL, int PD>
static int launch_rt_al(ModlArgs a, const RtPlan& rp, cudaStream_t st, TilePlan* plan) {
  a.rt_MC = rp.MC;
  a.rt_LPP = rp.LPP;
  a.rt_PPT = rp.PPT;
  a.rt_rot = rp.rot;
  const int tile_f = rp.PPT * 10 * a.M;
  a.rt_warp_f = (tile_f + (BWD ? rp.PPT * a.M : 0) + 3) & ~3;  // every warp's slot stays 16-byte aligned
  a.num_tiles = (a.n_px + rp.PPT - 1) / rp.PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = static_cast<size_t>(a.rt_warp_f) * 4 + 8;
  int warps = tune_shape(BWD, Shape{1, 16}).warps;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  if (warps * per_warp > static_cast<size_t>(di.max_smem_optin)) return VAEMDL_EUNSUPPORTED;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_rt_kernel<BWD, AR, AL, PD>;
  static std::mutex mu;
  static int c_dev = -1;
  static size_t c_smem = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || smem > c_smem) {  // the attribute is a maximum: raise it when a larger footprint shows up
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = di.sm_count;  // persistent, one CTA per SM (__launch_bounds__(512, 1))
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, rp.PPT, a.tw_base);
  if (a.partial && !partials_fit(a.n_px / a.HW, a.K)) return VAEMDL_EWORKSPACE;  // (before anything is enqueued)
  a.small = a.n_px < (1ll << 31) - 64;
  apply_l2_opt(a, total_warps, static_cast<long long>(tile_f) * (PD ? 2 : 4));
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = rp.PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

// n_mix 1..9 run on the pixel-pair kernel.  n_mix = 5 also has a component-pair instantiation with 32-row tiles, which
// is a little faster while the problem is so small that a warp only sees a handful of tiles (measured: 112 vs 117 us
// per step at 5 x 128 x 32 x 32, 345 vs 314 us backward at 16 x 64 x 64 x 64).
// Tile instantiations beyond n_mix 5 / 10 / 20 / 30 are compiled in their own translation units (modl_tiles_{a,b,c,d}.cu),
// x-conditioned classes only (AR = 0); the un-conditioned class (utils/mdl_plain.py) runs those n_mix on the run-time tiled
// kernel.  n_mix -> (components per lane) x (lanes per pixel):
//   a:  8 -> 8x1, 12 -> 6x2, 16 -> 8x2, 24 -> 6x4        b: 32 -> 8x4, 40 -> 10x4, 64 -> 8x8
//   c:  6 -> 6x1, 14 -> 14x1, 18 -> 6x3, 28 -> 14x2      d: 36 -> 12x3, 48 -> 12x4, 50 -> 10x5, 56 -> 14x4, 60 -> 12x5
int launch_tiled_extra_a(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan);
int launch_tiled_extra_b(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan);
int launch_tiled_extra_c(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan);
int launch_tiled_extra_d(bool bwd, const ModlArgs& a, cudaStream_t st, TilePlan* plan);
static int extra_tile_group(int M) {  // 1..4 = a..d, 0 = no extra tile instantiation
  switch (M) {
    case 8: case 12: case 16: case 24: return 1;
    case 32: case 40: case 64: return 2;
    case 6: case 14: case 18: case 28: return 3;
    case 36: case 48: case 50: case 56: case 60: return 4;
    default: return 0;
  }
}
static int extra_tile_ppt(int M, int AR) {  // pixels per tile of the extra instantiation serving n_mix M (0: none)
  if (AR != 0 || getenv("VAEMDL_NO_EXTRA_TILES")) return 0;
  switch (M) {
    case 6: case 8: case 14: return 32;       // one lane per pixel
    case 12: case 16: case 28: return 16;     // two
    case 18: case 36: return 10;              // three
    case 24: case 32: case 40: case 48: case 56: return 8;  // four
    case 50: case 60: return 6;               // five
    case 64: return 4;                        // eight
    default: return 0;
  }
}

static bool use_pixel_pairs(int M, long long n_px, bool bf16 = false) {
  const char* env = getenv("VAEMDL_PP");  // "0" / "1" force the choice for n_mix = 5 (A/B measurements, tests)
  if (M < 1 || M > 9 || bf16) return false;  // (bfloat16 parameters: tile<5,1> for n_mix = 5, the run-time kernel otherwise)
  if (M != 5) return true;
  if (env && (env[0] == '0' || env[0] == '1')) return env[0] == '1';
  return n_px >= 64ll * 148 * 16 * 6;
}

// Rotated component-pair order against shared-memory bank conflicts: which tile shapes carry it is decided at compile time
// by the bank model in modl_tile.cuh (Tile::ROT_KIND).  VAEMDL_ROT=0 switches it off (A/B).
static int pair_rot_on(int /*M*/) {
  static const int on = [] {
    const char* e = getenv("VAEMDL_ROT");
    return !(e && e[0] == '0');
  }();
  return on;
}

static int spread_runs() {
  const char* e = getenv("VAEMDL_SPREAD");  // "0": CTA-major run numbering (A/B)
  return !(e && e[0] == '0');
}

// Which separately launched kernels exchange per-pixel mixture sums between the forward and the backward pass
// (ModlArgs::pix_stats -> one-pass gradient).  Measured on B200 (tools/fused_probe.py, VAEMDL_NO_STATS A/B): n_mix = 30 gains
// 1-8 % and n_mix = 5 (32-row tiles) 4 %, n_mix = 10 / 20 LOSE 6-8 % -- with the second pass gone the warps spend 23 % of
// their samples waiting for the next tile instead of 10 % -- so those keep the two-pass kernel.  VAEMDL_STATS=all / none
// overrides (A/B).  The one-launch step (modl_step_kernel) always exchanges the sums.
static bool stats_off() {
  const char* e = getenv("VAEMDL_STATS");
  return getenv("VAEMDL_NO_STATS") != nullptr || (e && e[0] == 'n');
}
static bool stats_supported(int M, long long n_px, bool bf16, int AR = 0) {
  if (stats_off()) return false;
  if (M == 5 && !bf16 && AR == 0 && m5_two_slots()) return true;  // pixel-pair forward writes them, two-slot tile backward reads them
  if (use_pixel_pairs(M, n_px, bf16)) return false;
  if (bf16)  // the direct bf16 backward pass (aligned tile instantiations) needs them
    return (M == 10 || M == 20 || M == 30 || (AR == 0 && extra_tile_ppt(M, 0))) && getenv("VAEMDL_BF16_WIDEN") == nullptr;
  if (AR == 0 && extra_tile_ppt(M, 0)) return false;  // (the extra float32 tiles keep the two-pass gradient)
  const char* e = getenv("VAEMDL_STATS");
  if (e && e[0] == 'a') return M == 5 || M == 10 || M == 20 || M == 30;
  // n_mix 30: below ~1.2 M pixel-samples the two-pass gradient on tensor memory is ahead of the one-pass kernel (5 x 64 x 32 x 32:
  // step 229 vs 237 us, 5 x 128 x 32 x 32: 419 vs 440 us, 16 x 16 x 64 x 64: 670 vs 678 us; level from 1.5 M on; tools/ab_m30.sh)
  if (M == 30) return tm_enabled() ? n_px >= 1200000 : true;
  return M == 5;
}

template <bool BWD, int AR>
static int launch_modl(ModlArgs a, cudaStream_t st, TilePlan* plan = nullptr) {
  a.plain = AR;
  a.spread = spread_runs();
  a.pair_rot = pair_rot_on(a.M);
  if (!stats_supported(a.M, a.n_px, a.bf16 != 0, AR)) a.pix_stats = nullptr;
  // n_mix 5 with the forward pass's sums at hand: the one-pass gradient on the two-slot 32-row tile, whatever the size
  const bool m5_tile_bwd = BWD && AR == 0 && a.M == 5 && !a.bf16 && a.pix_stats && m5_two_slots();
  if (const int ppt = extra_tile_ppt(a.M, AR)) {
    (void)ppt;
    switch (extra_tile_group(a.M)) {
      case 1: return launch_tiled_extra_a(BWD, a, st, plan);
      case 2: return launch_tiled_extra_b(BWD, a, st, plan);
      case 3: return launch_tiled_extra_c(BWD, a, st, plan);
      default: return launch_tiled_extra_d(BWD, a, st, plan);
    }
  }
  if (!m5_tile_bwd && use_pixel_pairs(a.M, a.n_px, a.bf16 != 0)) {
    switch (a.M) {
      case 1: return launch_pp<1, BWD, AR>(a, st, plan);
      case 2: return launch_pp<2, BWD, AR>(a, st, plan);
      case 3: return launch_pp<3, BWD, AR>(a, st, plan);
      case 4: return launch_pp<4, BWD, AR>(a, st, plan);
      case 5: return launch_pp<5, BWD, AR>(a, st, plan);
      case 6: return launch_pp<6, BWD, AR>(a, st, plan);
      case 7: return launch_pp<7, BWD, AR>(a, st, plan);
      case 8: return launch_pp<8, BWD, AR>(a, st, plan);
      case 9: return launch_pp<9, BWD, AR>(a, st, plan);
    }
  }
  switch (a.M) {
    case 5:
      return launch_tiled<5, 1, BWD, AR>(a, st, plan);
    case 10:
      return launch_tiled<10, 1, BWD, AR>(a, st, plan);
    case 20:
      return launch_tiled<10, 2, BWD, AR>(a, st, plan);
    case 30:
      return launch_tiled<10, 3, BWD, AR>(a, st, plan);
    default: {
      static const bool force_generic = getenv("VAEMDL_GENERIC") != nullptr;  // A/B against the one-thread-per-pixel kernel
      if (!force_generic) return launch_rt<BWD, AR>(a, st, plan);
      if (a.bf16) return VAEMDL_EUNSUPPORTED;
      const DeviceInfo& di = device_info();
      long long blocks = (a.n_px + 127) / 128;
      const long long cap = static_cast<long long>(di.sm_count) * 8;
      if (blocks > cap) blocks = cap;
      modl_generic_kernel<BWD><<<static_cast<unsigned>(blocks), 128, 0, st>>>(a);
      return cuda_rc(cudaGetLastError());
    }
  }
}

static int tile_ppt(int M, long long n_px, bool bf16 = false, int AR = 0) {
  if (const int ppt = extra_tile_ppt(M, AR)) return ppt;
  if (use_pixel_pairs(M, n_px, bf16)) return 64;
  switch (M) {
    case 5:
    case 10:
      return 32;
    case 20:
      return 16;
    case 30:
      return 10;
    default:
      if (getenv("VAEMDL_GENERIC")) return 0;  // one-thread-per-pixel kernel: atomics
      return rt_plan(M, false, bf16).PPT;
  }
}

// bin geometry of the class (utils/discretized_logistic.py:10-21); the x-conditioned classes are fixed to the default
constexpr double kMaxH = 27.0;  // largest half bin width in units of the narrowest scale the linear-domain product carries
struct BinGeom {
  float low = -1.0f, high = 1.0f, levels = 256.0f;
  bool is_default() const { return low == -1.0f && high == 1.0f && levels == 256.0f; }
};
static int set_bins(ModlArgs& a, const BinGeom& g, int AR) {
  if (!(g.levels > 1.0f) || !(g.high > g.low)) return VAEMDL_EINVAL;
  if (AR == 0 && !g.is_default()) return VAEMDL_EUNSUPPORTED;
  // The kernels multiply the three sub-pixel terms in the linear domain, each a ratio whose denominator is >= exp(-h),
  // h = exp(-logscale) * dx <= e^7 * dx (log-scales are clamped at -7, utils/mdl_plain.py:150): the product of three stays
  // inside the float32 range while h <= kMaxH, i.e. bin width <= 2 * kMaxH / e^7 = 0.049 (levels >= 42 on [-1, 1]).
  if (!g.is_default() && (static_cast<double>(g.high) - g.low) / (static_cast<double>(g.levels) - 1.0) / 2.0 * exp(7.0) > kMaxH)
    return VAEMDL_EUNSUPPORTED;
  a.low = g.low;
  a.high = g.high;
  const double width = (static_cast<double>(g.high) - static_cast<double>(g.low)) / (static_cast<double>(g.levels) - 1.0);
  a.width = static_cast<float>(width);          // utils/discretized_logistic.py:18
  a.dx = static_cast<float>(width / 2.0);       // :21
  // h = exp(-ls) * dx >= kHSmall  <=>  ls <= log(dx / kHSmall); nudged up so the cut errs towards the exact exp(-h)
  a.ls_narrow = g.is_default() ? kLsNarrow : static_cast<float>(log(width / 2.0 / static_cast<double>(kHSmall)) + 1e-4);
  return VAEMDL_OK;
}

static int check_common(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                        int x_batch, int H, int W, int M) {
  if (!params || !x) return VAEMDL_EINVAL;
  if (n_img <= 0 || x_batch <= 0 || H <= 0 || W <= 0) return VAEMDL_EINVAL;
  if (x_dtype != VAEMDL_X_F32 && x_dtype != VAEMDL_X_U8) return VAEMDL_EINVAL;
  if (x_range != VAEMDL_RANGE_UNIT && x_range != VAEMDL_RANGE_SYM) return VAEMDL_EINVAL;
  if (x_dtype == VAEMDL_X_U8 && x_range != VAEMDL_RANGE_UNIT) return VAEMDL_EINVAL;
  if (edge_mode != VAEMDL_EDGE_MDL && edge_mode != VAEMDL_EDGE_OPENAI) return VAEMDL_EINVAL;
  if (M < 1 || M > VAEMDL_MAX_MIX) return VAEMDL_EUNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(params) & 15u) return VAEMDL_EALIGN;
  return VAEMDL_OK;
}

}  // namespace vaemdl


namespace vaemdl {
template <int AR>
static int modl_fwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                         int x_batch, int H, int W, int M, float* lp_pixel, float* ll_image, double* ll_image_f64,
                         const IwaeOut& iw, void* workspace, size_t workspace_bytes, cudaStream_t st, int bf16 = 0,
                         float* pix_stats = nullptr, const BinGeom& bins = BinGeom{}) {
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  const bool iwae = iw.S > 0;
  if (!lp_pixel && !ll_image && !ll_image_f64 && !iwae) return VAEMDL_EINVAL;
  if (iwae && static_cast<long long>(iw.S) * iw.B != n_img) return VAEMDL_EINVAL;
  const bool want_ll = ll_image || ll_image_f64 || iwae;
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.lp_pixel = lp_pixel;
  a.HW = H * W;
  a.n_px = n_img * a.HW;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.bf16 = bf16;
  a.pix_stats = reinterpret_cast<float2*>(pix_stats);
  if ((rc = set_bins(a, bins, AR))) return rc;
  const int ppt = tile_ppt(M, a.n_px, bf16 != 0, AR);
  const bool use_partials = want_ll && ppt > 0 && a.HW >= ppt;
  char* ws = static_cast<char*>(workspace);
  size_t tail_off = 0;
  if (want_ll) {
    const size_t need = vaemdl_modl_workspace_bytes(n_img, H, W);
    if (!workspace || workspace_bytes < need) return VAEMDL_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
    tail_off = partial_elems(n_img) * sizeof(double);
    if (use_partials) {
      a.partial = reinterpret_cast<double*>(ws);
    } else {
      a.ll_atomic = ll_image_f64 ? ll_image_f64 : reinterpret_cast<double*>(ws);
      cudaError_t e = cudaMemsetAsync(a.ll_atomic, 0, sizeof(double) * n_img, st);
      if (e != cudaSuccess) return cuda_rc(e);
    }
  }
  unsigned* counter = want_ll ? reinterpret_cast<unsigned*>(ws + tail_off + static_cast<size_t>(n_img) * sizeof(double)) : nullptr;
  if (iwae && use_partials && iw.elbo) a.zero_me = counter;
  TilePlan plan;
  rc = launch_modl<false, AR>(a, st, &plan);
  if (rc) return rc;
  if (use_partials) {
    const PartialGeom geom{a.partial, plan.tw_base, plan.tw_rem, plan.K, plan.PPT, a.HW};
    return finish_partials(geom, n_img, ll_image, ll_image_f64, iw, reinterpret_cast<double*>(ws + tail_off), counter, st);
  }
  if (!want_ll) return VAEMDL_OK;
  // float64 atomics route (any-M kernel, images smaller than a tile)
  if (ll_image) {
    cast_f64_f32_kernel<<<static_cast<unsigned>((n_img + 255) / 256), 256, 0, st>>>(a.ll_atomic, ll_image, n_img);
    rc = cuda_rc(cudaGetLastError());
  }
  if (rc || !iwae) return rc;
  rc = vaemdl_iwae_tail(nullptr, a.ll_atomic, iw.extra, iw.S, iw.B, iw.B_total, iw.log_w, iw.lme_b, iw.elbo, iw.g_ll, st);
  if (rc || !iw.elbo) return rc;
  return peer_push(iw.peer, iw.elbo, st);
}
}  // namespace vaemdl

namespace vaemdl {
template <int AR>
static int modl_bwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                         int x_batch, int H, int W, int M, const float* g_image, const float* g_pixel, float* dparams,
                         cudaStream_t st, int bf16 = 0, const float* pix_stats = nullptr, const BinGeom& bins = BinGeom{}) {
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  if (!dparams || (!g_image && !g_pixel)) return VAEMDL_EINVAL;
  if (reinterpret_cast<uintptr_t>(dparams) & 15u) return VAEMDL_EALIGN;
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.g_image = g_image;
  a.g_pixel = g_pixel;
  a.dparams = dparams;
  a.HW = H * W;
  a.n_px = n_img * a.HW;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.bf16 = bf16;
  a.pix_stats = reinterpret_cast<float2*>(const_cast<float*>(pix_stats));
  if (pix_stats && (reinterpret_cast<uintptr_t>(pix_stats) & 7u)) return VAEMDL_EALIGN;
  if ((rc = set_bins(a, bins, AR))) return rc;
  return launch_modl<true, AR>(a, st);
}

template <int AR>
static int modl_iwae_fwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                              long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                              float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                              void* workspace, size_t workspace_bytes, cudaStream_t st, int bf16 = 0,
                              float* pix_stats = nullptr, const BinGeom& bins = BinGeom{}) {
  if (S <= 0 || B <= 0 || B_total < 0) return VAEMDL_EINVAL;
  if (pix_stats && (reinterpret_cast<uintptr_t>(pix_stats) & 7u)) return VAEMDL_EALIGN;
  if (elbo && !lme_b) return VAEMDL_EINVAL;
  IwaeOut iw;
  iw.S = S;
  iw.B = B;
  iw.B_total = B_total;
  iw.extra = extra;
  iw.log_w = log_w;
  iw.lme_b = lme_b;
  iw.elbo = elbo;
  iw.g_ll = g_ll;
  iw.peer = take_peer();
  return modl_fwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, static_cast<long long>(S) * B, x_batch, H, W, M, nullptr,
                           ll_image, ll_image_f64, iw, workspace, workspace_bytes, st, bf16, pix_stats, bins);
}
}  // namespace vaemdl

namespace vaemdl {
// VAEMDL_FUSED = "1": the one-launch cooperative step (modl_step_kernel) whenever the shape is eligible; otherwise three
// launches.  Round 1 took the one-launch step for small training shapes (101.6 -> 96.6 us at BASELINE configs[0]).  Since the
// gradient kernel forms its first tile's derivatives while the finish kernel runs (tile_body LATE_G) the three launches are
// AHEAD at every shape measured (5 x 64 x 32 x 32: n_mix 10 92.5 vs 96.6 us, n_mix 20 154 vs 167 us, n_mix 30 235 vs 234 us;
// n_mix 5 x batch 128 on two slots per warp 93.2 vs 99.1 us), so the cooperative kernel is now opt-in.
static int fused_mode() {
  const char* e = getenv("VAEMDL_FUSED");
  if (!e) return -1;
  return e[0] == '0' ? 0 : 1;
}
static bool fused_eligible(int S, long long n_px, int HW, int M, int AR = 0) {
  const int mode = fused_mode();
  if (mode == 0 || S > 32) return false;
  if (use_pixel_pairs(M, n_px)) return false;
  if (M != 5 && M != 10 && M != 20 && M != 30) return false;
  const int ppt = tile_ppt(M, n_px);
  if (HW < ppt) return false;
  static int coop = -1;
  if (coop < 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
    coop = v;
  }
  if (!coop) return false;
  (void)AR;
  return mode == 1;
}

// One IWAE step of the observation model: forward, per-image sums, log-mean-exp, elbo, softmax weights, parameter gradient.
// One cooperative launch when the shape is eligible, else forward + finish + backward (3 launches).
template <int AR>
static int modl_iwae_step_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                               long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                               float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                               float* dparams, void* workspace, size_t workspace_bytes, cudaStream_t st, int* launches,
                               const BinGeom& bins = BinGeom{}) {
  if (S <= 0 || B <= 0 || B_total < 0 || !lme_b || !g_ll || !dparams) return VAEMDL_EINVAL;
  const long long n_img = static_cast<long long>(S) * B;
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(dparams) & 15u) return VAEMDL_EALIGN;
  const int HW = H * W;
  const long long n_px = n_img * HW;
  const size_t need = vaemdl_modl_workspace_bytes(n_img, H, W);
  if (!workspace || workspace_bytes < need) return VAEMDL_EWORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
  char* ws = static_cast<char*>(workspace);
  // per-pixel mixture sums handed from the forward to the backward pass (one-pass gradient): behind the forward
  // workspace, when the caller sized the buffer with vaemdl_modl_step_workspace_bytes and the kernels support it
  const bool room = workspace_bytes >= need + static_cast<size_t>(n_px) * sizeof(float2);
  const bool fused = room && !stats_off() && fused_eligible(S, n_px, HW, M, AR);  // (implies n_mix in {5, 10, 20, 30})
  float* stats = nullptr;
  if (room && (fused || stats_supported(M, n_px, false, AR))) stats = reinterpret_cast<float*>(ws + need);
  if (!fused) {
    if (launches) *launches = 3;
    rc = modl_iwae_fwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                                ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes, st, 0, stats, bins);
    if (rc) return rc;
    return modl_bwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, g_ll, nullptr, dparams, st, 0,
                             stats, bins);
  }
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.partial = reinterpret_cast<double*>(ws);
  a.g_image = g_ll;
  a.dparams = dparams;
  a.HW = HW;
  a.n_px = n_px;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.plain = AR;
  a.spread = spread_runs();
  a.pair_rot = pair_rot_on(M);
  a.pix_stats = reinterpret_cast<float2*>(stats);
  if ((rc = set_bins(a, bins, AR))) return rc;
  StepFinish f{};
  f.extra = extra;
  f.ll = ll_image;
  f.ll64 = ll_image_f64;
  f.log_w = log_w;
  f.lme_b = lme_b;
  f.elbo = elbo;
  f.g_ll = g_ll;
  f.lme64 = reinterpret_cast<double*>(ws + partial_elems(n_img) * sizeof(double));
  f.B = B;
  f.S = S;
  f.b_norm = static_cast<float>(B_total > 0 ? B_total : B);
  f.small = n_px < (1ll << 31);
  f.peer = elbo ? take_peer() : PeerOut{};
  if (launches) *launches = 1;
  switch (M) {
    case 5:
      rc = launch_step<5, 1, AR>(a, f, n_img, st);
      break;
    case 10:
      rc = launch_step<10, 1, AR>(a, f, n_img, st);
      break;
    case 20:
      rc = launch_step<10, 2, AR>(a, f, n_img, st);
      break;
    default:
      rc = launch_step<10, 3, AR>(a, f, n_img, st);
      break;
  }
  if (rc == static_cast<int>(cudaErrorCooperativeLaunchTooLarge) || rc == static_cast<int>(cudaErrorLaunchOutOfResources)) {
    // the grid cannot be co-resident right now (another context holds SMs, MPS partition, ...): nothing was enqueued,
    // the step runs as three ordinary launches instead
    cudaGetLastError();
    give_peer(f.peer);
    if (launches) *launches = 3;
    rc = modl_iwae_fwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                                ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes, st, 0, stats, bins);
    if (rc) return rc;
    return modl_bwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, g_ll, nullptr, dparams, st, 0,
                             stats, bins);
  }
  return rc;
}
}  // namespace vaemdl
