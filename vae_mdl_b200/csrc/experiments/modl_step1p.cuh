#pragma once
// NOT PART OF THE BUILD.  Kept as the record of a measured negative result (round 2, profiles/r02i_*): on B200 this kernel
// runs BASELINE configs[0] in 100.7 us against 97.2 us for modl_step_kernel (n_mix 5 x batch 128: 110.7 vs 99.2 us) although it
// executes 20 % fewer instructions (34.2 M vs 42.7 M): only ~50 MB of the gradient survive in L2 between the two passes
// (DRAM writes 165 MB instead of 84 MB), and pass 2 is a chain of dependent bulk copies per warp with nothing to compute in
// between (33 % of the issue slots busy, 24 % of the stall samples on the copy barriers).
//
// modl_step1p.cuh -- one IWAE step of the observation model with ONE evaluation of the logistic terms per pixel-sample.
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
//
// For the training shapes of models/model05.py (BASELINE configs[0]: 131 MB of parameters, ~4 tiles per warp) the
// forward + backward pair is bound by instruction issue, not by DRAM (ncu, profiles/r02g: 38 % of the DRAM peak, 46 % issue
// slots busy, 42.7 M warp instructions of which 15 M are the forward pass's evaluation of terms the backward pass
// evaluates again).  The upstream gradient of an image, g = -softmax_s(log_w) / B (models/loss.py:34-37), is one scalar per
// image, so the gradient row of a pixel-sample factors into g * (r_m * dlogP_m/dtheta): everything but g is known in the
// forward pass.  This kernel therefore
//   pass 1  walks each warp's run of tiles ONCE with the gradient arithmetic: per-pixel log-prob -> per-image float64
//           partial sums (the forward result), unscaled gradient rows -> dparams (bulk store; the run's last tile stays in
//           the warp's shared-memory slot);
//   grid barrier, IWAE finish spread over the grid (same arithmetic and order as finish_kernel), grid barrier;
//   pass 2  walks the run backwards and multiplies every row by its image's g: the resident tile first, the others come
//           back from L2 / DRAM by bulk copy, are scaled in shared memory (one multiply per element) and stored again.
// 2/3 of the instructions of forward + finish + backward; the price is that all but the resident tiles of the gradient
// are written twice, which is why only shapes whose gradient largely stays in the 126 MB L2 take this route.
#include "modl_tile.cuh"

namespace vaemdl {

template <int MC, int LPP, int AR>
__global__ void __launch_bounds__(512, 1) modl_step1p_kernel(const StepArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  using T = Tile<MC, LPP>;
  constexpr int M = T::M, PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F, NPAIR = T::NPAIR;
  constexpr bool AL = T::ALIGNED;
  constexpr int WARP_F = TILE_F + T::AUX_F;
  static_assert(T::AUX_F >= PPT, "the aux strip doubles as the per-row weight table of pass 2");
  const ModlArgs& a = sa.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slot + TILE_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();

  const long long gw = run_index(a, warp, nwarps);
  const bool lane_used = (lane / LPP) < PPT;
  const int p = lane_used ? (lane / LPP) : 0;
  const int sub = lane % LPP;
  const int m0 = sub * MC;
  const int rot = a.pair_rot ? T::pair_rot(lane) : 0;
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const long long t_cnt = t_end - t_begin;
  const uint64_t pol_first = policy_evict_first();
  uint32_t phase = 0;  // completed loads on this warp's mbarrier

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  // tile t of `base` (the parameters in pass 1, the unscaled gradient in pass 2) -> the slot
  auto issue = [&](const float* base, long long t) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
    const float* src = base + t * TILE_F;
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s_hint(slot, src, bytes, bar, pol_first);  // read once: do not displace the gradient rows pass 2 returns to
      }
    } else {  // ragged last tile
      for (int i = lane; i < rows * ROWF; i += 32) slot[i] = src[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };
  // the slot -> gradient tile t (bulk store; a ragged tile goes out with plain stores)
  auto store_tile = [&](long long t) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
    float* dst = a.dparams + t * TILE_F;
    if ((bytes & 15u) == 0) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(dst, slot, bytes);
        bulk_commit();
      }
    } else {
      __syncwarp();
      for (int q = lane; q < rows * ROWF; q += 32) dst[q] = slot[q];
      __syncwarp();
    }
  };

  // ---------------------------------------------------------------------------------------------------------------------
  // pass 1: forward order, gradient arithmetic, upstream weight 1
  // ---------------------------------------------------------------------------------------------------------------------
  if (t_cnt > 0) issue(a.params, t_begin);
  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_begin * PPT + p) / a.HW;
  int pix_own = static_cast<int>((t_begin * PPT + p) - n_own * a.HW);
  double acc0 = 0.0, acc1 = 0.0;
  long long n_base = (t_begin * PPT) / a.HW;

  auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    const bool in = p < rows;  // lanes past a ragged last tile shadow the tile's first pixel
    n_out = in ? n_lane : n_first;
    nfirst_out = n_first;
    raw = load_pixel_raw(a, n_out, in ? pix_lane : pix_first);
  };
  long long n_cur = 0, nfirst_cur = 0;
  PixRaw raw_cur{};
  if (t_cnt > 0) fetch(t_begin, n_own, pix_own, n_cur, nfirst_cur, raw_cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_begin + it;
    const int rows = tile_rows(t);
    const int pp = p < rows ? p : 0;
    const bool active = lane_used && (p < rows);
    const long long i = t * PPT + pp;
    const long long n = n_cur, n_first = nfirst_cur;
    Pixel px;
    decode_pixel<AR>(a, raw_cur, px);
    n_own += step_n;
    pix_own += step_pix;
    if (pix_own >= a.HW) {
      pix_own -= a.HW;
      ++n_own;
    }
    if (it + 1 < t_cnt) fetch(t + 1, n_own, pix_own, n_cur, nfirst_cur, raw_cur);

    float* rowp = slot + pp * ROWF;
    float* auxp = aux + pp * M;
    mbar_wait(bar, phase & 1u);
    ++phase;

    float lmax;
    if constexpr (AL) {
      lmax = -INFINITY;
#pragma unroll
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const float2 v = *reinterpret_cast<const float2*>(rowp + m0 + 2 * prr);
        lmax = fmaxf(lmax, fmaxf(v.x, v.y));
      }
    } else {
      lmax = rowp[m0];
#pragma unroll
      for (int m = 1; m < MC; ++m) lmax = fmaxf(lmax, rowp[m0 + m]);
    }
    lmax = group_max<LPP>(lmax, lane);

    f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
      const int m = m0 + 2 * prr;
      const bool single = (MC % 2 == 1) && (prr == NPAIR - 1);
      f2 lg = ld_pair<AL>(rowp, m, single);
      if (single) lg = pk(lo(lg), -INFINITY);
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mu[c] = ld_pair<AL>(rowp, (1 + 3 * c) * M + m, single);
        sc[c] = ld_pair<AL>(rowp, (2 + 3 * c) * M + m, single);
        kp[c] = ld_pair<AL>(rowp, (3 + 3 * c) * M + m, single);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < (AR ? a.ls_narrow : kLsNarrow));
      const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, true, Pixel, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, true, Pixel, AR>(px, mu, sc, kp, u);
      sumW2 = sumW2 + W;
      sumWP2 = fma2(W, P, sumWP2);
      if (active) {  // unscaled derivatives overwrite the component's parameters in place; W * P waits in the aux strip
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          st_pair<AL>(rowp, (1 + 3 * c) * M + m, single, u[3 * c + 0]);
          st_pair<AL>(rowp, (2 + 3 * c) * M + m, single, u[3 * c + 1]);
          st_pair<AL>(rowp, (3 + 3 * c) * M + m, single, u[3 * c + 2]);
        }
        st_pair<AL>(auxp, m, single, W * P);
      }
    }
    const float S = group_sum<LPP>(lo(sumWP2) + hi(sumWP2), lane);
    const float SW = group_sum<LPP>(lo(sumW2) + hi(sumW2), lane);
    const bool tiny = !(S > kTinySum);  // also catches NaN
    const float* grow = param_row(a, i, ROWF);
    float lt = 0.f, ll = 0.f;
    if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, false);

    {  // the forward result: per-pixel log-prob -> float64 per-image partial sums (same bookkeeping as tile_body)
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;
      if (tiny) lp = lt - ll;
      const bool owner = active && sub == 0;
      const float val = owner ? lp : 0.0f;
      while (n_base < n_first) {
        const double done = warp_sum(acc0);
        if (lane == 0) a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = done;
        acc0 = acc1;
        acc1 = 0.0;
        ++n_base;
      }
      if (n == n_base)
        acc0 += static_cast<double>(val);
      else
        acc1 += static_cast<double>(val);
    }

    {  // responsibilities; gradient rows with upstream weight 1
      const float rS = rcpa(S), rSW = rcpa(SW);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        const bool single = (MC % 2 == 1) && (prr == NPAIR - 1);
        f2 lg = ld_pair<AL>(rowp, m, single);
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        const f2 wp = ld_pair<AL>(auxp, m, single);
        f2 r = wp * rS;
        f2 pi = W * rSW;
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, m, px, a.plain != 0, false) - lt),
                 single ? 0.0f : expf(modl_logt(grow, M, m + 1, px, a.plain != 0, false) - lt));
          pi = pk(expf(grow[m] - ll), single ? 0.0f : expf(grow[m + 1] - ll));
        }
        if (active) {
          st_pair<AL>(rowp, m, single, r - pi);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<AL>(rowp, j * M + m, single, ld_pair<AL>(rowp, j * M + m, single) * r);
        }
      }
    }
    if (it + 1 < t_cnt) {  // not the run's last tile: out to dparams, next parameter tile in
      store_tile(t);
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      issue(a.params, t + 1);
    }
  }
  if (t_cnt > 0) {
    const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;
    const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
    if (lane == 0) {
      a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
      if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
    }
  }
  if (lane == 0) bulk_wait_all<0>();  // this warp's gradient tiles are complete in global memory before pass 2 re-reads them

  // ---------------------------------------------------------------------------------------------------------------------
  // IWAE finish spread over the grid (step_finish: the same arithmetic, in the same order, as finish_kernel)
  // ---------------------------------------------------------------------------------------------------------------------
  const long long gw_lin = static_cast<long long>(blockIdx.x) * nwarps + warp;
  const long long total_warps = static_cast<long long>(gridDim.x) * nwarps;
  __threadfence();
  grid.sync();
  step_finish(sa.f, gw_lin, total_warps, lane);
  __threadfence();
  grid.sync();
  asm volatile("fence.proxy.async;" ::: "memory");
  if (sa.f.elbo && gw_lin == total_warps - 1) {  // batch mean, fixed order
    double tt = 0.0;
    for (long long b = lane; b < sa.f.B; b += 32) tt += sa.f.lme64[b];
    tt = warp_sum(tt);
    if (lane == 0) sa.f.elbo[0] = static_cast<float>(tt / static_cast<double>(sa.f.b_norm));  // models/loss.py:37
  }

  // ---------------------------------------------------------------------------------------------------------------------
  // pass 2: backward order, every row times the upstream gradient of its image
  // ---------------------------------------------------------------------------------------------------------------------
  constexpr int VEC = (ROWF % 4 == 0) ? 4 : 2;  // elements per shared-memory access of the scaling loop
  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_end - 1 - it;
    const int rows = tile_rows(t);
    if (it > 0) {
      mbar_wait(bar, phase & 1u);
      ++phase;
    }
    // per-row weights: lane r of the first PPT lanes looks up the image of row r
    if (lane < rows) {
      const long long i = t * PPT + lane;
      const long long n = a.small ? static_cast<long long>(static_cast<unsigned>(i) / static_cast<unsigned>(a.HW)) : i / a.HW;
      aux[lane] = a.g_image[n];
    }
    __syncwarp();
    const int nvec = rows * ROWF / VEC;
    if constexpr (VEC == 4) {
      float4* s4 = reinterpret_cast<float4*>(slot);
      for (int q = lane; q < nvec; q += 32) {
        const float g = aux[(q * 4) / ROWF];
        float4 v = s4[q];
        v.x *= g;
        v.y *= g;
        v.z *= g;
        v.w *= g;
        s4[q] = v;
      }
    } else {
      float2* s2 = reinterpret_cast<float2*>(slot);
      for (int q = lane; q < nvec; q += 32) {
        const float g = aux[(q * 2) / ROWF];
        float2 v = s2[q];
        v.x *= g;
        v.y *= g;
        s2[q] = v;
      }
    }
    store_tile(t);
    if (it + 1 < t_cnt) {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      issue(a.dparams, t - 1);
    }
  }
  if (lane == 0) bulk_wait_all<0>();
}

}  // namespace vaemdl
