// experiments/modl_tm_fwd.cuh -- NOT part of the build.  The forward kernel with its tile in tensor memory (round 2, second session).
// Built, bit-identical to the shared-memory forward kernel -- and slower at every shape measured (profiles/r03e_tm_fwd.txt):
// BASELINE configs[0] forward 34.2 -> 37.5 us, n_mix 20 / 30 at the same shape 51.3 -> 54.6 / 74.4 -> 82.3 us, headline 130 -> 138 us.
// Moving the tile from the slot into tensor memory (50 shared-memory loads + 15 tcgen05.st per lane and tile, then a wait) adds more
// serial latency to a warp's tile than the hidden load wait gives back: the forward kernel at these shapes is bound by the dependency
// latency of its arithmetic, 3.5-4 warps per scheduler, not by the wait for the next tile.  (The gradient kernel is different: its
// second pass touches every value anyway, so the swap into tensor memory rides on loads and stores it has to do.)
// To try it again: include this file after modl_tm.cuh and launch modl_tile_tm_fwd_kernel with the forward kernel's grid.
#pragma once
#include "../modl_tm.cuh"

namespace vaemdl {

// ---- forward kernel with the tile in tensor memory: for the small training shapes -------------------------------------------
// With one shared-memory slot per warp the forward kernel cannot take its next tile before every lane has read its row, i.e.
// before the tile is done: at a handful of tiles per warp (BASELINE configs[0]: 4.9) 14 % of the warp samples wait for loads.
// Here a warp first moves the tile from the slot into tensor memory (one 24-column block per component pair, as above),
// refills the slot at once and evaluates the tile from tensor memory while the next one lands.  Same arithmetic, same order,
// same split of the tile range over the warps as tile_body<.., false, 1, ..>: bit-identical outputs.  At the large shapes the
// shared-memory kernel is at 96-100 % of the HBM roofline already and the extra copy only costs (the host chooses).
template <int MC, int LPP, int AR>
__device__ __forceinline__ void tile_body_tm_fwd(const ModlArgs& a, unsigned char* smem_raw) {
  using T = Tile<MC, LPP>;
  constexpr int M = T::M, PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F, NPAIR = T::NPAIR;
  static_assert(tm_supported<MC, LPP>() && NPAIR * 24 <= 128, "forward on tensor memory: sixteen warps, 128 columns each");
  static_assert((ROWF * 4) % 16 == 0, "rows must be bulk-copyable one by one");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * TILE_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * TILE_F * 4) + warp;
  uint32_t* tmem_base_p = reinterpret_cast<uint32_t*>(smem_raw + static_cast<size_t>(nwarps) * (TILE_F * 4 + 8));

  if (warp == 0) tmem_alloc_512(tmem_base_p);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tm = *tmem_base_p + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(128 * (warp >> 2));
  if (a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  pdl_trigger();  // let the finish kernel's launch overlap this kernel's tail

  const long long gw = run_index(a, warp, nwarps);
  const bool lane_used = (lane / LPP) < PPT;
  const int p = lane_used ? (lane / LPP) : 0;
  const int sub = lane % LPP;
  const int m0 = sub * MC;
  const int rot = (T::ROT && a.pair_rot) ? T::pair_rot(lane) : 0;
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const int t_cnt = static_cast<int>(t_end - t_begin);
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  auto issue = [&](long long t) {
    if (lane == 0) {
      const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * 4;
      const uint32_t bytes = static_cast<uint32_t>(tile_rows(t)) * ROWF * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      if (a.keep_tiles > 0)
        bulk_g2s_hint(slot, src, bytes, bar, (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
      else
        bulk_g2s(slot, src, bytes, bar);
    }
  };

  if (t_cnt > 0) {
    issue(t_begin);
    const long long step_n = PPT / a.HW;
    const int step_pix = static_cast<int>(PPT - step_n * a.HW);
    long long n_own = (t_begin * PPT + p) / a.HW;
    int pix_own = static_cast<int>((t_begin * PPT + p) - n_own * a.HW);
    double acc0 = 0.0, acc1 = 0.0;
    long long n_base = (t_begin * PPT) / a.HW;
    auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw) {
      const long long n_first = __shfl_sync(kFull, n_lane, 0);
      const int pix_first = __shfl_sync(kFull, pix_lane, 0);
      const bool in = p < tile_rows(t);  // lanes past a ragged last tile shadow the tile's first pixel
      const long long n = in ? n_lane : n_first;
      raw = load_pixel_raw(a, n, in ? pix_lane : pix_first);
      n_out = n;
      nfirst_out = n_first;
    };
    long long n_cur = 0, nfirst_cur = 0;
    PixRaw raw_cur{};
    fetch(t_begin, n_own, pix_own, n_cur, nfirst_cur, raw_cur);

    for (int it = 0; it < t_cnt; ++it) {
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

      // slot -> tensor memory, row by row (thread i moves its own row), and the maximum logit on the way
      const float* rowp = slot + pp * ROWF;
      float lmax = -INFINITY;
      mbar_wait(bar, static_cast<uint32_t>(it) & 1u);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        Blk b;
        b.r[0] = b.r[1] = b.r[22] = b.r[23] = 0u;
#pragma unroll
        for (int g = 0; g < 10; ++g) blk_set(b, 1 + g, ld_pair<true>(rowp, g * M + m, false));
        const f2 lg = blk_get(b, 1);
        lmax = fmaxf(lmax, fmaxf(lo(lg), hi(lg)));
        tmem_st24(tm + 24u * pr, b);
      }
      tmem_wait_st();
      lmax = group_max<LPP>(lmax, lane);
      __syncwarp();
      if (it + 1 < t_cnt) issue(t + 1);  // every lane has read its row: the slot takes the next tile

      f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
      Blk b;
      tmem_ld24_issue(tm, b);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        tmem_ld24_wait(b);
        const f2 lg = blk_get(b, 1);
        f2 mu[3], sc[3], kp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mu[c] = blk_get(b, 2 + 3 * c);
          sc[c] = blk_get(b, 3 + 3 * c);
          kp[c] = blk_get(b, 4 + 3 * c);
        }
        const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
        const bool narrow = __any_sync(kFull, smin < (AR ? a.ls_narrow : kLsNarrow));
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        f2 u[9];
        f2 P;
        if (narrow)
          P = pair_eval<true, false, Pixel, AR>(px, mu, sc, kp, u);
        else
          P = pair_eval<false, false, Pixel, AR>(px, mu, sc, kp, u);
        __syncwarp();
        if (pr + 1 < NPAIR) tmem_ld24_issue(tm + 24u * (pr + 1), b);
        sumW2 = sumW2 + W;
        sumWP2 = fma2(W, P, sumWP2);
      }
      const float S = group_sum<LPP>(lo(sumWP2) + hi(sumWP2), lane);
      const float SW = group_sum<LPP>(lo(sumW2) + hi(sumW2), lane);
      const bool tiny = !(S > kTinySum);  // also catches NaN
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;  // utils/mdl.py:78-89 in one step
      if (tiny) {
        float lt, ll;
        modl_pixel_logdomain(param_row(a, i, ROWF), M, px, a.plain != 0, lt, ll, false);
        lp = lt - ll;
      }
      __syncwarp();
      const bool owner = active && sub == 0;
      if (a.pix_stats && owner) a.pix_stats[i] = make_float2(S, SW);
      if (a.lp_pixel && owner) a.lp_pixel[i] = lp;
      const float val = owner ? lp : 0.0f;
      if (a.partial) {
        while (n_base < n_first) {  // the warp has left image n_base: its sum leaves the registers (warp-uniform)
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
      } else if (a.ll_atomic) {
        if (owner) atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    }
    if (a.partial) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
  tmem_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc_512(*tmem_base_p);
}

template <int MC, int LPP, int AR>
__global__ void __launch_bounds__(512, 1) modl_tile_tm_fwd_kernel(const ModlArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  tile_body_tm_fwd<MC, LPP, AR>(a, smem_raw);
}

}  // namespace vaemdl
