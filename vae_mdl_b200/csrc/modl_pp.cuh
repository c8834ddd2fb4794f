#pragma once
// modl_pp.cuh -- the pixel-pair kernel for n_mix 1..9 (two pixels per packed register).
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
#include "modl_core.cuh"
#include "modl_tile.cuh"

namespace vaemdl {

// ---- the pixel-pair kernel (small n_mix, e.g. the reference's own default n_mix = 5) ---------------------------------------
// Same pipeline as modl_tile_kernel (per-warp TMA bulk loads, runs of consecutive tiles, in-place gradient staging), but
// the two halves of a packed register hold the SAME mixture component of TWO pixels: lane l owns rows l and l + 32 of a
// 64-row tile.  Nothing is wasted on an odd component count, the tile is as large as the n_mix = 10 one (12.8 KB at
// n_mix = 5), and the two rows of a lane sit 32 rows apart so that the scalar shared-memory loads spread over the banks.
template <int M>
struct TilePP {
  static constexpr int PPT = 64;
  static constexpr int ROWF = 10 * M;
  static constexpr int TILE_F = PPT * ROWF;
  static constexpr int TILE_B = TILE_F * 4;
  static constexpr int AUX_F = PPT * M;
};

__device__ __forceinline__ Pixel half_pixel(const PixelPair& pp, bool hi_half) {
  Pixel px;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    px.x[c] = hi_half ? hi(pp.x[c]) : lo(pp.x[c]);
    px.left[c] = hi_half ? pp.lh[c] : pp.ll[c];
    px.right[c] = hi_half ? pp.rh[c] : pp.rl[c];
  }
  px.dx = pp.dx;
  px.width = pp.width;
  return px;
}

template <int M, bool BWD, int MAXT, int AR>
__global__ void __launch_bounds__(MAXT, 1) modl_pp_kernel(const ModlArgs a) {
  using T = TilePP<M>;
  constexpr int PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F;
  constexpr int WARP_F = TILE_F + (BWD ? T::AUX_F : 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slot + TILE_F;
  // Even n_mix: the row stride 10*M words shares a large power of two with the 32 banks (16-way conflicts at M = 8), so
  // every group of ROTB lanes walks the components in its own rotated order (<= 2-way for every M; the order of the
  // per-pixel sum then depends on the lane, the result stays reproducible run to run).
  constexpr int ROTB = (M % 2 != 0 || M < 2) ? 0 : (M == 8 ? 4 : 8);
  const int rot0 = ROTB ? (lane / (ROTB ? ROTB : 1)) % M : 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  if constexpr (BWD)
    pdl_wait();
  else
    pdl_trigger();

  const long long gw = run_index(a, warp, nwarps);
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const long long t_cnt = t_end - t_begin;
  const bool rev = BWD && a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  auto issue = [&](long long t) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
    const float* src = a.params + t * TILE_F;
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(slot, src, bytes, bar, pol_first);
          else
            bulk_g2s(slot, src, bytes, bar);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(slot, src, bytes, bar, (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(slot, src, bytes, bar);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32) slot[i] = src[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };
  if (t_cnt > 0) issue(t_first);

  // (image, pixel-in-image) of this lane's FIRST pixel-sample (row `lane` of the tile), advanced incrementally
  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + lane) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + lane) - n_own * a.HW);
  double acc0 = 0.0, acc1 = 0.0;
  const long long n_warp_first = (t_begin * PPT) / a.HW;
  long long n_base = n_warp_first;

  struct Fetched {
    long long nA, nB, n_first;
    int pixA, pixB;
    PixRaw rawA, rawB;
    float gA, gB;
  };
  auto fetch = [&](long long t, long long n_lane, int pix_lane, Fetched& f) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    long long nB = n_lane;
    int pixB = pix_lane + 32;  // the lane's second row, 32 rows further on
    while (pixB >= a.HW) {
      pixB -= a.HW;
      ++nB;
    }
    const bool inA = lane < rows, inB = lane + 32 < rows;  // rows past a ragged last tile shadow the tile's first pixel
    f.nA = inA ? n_lane : n_first;
    f.pixA = inA ? pix_lane : pix_first;
    f.nB = inB ? nB : n_first;
    f.pixB = inB ? pixB : pix_first;
    f.n_first = n_first;
    f.rawA = load_pixel_raw(a, f.nA, f.pixA);
    f.rawB = load_pixel_raw(a, f.nB, f.pixB);
    f.gA = f.gB = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) {
        f.gA = a.g_image[f.nA];
        f.gB = a.g_image[f.nB];
      }
      if (a.g_pixel) {
        f.gA += a.g_pixel[f.nA * a.HW + f.pixA];
        f.gB += a.g_pixel[f.nB * a.HW + f.pixB];
      }
    }
  };
  Fetched cur{};
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    const int rows = tile_rows(t);
    const bool actA = lane < rows, actB = lane + 32 < rows;
    const int ppA = actA ? lane : 0, ppB = actB ? lane + 32 : 0;
    const long long iA = t * PPT + ppA, iB = t * PPT + ppB;
    const long long nA = cur.nA, nB = cur.nB, n_first = cur.n_first;
    const f2 g2 = pk(cur.gA, cur.gB);
    PixelPair px;
    {
      Pixel pa, pb;
      decode_pixel<AR>(a, cur.rawA, pa);
      decode_pixel<AR>(a, cur.rawB, pb);
      px.dx = pa.dx;
      px.width = pa.width;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        px.x[c] = pk(pa.x[c], pb.x[c]);
        px.ll[c] = pa.left[c];
        px.lh[c] = pb.left[c];
        px.rl[c] = pa.right[c];
        px.rh[c] = pb.right[c];
      }
    }
    if (!rev) {
      n_own += step_n;
      pix_own += step_pix;
      if (pix_own >= a.HW) {
        pix_own -= a.HW;
        ++n_own;
      }
    } else {
      n_own -= step_n;
      pix_own -= step_pix;
      if (pix_own < 0) {
        pix_own += a.HW;
        --n_own;
      }
    }
    if (it + 1 < t_cnt) fetch(t + t_dir, n_own, pix_own, cur);

    float* rowA = slot + ppA * ROWF;
    float* rowB = slot + ppB * ROWF;
    float* auxA = aux + ppA * M;
    float* auxB = aux + ppB * M;
    mbar_wait(bar, parity);

    f2 lmax = pk(rowA[rot0], rowB[rot0]);
#pragma unroll
    for (int mi = 1; mi < M; ++mi) {
      const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
      lmax = pk(fmaxf(lo(lmax), rowA[m]), fmaxf(hi(lmax), rowB[m]));
    }

    f2 sumW = sp(0.0f), sumWP = sp(0.0f);
#pragma unroll 1
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
      const f2 lg = pk(rowA[m], rowB[m]);
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mu[c] = pk(rowA[(1 + 3 * c) * M + m], rowB[(1 + 3 * c) * M + m]);
        sc[c] = pk(rowA[(2 + 3 * c) * M + m], rowB[(2 + 3 * c) * M + m]);
        kp[c] = pk(rowA[(3 + 3 * c) * M + m], rowB[(3 + 3 * c) * M + m]);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < (AR ? a.ls_narrow : kLsNarrow));
      const f2 W = ex2_2((lg - lmax) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, BWD, PixelPair, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, BWD, PixelPair, AR>(px, mu, sc, kp, u);
      sumW = sumW + W;
      sumWP = fma2(W, P, sumWP);
      if constexpr (BWD) {
        // unscaled gradients overwrite the component's parameters in place; W*P goes to the aux strip (owners only)
        const f2 wp = W * P;
        if (actA) {
#pragma unroll
          for (int j = 0; j < 9; ++j) rowA[(1 + j) * M + m] = lo(u[j]);
          auxA[m] = lo(wp);
        }
        if (actB) {
#pragma unroll
          for (int j = 0; j < 9; ++j) rowB[(1 + j) * M + m] = hi(u[j]);
          auxB[m] = hi(wp);
        }
      }
    }
    const bool tinyA = !(lo(sumWP) > kTinySum), tinyB = !(hi(sumWP) > kTinySum);  // also catches NaN
    const float* growA = a.params + iA * ROWF;
    const float* growB = a.params + iB * ROWF;

    if constexpr (!BWD) {
      __syncwarp();
      if (it + 1 < t_cnt) issue(t + t_dir);  // every lane has read its rows: re-arm the slot with the warp's next tile
      float lpA = (lg2_split(lo(sumWP)) - lg2_split(lo(sumW))) * kLn2;  // utils/mdl.py:78-89 in one step
      float lpB = (lg2_split(hi(sumWP)) - lg2_split(hi(sumW))) * kLn2;
      if (tinyA) {
        float lt, ll;
        modl_pixel_logdomain(growA, M, half_pixel(px, false), a.plain != 0, lt, ll);
        lpA = lt - ll;
      }
      if (tinyB) {
        float lt, ll;
        modl_pixel_logdomain(growB, M, half_pixel(px, true), a.plain != 0, lt, ll);
        lpB = lt - ll;
      }
      if (a.lp_pixel) {
        if (actA) a.lp_pixel[iA] = lpA;
        if (actB) a.lp_pixel[iB] = lpB;
      }
      if (a.pix_stats) {  // (mixture sum, logit normaliser) per pixel-sample for a one-pass gradient kernel
        if (actA) a.pix_stats[iA] = make_float2(lo(sumWP), lo(sumW));
        if (actB) a.pix_stats[iB] = make_float2(hi(sumWP), hi(sumW));
      }
      const float valA = actA ? lpA : 0.0f, valB = actB ? lpB : 0.0f;
      if (a.partial) {
        // a tile holds pixels of at most two images (HW >= 64 on this route): n_first and n_first + 1
        while (n_base < n_first) {
          const double done = warp_sum(acc0);
          if (lane == 0) a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (nA == n_base)
          acc0 += static_cast<double>(valA);
        else
          acc1 += static_cast<double>(valA);
        if (nB == n_base)
          acc0 += static_cast<double>(valB);
        else
          acc1 += static_cast<double>(valB);
      } else if (a.ll_atomic) {
        if (actA) atomicAdd(a.ll_atomic + nA, static_cast<double>(valA));
        if (actB) atomicAdd(a.ll_atomic + nB, static_cast<double>(valB));
      }
    } else {
      const f2 rS = rcp_2(sumWP), rSW = rcp_2(sumW);
      float ltA = 0.f, llA = 0.f, ltB = 0.f, llB = 0.f;
      Pixel pxa, pxb;
      if (tinyA || tinyB) {
        pxa = half_pixel(px, false);
        pxb = half_pixel(px, true);
        if (tinyA) modl_pixel_logdomain(growA, M, pxa, a.plain != 0, ltA, llA);
        if (tinyB) modl_pixel_logdomain(growB, M, pxb, a.plain != 0, ltB, llB);
      }
#pragma unroll 1
      for (int mi = 0; mi < M; ++mi) {
        const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
        const f2 lg = pk(rowA[m], rowB[m]);
        const f2 W = ex2_2((lg - lmax) * kLog2e);
        const f2 wp = pk(auxA[m], auxB[m]);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tinyA || tinyB) {
          float rA = lo(r), rB = hi(r), piA = lo(pi), piB = hi(pi);
          if (tinyA) {
            rA = expf(modl_logt(growA, M, m, pxa, a.plain != 0) - ltA);
            piA = expf(growA[m] - llA);
          }
          if (tinyB) {
            rB = expf(modl_logt(growB, M, m, pxb, a.plain != 0) - ltB);
            piB = expf(growB[m] - llB);
          }
          r = pk(rA, rB);
          pi = pk(piA, piB);
        }
        const f2 gr = r * g2;
        const f2 dl = (r - pi) * g2;
        if (actA) {
          rowA[m] = lo(dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) rowA[j * M + m] *= lo(gr);
        }
        if (actB) {
          rowB[m] = hi(dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) rowB[j * M + m] *= hi(gr);
        }
      }
      // hand the gradient tile to the TMA engine
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
      float* dst = a.dparams + t * TILE_F;
      if ((bytes & 15u) == 0) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.bwd_hint & 2)
            bulk_s2g_hint(dst, slot, bytes, pol_first);
          else
            bulk_s2g(dst, slot, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int q = lane; q < rows * ROWF; q += 32) dst[q] = slot[q];
        __syncwarp();
      }
      if (it + 1 < t_cnt) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        issue(t + t_dir);
      }
    }
  }
  if constexpr (BWD) {
    if (lane == 0) bulk_wait_all<0>();
  } else {
    if (a.partial && t_cnt > 0) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

}  // namespace vaemdl
