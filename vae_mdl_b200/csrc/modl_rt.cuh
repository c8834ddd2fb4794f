#pragma once
// modl_rt.cuh -- the run-time tiled kernel (any n_mix) and the one-thread-per-pixel A/B kernel.
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
#include "modl_core.cuh"
#include "modl_tile.cuh"

namespace vaemdl {

// ---- the run-time tiled kernel (any n_mix) --------------------------------------------------------------------------------
// Same pipeline as modl_tile_kernel with one slot per warp, but n_mix and the split of a pixel over lanes are run-time
// values chosen by the host (rt_plan): LPP lanes share a pixel, lane `sub` of the group owns components
// [sub*MC, min(M, (sub+1)*MC)), two per packed register; PPT = 32 / LPP pixels per tile (one fewer when PPT * M would be
// odd: bulk copies need 16-byte multiples).  A lane walks its component pairs in an order rotated by rot * pixel so that
// the 32 scalar shared-memory loads of one instruction spread over the banks whatever 10 * M is modulo 32.
__device__ __forceinline__ float group_sum_rt(float v, int base, int LPP) {
  float s = __shfl_sync(kFull, v, base);
  for (int j = 1; j < LPP; ++j) s += __shfl_sync(kFull, v, (base + j) & 31);
  return s;
}
__device__ __forceinline__ float group_max_rt(float v, int base, int LPP) {
  float s = __shfl_sync(kFull, v, base);
  for (int j = 1; j < LPP; ++j) s = fmaxf(s, __shfl_sync(kFull, v, (base + j) & 31));
  return s;
}

// AL: n_mix and MC even -> every component pair sits on an 8-byte boundary and is moved with 64-bit shared accesses
template <bool BWD, int AR, bool AL, int PD = 0>
__global__ void __launch_bounds__(512, 1) modl_rt_kernel(const ModlArgs a) {
  const int M = a.M, MC = a.rt_MC, LPP = a.rt_LPP, PPT = a.rt_PPT;
  const int ROWF = 10 * M, TILE_F = PPT * ROWF, NPAIR = (MC + 1) >> 1;
  const int WARP_F = a.rt_warp_f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slot + TILE_F;
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
  const int p_raw = lane / LPP;
  const bool lane_used = p_raw < PPT;
  const int p = lane_used ? p_raw : 0;  // idle lanes shadow pixel 0 (they never write)
  const int sub = lane - p_raw * LPP;
  const int gbase = lane - sub;         // first lane of this pixel's group
  const int m0 = sub * MC;
  const int m_end = m0 + MC < M ? m0 + MC : M;
  const int m_safe = m0 < M ? m0 : 0;   // a group's last lane may own nothing (M = 9 over 4 lanes)
  int rot = (a.rt_rot * p) % NPAIR;

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
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
    const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * (PD ? 2 : 4);
    char* land = reinterpret_cast<char*>(slot) + (PD ? bytes : 0u);  // bf16 lands behind the room its float32 image needs
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(land, src, bytes, bar, pol_first);
          else
            bulk_g2s(land, src, bytes, bar);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(land, src, bytes, bar, (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(land, src, bytes, bar);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32)
        slot[i] = PD ? bf16_bits_to_f32(reinterpret_cast<const unsigned short*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };
  if (t_cnt > 0) issue(t_first);

  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + p) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + p) - n_own * a.HW);
  double acc0 = 0.0, acc1 = 0.0;
  long long n_base = (t_begin * PPT) / a.HW;

  auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw,
                   float& g_out) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    const bool in = p < rows;
    const long long n = in ? n_lane : n_first;
    const int pix = in ? pix_lane : pix_first;
    raw = load_pixel_raw(a, n, pix);
    g_out = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) g_out = a.g_image[n];
      if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
    }
    n_out = n;
    nfirst_out = n_first;
  };
  long long n_cur = 0, nfirst_cur = 0;
  PixRaw raw_cur{};
  float g_cur = 0.0f;
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    const int rows = tile_rows(t);
    const int pp = p < rows ? p : 0;
    const bool active = lane_used && (p < rows);
    const long long i = t * PPT + pp;
    const long long n = n_cur, n_first = nfirst_cur;
    const float g = g_cur;
    Pixel px;
    decode_pixel<AR>(a, raw_cur, px);
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
    if (it + 1 < t_cnt) fetch(t + t_dir, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

    float* rowp = slot + pp * ROWF;
    float* auxp = aux + pp * M;
    mbar_wait(bar, parity);
    if constexpr (PD != 0) {
      if (((rows * ROWF * 2) & 15) == 0) widen_bf16_inplace(slot, rows * ROWF, lane);
    }

    float lmax = -INFINITY;
    for (int m = m0; m < m_end; ++m) lmax = fmaxf(lmax, rowp[m]);
    lmax = group_max_rt(lmax, gbase, LPP);

    f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int prr = pr + rot >= NPAIR ? pr + rot - NPAIR : pr + rot;
      const int m = m0 + 2 * prr;
      const bool vlo = m < m_end, vhi = m + 1 < m_end;
      const int ml = vlo ? m : m_safe, mh = vhi ? m + 1 : ml;
      f2 lg = ld_pair<AL>(rowp, ml, !vhi);
      lg = pk(vlo ? lo(lg) : -INFINITY, vhi ? hi(lg) : -INFINITY);  // padding halves get zero weight
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* q = rowp + (1 + 3 * c) * M;
        mu[c] = ld_pair<AL>(q, ml, !vhi);
        sc[c] = ld_pair<AL>(q, M + ml, !vhi);
        kp[c] = ld_pair<AL>(q, 2 * M + ml, !vhi);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < (AR ? a.ls_narrow : kLsNarrow));
      const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, BWD, Pixel, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, BWD, Pixel, AR>(px, mu, sc, kp, u);
      if (!vlo) P = sp(0.0f);  // a padding pair was evaluated on clamped (backward: possibly overwritten) values
      sumW2 = sumW2 + W;
      sumWP2 = fma2(W, P, sumWP2);
      if constexpr (BWD) {
        const f2 wp = W * P;
        if (active && vlo) {  // (AL: a pair is valid or padding as a whole)
#pragma unroll
          for (int j = 0; j < 9; ++j) st_pair<AL>(rowp, (1 + j) * M + ml, !vhi, u[j]);
          st_pair<AL>(auxp, ml, !vhi, wp);
        }
      }
    }
    const float S = group_sum_rt(lo(sumWP2) + hi(sumWP2), gbase, LPP);
    const float SW = group_sum_rt(lo(sumW2) + hi(sumW2), gbase, LPP);
    const bool tiny = !(S > kTinySum);  // also catches NaN
    const float* grow = param_row(a, i, ROWF);

    if constexpr (!BWD) {
      if constexpr (PD != 0) fence_async_smem();
      __syncwarp();
      if (it + 1 < t_cnt) issue(t + t_dir);  // every lane has read its row: re-arm the slot with the warp's next tile
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;
      if (tiny) {
        float lt, ll;
        modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
        lp = lt - ll;
      }
      const bool owner = active && sub == 0;
      if (a.lp_pixel && owner) a.lp_pixel[i] = lp;
      const float val = owner ? lp : 0.0f;
      if (a.partial) {
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
      } else if (a.ll_atomic) {
        if (owner) atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    } else {
      const float rS = rcpa(S), rSW = rcpa(SW);
      float lt = 0.f, ll = 0.f;
      if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = pr + rot >= NPAIR ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        const bool vlo = m < m_end, vhi = m + 1 < m_end;
        const int ml = vlo ? m : m_safe, mh = vhi ? m + 1 : ml;
        const f2 lg = ld_pair<AL>(rowp, ml, !vhi);
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        const f2 wp = ld_pair<AL>(auxp, ml, !vhi);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, ml, px, a.plain != 0, PD != 0) - lt), expf(modl_logt(grow, M, mh, px, a.plain != 0, PD != 0) - lt));
          pi = pk(expf(ld_param(grow, ml, PD != 0) - ll), expf(ld_param(grow, mh, PD != 0) - ll));
        }
        const f2 gr = r * g;
        const f2 dl = (r - pi) * g;
        if (active && vlo) {
          st_pair<AL>(rowp, ml, !vhi, dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<AL>(rowp, j * M + ml, !vhi, ld_pair<AL>(rowp, j * M + ml, !vhi) * gr);
        }
      }
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
      char* dst = reinterpret_cast<char*>(a.dparams) + t * TILE_F * (PD ? 2 : 4);
      if ((bytes & 15u) == 0) {
        if constexpr (PD != 0) {
          __syncwarp();
          narrow_bf16_inplace(slot, rows * ROWF, lane);
        }
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
        for (int q = lane; q < rows * ROWF; q += 32) {
          if (PD)
            reinterpret_cast<unsigned short*>(dst)[q] = f32_to_bf16_bits(slot[q]);
          else
            reinterpret_cast<float*>(dst)[q] = slot[q];
        }
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

// ---- any-M kernel: one thread per pixel-sample, parameters straight from global memory (correct, not tuned) --------------
template <bool BWD>
__global__ void __launch_bounds__(128) modl_generic_kernel(const ModlArgs a) {
  const int M = a.M;
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_iter = (a.n_px + stride - 1) / stride;  // every lane runs the same trip count (warp votes inside)
  for (long long itn = 0; itn < n_iter; ++itn) {
    const long long i_raw = itn * stride + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = i_raw < a.n_px;
    const long long i = active ? i_raw : 0;
    Pixel px;
    const long long n = i / a.HW;
    load_pixel(a, n, static_cast<int>(i - n * a.HW), px);
    const float* row = a.params + i * 10 * M;
    float lmax = row[0];
    for (int m = 1; m < M; ++m) lmax = fmaxf(lmax, row[m]);
    float sumW = 0.f, sumWP = 0.f;
    for (int m = 0; m < M; ++m) {
      const float mu[3] = {row[M + m], row[4 * M + m], row[7 * M + m]};
      const float sc[3] = {row[2 * M + m], row[5 * M + m], row[8 * M + m]};
      const float kp[3] = {row[3 * M + m], row[6 * M + m], row[9 * M + m]};
      const float W = ex2a((row[m] - lmax) * kLog2e);
      sumW += W;
      sumWP = fmaf(W, mix_fwd(px, mu, sc, kp, a.plain != 0), sumWP);
    }
    const bool tiny = !(sumWP > kTinySum);
    float lt = 0.f, ll = 0.f;
    if (tiny) modl_pixel_logdomain(row, M, px, a.plain != 0, lt, ll);
    if constexpr (!BWD) {
      const float lp = tiny ? (lt - ll) : (lg2_split(sumWP) - lg2_split(sumW)) * kLn2;
      if (active) {
        if (a.lp_pixel) a.lp_pixel[i] = lp;
        if (a.ll_atomic) atomicAdd(a.ll_atomic + n, static_cast<double>(lp));
      }
    } else {
      float g = 0.f;
      if (a.g_image) g += a.g_image[n];
      if (a.g_pixel) g += a.g_pixel[i];
      const float rS = rcpa(sumWP), rSW = rcpa(sumW);
      float* orow = a.dparams + i * 10 * M;
      for (int m = 0; m < M; ++m) {
        const float mu[3] = {row[M + m], row[4 * M + m], row[7 * M + m]};
        const float sc[3] = {row[2 * M + m], row[5 * M + m], row[8 * M + m]};
        const float kp[3] = {row[3 * M + m], row[6 * M + m], row[9 * M + m]};
        const float W = ex2a((row[m] - lmax) * kLog2e);
        float u[9];
        const float P = mix_bwd(px, mu, sc, kp, u, a.plain != 0);
        float r = W * P * rS, pi = W * rSW;
        if (tiny) {
          r = expf(modl_logt(row, M, m, px, a.plain != 0) - lt);
          pi = expf(row[m] - ll);
        }
        if (active) {
          orow[m] = g * (r - pi);
          const float gr = g * r;
#pragma unroll
          for (int j = 0; j < 9; ++j) orow[(1 + j) * M + m] = gr * u[j];
        }
      }
    }
  }
}

}  // namespace vaemdl
