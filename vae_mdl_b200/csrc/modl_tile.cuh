#pragma once
// modl_tile.cuh -- the tiled kernel for n_mix = MC * LPP (5, 8, 10, 12, 16, 20, 24, 30, 32, 40, 64) and the one-launch cooperative
// step built on it.
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
#include "modl_core.cuh"

namespace vaemdl {

// ---- shared-memory bank model of the tiled kernel (compile time) ---------------------------------------------------------
// A row is 10 M words, so rows of different lanes start in a few banks only (M = 16 / 32 / 64: ALL rows start in bank 0) and
// the 64-bit accesses of a half-warp to "component pair pr of my row" collide.  A lane can walk its pairs in a rotated order,
// the rotation taken from a small family of lane functions.  rot_wavefronts() counts the wavefronts of one pass over the pairs
// (two half-warps per access, the degree of the worst bank each); best_rot_kind() picks the cheapest member -- but only when
// row order costs at least 3x the conflict-free count: measured on B200 (tools/ab_rot.sh, tools/ncu_smem.sh), shapes at 2.0x
// (n_mix 10 / 20 / 40) gain nothing from it (n_mix 10 even loses 7 % back to back although 65 % of the conflict wavefronts
// go away: the shared-memory pipe is not their limit), at 2.5x (n_mix 30) the rotated order is level backward and 3 points
// behind forward (the rotation arithmetic in the rolled loops), n_mix 16 / 32 (8x / 4x) go from 32-57 % to 84-87 % of the
// HBM roofline.  Where the rotation is off the pair index stays a compile-time sequence.
__host__ __device__ constexpr int rot_kind_value(int kind, int lane, int LPP, int NP) {
  const int p = lane / LPP;
  int r = 0;
  switch (kind) {
    case 1: r = (lane >> 3) & 1; break;
    case 2: r = ((lane >> 3) & 1) * 2; break;
    case 3: r = p; break;
    case 4: r = lane >> 1; break;
    case 5: r = lane >> 2; break;
    case 6: r = lane >> 3; break;
    case 7: r = 3 * p; break;
    default: r = 0; break;
  }
  return r % NP;
}
constexpr int rot_wavefronts(int M, int MC, int LPP, int kind) {
  const int PPT = 32 / LPP, NP = MC / 2, ROWF = 10 * M;
  int tot = 0;
  for (int pr = 0; pr < NP; ++pr) {
    for (int half = 0; half < 2; ++half) {
      int cnt[32] = {};
      int addr[32][16] = {};
      for (int lane = 16 * half; lane < 16 * half + 16; ++lane) {
        int p = lane / LPP;
        const int sub = lane % LPP;
        if (p >= PPT) p = 0;
        const int prr = (pr + rot_kind_value(kind, lane, LPP, NP)) % NP;
        const int a = p * ROWF + sub * MC + 2 * prr;
        for (int w = a; w < a + 2; ++w) {
          const int b = w % 32;
          bool seen = false;
          for (int q = 0; q < cnt[b]; ++q) seen = seen || addr[b][q] == w;
          if (!seen) addr[b][cnt[b]++] = w;
        }
      }
      int mx = 0;
      for (int b = 0; b < 32; ++b) mx = cnt[b] > mx ? cnt[b] : mx;
      tot += mx;
    }
  }
  return tot;
}
constexpr int best_rot_kind(int M, int MC, int LPP) {
  if (MC % 2 != 0 || MC < 4) return 0;
  const int ideal = 2 * (MC / 2);
  int best = 0, bw = rot_wavefronts(M, MC, LPP, 0);
  if (bw < 3 * ideal) return 0;  // row order is within 3x of conflict-free: leave it alone (see above)
  for (int k = 1; k <= 7; ++k) {
    const int w = rot_wavefronts(M, MC, LPP, k);
    if (w < bw) {
      bw = w;
      best = k;
    }
  }
  return best;
}

// ---- the tiled kernel -------------------------------------------------------------------------------------------------
// M = MC * LPP mixtures; LPP lanes share a pixel, each owning MC consecutive components, processed two at a time.
template <int MC, int LPP>
struct Tile {
  static constexpr int M = MC * LPP;
  static constexpr int PPT = 32 / LPP;  // pixels per warp tile
  static constexpr int ROWF = 10 * M;
  static constexpr int TILE_F = PPT * ROWF;
  static constexpr int TILE_B = TILE_F * 4;
  static constexpr int AUX_F = PPT * M;  // backward: W*P per (pixel, component)
  static constexpr int NPAIR = (MC + 1) / 2;
  static constexpr bool ALIGNED = (M % 2 == 0) && (MC % 2 == 0);  // component pairs sit on 8-byte boundaries
  static_assert(TILE_B % 16 == 0, "bulk copies need 16-byte multiples");
  static constexpr int ROT_KIND = ALIGNED ? best_rot_kind(M, MC, LPP) : 0;  // (bank model above)
  static constexpr bool ROT = ROT_KIND != 0;
  __device__ static __forceinline__ int pair_rot(int lane) { return rot_kind_value(ROT_KIND, lane, LPP, NPAIR); }
};

// a pair of consecutive floats at row[off], row[off+1]; `single`: only row[off] exists (odd MC), both halves get it
template <bool ALIGNED>
__device__ __forceinline__ f2 ld_pair(const float* row, int off, bool single) {
  if constexpr (ALIGNED) {
    const float2 t = *reinterpret_cast<const float2*>(row + off);
    return pk(t.x, t.y);
  } else {
    const float a = row[off];
    const float b = single ? a : row[off + 1];
    return pk(a, b);
  }
}
template <bool ALIGNED>
__device__ __forceinline__ void st_pair(float* row, int off, bool single, f2 v) {
  if constexpr (ALIGNED) {
    *reinterpret_cast<float2*>(row + off) = make_float2(lo(v), hi(v));
  } else {
    row[off] = lo(v);
    if (!single) row[off + 1] = hi(v);
  }
}

// A component pair of a row that holds bfloat16 values (PD = 2): the two neighbours share one aligned 32-bit word (the
// pair index is even); widening is a shift and a mask, narrowing one cvt.rn.bf16x2.f32.
__device__ __forceinline__ f2 ld_pair_bf16(const float* row_as_bf16, int off) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned short*>(row_as_bf16) + off);
  return pk(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ void st_pair_bf16(float* row_as_bf16, int off, f2 v) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo(v), hi(v));  // .x = low half-word = the lower index
  *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(row_as_bf16) + off) = *reinterpret_cast<const uint32_t*>(&b);
}
template <bool ALIGNED, bool DIRECT>
__device__ __forceinline__ f2 ld_row(const float* row, int off, bool single) {
  if constexpr (DIRECT)
    return ld_pair_bf16(row, off);
  else
    return ld_pair<ALIGNED>(row, off, single);
}
template <bool ALIGNED, bool DIRECT>
__device__ __forceinline__ void st_row(float* row, int off, bool single, f2 v) {
  if constexpr (DIRECT)
    st_pair_bf16(row, off, v);
  else
    st_pair<ALIGNED>(row, off, single, v);
}

struct PixRaw {
  unsigned v[3];
};
__device__ __forceinline__ PixRaw load_pixel_raw(const ModlArgs& a, long long n, int pix) {
  // image n is scored against x[n % x_batch].  The 64-bit remainder is a ~100-instruction subroutine executed per tile and
  // lane (every importance sample but the first has n >= x_batch): 32-bit arithmetic whenever the problem allows it.
  long long xb;
  if (a.x_batch == 1)
    xb = 0;
  else if (n < a.x_batch)
    xb = n;
  else if (a.small)
    xb = static_cast<long long>(static_cast<unsigned>(n) % static_cast<unsigned>(a.x_batch));
  else
    xb = n % a.x_batch;
  const long long xo = (xb * a.HW + pix) * 3;
  PixRaw r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    r.v[c] = a.x_u8 ? static_cast<unsigned>(static_cast<const uint8_t*>(a.x)[xo + c])
                    : __float_as_uint(static_cast<const float*>(a.x)[xo + c]);
  return r;
}
template <int AR>
__device__ __forceinline__ void decode_pixel(const ModlArgs& a, const PixRaw& r, Pixel& px) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = a.x_u8 ? u8_to_unit(r.v[c]) : __uint_as_float(r.v[c]);  // utils/data.py:15-16
    if (a.x_unit) v = __fmaf_rn(v, 2.0f, -1.0f);                                                  // utils/mdl.py:65
    px.x[c] = v;
    if constexpr (AR != 0) {  // utils/discretized_logistic.py:71-76 with the class's own low / high
      px.left[c] = v <= a.low;
      px.right[c] = v >= a.high;
    } else {
      px.left[c] = a.edge_openai ? (v < -0.999f) : (v <= -1.0f);
      px.right[c] = a.edge_openai ? (v > 0.999f) : (v >= 1.0f);
    }
  }
  px.dx = bin_dx<AR>(a);
  px.width = bin_width<AR>(a);
}

// FUSED: the forward and the backward pass of one step run inside ONE cooperative kernel (modl_step_kernel): both use
// the backward shared-memory layout, the mbarrier is initialised once and its phase carries over, the forward pass
// leaves its last tile in the slot and the (reversed) backward pass starts on it without loading anything.
// PD = 1: bfloat16 parameters / gradient in global memory (widened / narrowed in place in the slot, see widen_bf16_inplace)
// PD = 2: bfloat16 parameters / gradient that STAY bfloat16 in the slot (half the shared memory: two slots per warp, the next
//         tile lands while this one is processed); every component pair is widened as it is read and the final gradient
//         narrowed as it is written, one rounding.  Needs 4-byte aligned pairs (T::ALIGNED) and, backward, the forward
//         pass's per-pixel sums (ST): unscaled derivatives must never be rounded to bfloat16.
// ST: the backward pass takes every pixel's mixture sums from a.pix_stats (written by the forward pass of the same step)
// instead of forming them itself: the gradient of a component is scaled as soon as it is computed -- one pass over the
// row, no aux strip.  (The forward body of a FUSED step is instantiated with the same ST: both share the slot layout.)
template <int MC, int LPP, bool BWD, int NSLOT, int AR, bool FUSED, int PD = 0, bool ST = false>
__device__ __forceinline__ void tile_body(const ModlArgs& a, unsigned char* smem_raw) {
  using T = Tile<MC, LPP>;
  constexpr int M = T::M, PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F, NPAIR = T::NPAIR;
  constexpr bool AL = T::ALIGNED;
  // ST: instead of the aux strip the slot is followed by the tile's (S, SW) pairs, which travel with the tile (bulk copy)
  constexpr int STAT_F = 2 * PPT;
  constexpr bool DIRECT = PD == 2;
  constexpr int SLOT_F = DIRECT ? TILE_F / 2 : TILE_F;            // floats of shared memory per slot
  constexpr int NSTAT = (ST && NSLOT > 1) ? NSLOT : 1;            // one (S, SW) strip per slot in flight
  constexpr int WARP_F = NSLOT * SLOT_F + (((BWD || FUSED) && !ST) ? T::AUX_F : 0) + (ST ? NSTAT * STAT_F : 0);
  static_assert(!FUSED || NSLOT == 1, "the fused step keeps one slot per warp");
  static_assert(PD != 1 || (NSLOT == 1 && !FUSED), "widened bf16 parameters: one slot per warp, three-launch step");
  static_assert(!DIRECT || (AL && !FUSED && (!BWD || ST)), "direct bf16: aligned pairs, no unscaled derivatives in bf16");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slots = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slots + NSLOT * SLOT_F;   // (ST: the tiles' (S, SW) pairs live here, one strip per slot)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp * NSLOT;

  if constexpr (!(FUSED && BWD)) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) mbar_init(&bars[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
  }
  // A backward kernel launched programmatically behind the finish kernel may not read g_image (nor the forward pass's
  // per-pixel sums) before griddepcontrol.wait.  The two-pass kernel needs the upstream gradient only in the SECOND pass
  // over a row, so it loads its first tile and forms that tile's unscaled derivatives while the finish kernel is still
  // running (LATE_G): the finish kernel (6-8 us per step) leaves the critical path.  The parameters themselves are inputs
  // of the whole step, not products of the preceding kernels.
  constexpr bool LATE_G = BWD && !FUSED && !ST;
  if constexpr (!FUSED) {
    if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
    if constexpr (BWD) {
      if constexpr (!LATE_G) pdl_wait();
    } else {
      pdl_trigger();  // let the finish kernel's launch overlap this kernel's tail
    }
  }

  const long long gw = run_index(a, warp, nwarps);
  const bool lane_used = (lane / LPP) < PPT;
  const int p = lane_used ? (lane / LPP) : 0;  // idle lanes (LPP=3: lanes 30,31) shadow pixel 0
  const int sub = lane % LPP;
  const int m0 = sub * MC;
  const int rot = (T::ROT && a.pair_rot) ? T::pair_rot(lane) : 0;  // this lane's first component pair (bank-conflict rotation)

  // this warp's run of CONSECUTIVE tiles (balanced split of the tile range over all warps of the grid): per-image
  // sums then accumulate in registers across tiles and leave the warp once per image instead of once per tile
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  // bring tile t into slot s (bulk copy when the byte count allows it, plain loads for a ragged tail tile)
  const long long t_cnt = t_end - t_begin;
  const bool rev = BWD && a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
  auto issue = [&](long long t, int s) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
    const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * (PD ? 2 : 4);
    float* slot_f = slots + s * SLOT_F;
    char* dst = reinterpret_cast<char*>(slot_f) + (PD == 1 ? bytes : 0u);  // (PD = 1: bf16 lands behind the room its float32 image needs)
    [[maybe_unused]] float* stat_s = aux + (NSTAT > 1 ? s * STAT_F : 0);
    [[maybe_unused]] const uint32_t sbytes = static_cast<uint32_t>(rows) * 8u;  // the tile's (S, SW) pairs (BWD && ST)
    if ((bytes & 15u) == 0 && (!(BWD && ST) || (sbytes & 15u) == 0)) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[s], bytes + ((BWD && ST) ? sbytes : 0u));
        if constexpr (BWD && ST) bulk_g2s(stat_s, a.pix_stats + t * PPT, sbytes, &bars[s]);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(dst, src, bytes, &bars[s], pol_first);
          else
            bulk_g2s(dst, src, bytes, &bars[s]);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(dst, src, bytes, &bars[s], (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(dst, src, bytes, &bars[s]);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32) {
        if constexpr (DIRECT)
          reinterpret_cast<unsigned short*>(slot_f)[i] = reinterpret_cast<const unsigned short*>(src)[i];
        else
          slot_f[i] = PD ? bf16_bits_to_f32(reinterpret_cast<const unsigned short*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
      }
      if constexpr (BWD && ST) {
        if (lane < rows) reinterpret_cast<float2*>(stat_s)[lane] = a.pix_stats[t * PPT + lane];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&bars[s], 0);
    }
  };

  // forward: every slot is in flight from the start; backward: slots are refilled one tile ahead (see below);
  // fused backward: the first tile (the forward pass's last) is already in the slot
  if constexpr (!(FUSED && BWD)) {
#pragma unroll
    for (int s = 0; s < (BWD ? 1 : NSLOT); ++s) {
      if (s < t_cnt) issue(t_first + s * t_dir, s);
    }
  }
  // loads this warp's barrier has completed before this pass (fused backward: the whole forward pass but the resident tile)
  const uint32_t phase0 = (FUSED && BWD) ? static_cast<uint32_t>(t_cnt - 1) : 0u;

  // (image, pixel-in-image) of this lane's pixel-sample, advanced incrementally: one 64-bit division per kernel
  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + p) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + p) - n_own * a.HW);
  // float64 running sums of the image the warp is in (acc0, image n_base) and of the next one (acc1), per lane
  double acc0 = 0.0, acc1 = 0.0;
  const long long n_warp_first = (t_begin * PPT) / a.HW;
  long long n_base = n_warp_first;

  // software prefetch of the (L2-resident) pixel and upstream-gradient values one tile ahead
  auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw,
                   float& g_out, float2& st_out, bool want_g = true) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    const bool in = p < rows;  // lanes past a ragged last tile shadow the tile's first pixel
    const long long n = in ? n_lane : n_first;
    const int pix = in ? pix_lane : pix_first;
    raw = load_pixel_raw(a, n, pix);
    g_out = 0.0f;
    if constexpr (BWD) {
      if (want_g) {
        if (a.g_image) g_out = a.g_image[n];
        if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
      }
    }
    (void)st_out;
    n_out = n;
    nfirst_out = n_first;
  };
  // the upstream gradient of pixel `pix` of image n on its own (LATE_G: the first tile's, read after griddepcontrol.wait)
  [[maybe_unused]] auto fetch_g = [&](long long n, int pix) -> float {
    float g_out = 0.0f;
    if (a.g_image) g_out = a.g_image[n];
    if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
    return g_out;
  };

  long long n_cur = 0, nfirst_cur = 0;
  PixRaw raw_cur{};
  float g_cur = 0.0f;
  float2 st_cur = make_float2(1.0f, 1.0f);
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur, st_cur, !LATE_G);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const int s = static_cast<int>(it % NSLOT);
    const uint32_t parity = (phase0 + static_cast<uint32_t>(it / NSLOT)) & 1u;
    const int rows = tile_rows(t);
    const int pp = p < rows ? p : 0;
    const bool active = lane_used && (p < rows);
    const long long i = t * PPT + pp;  // this lane's pixel-sample
    const long long n = n_cur, n_first = nfirst_cur;
    float g = g_cur;
    const float2 st = st_cur;
    Pixel px;
    decode_pixel<AR>(a, raw_cur, px);
    [[maybe_unused]] int pix_this = pix_own;  // this lane's pixel within its image (before the advance below)
    if constexpr (LATE_G) {
      const int pix_first = __shfl_sync(kFull, pix_own, 0);  // (every lane takes part: lanes past a ragged tile shadow lane 0)
      if (!(p < rows)) pix_this = pix_first;
    }
    // advance the index and prefetch the next tile's pixel / upstream gradient
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
    // (LATE_G: the first iteration's look-ahead reads the upstream gradient too, so it waits until after griddepcontrol.wait)
    if (it + 1 < t_cnt && !(LATE_G && it == 0)) fetch(t + t_dir, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur, st_cur);

    float* slot = slots + s * SLOT_F;
    // (DIRECT: the row holds ROWF bfloat16 values; rowp is only ever handed to ld_row / st_row)
    float* rowp = DIRECT ? reinterpret_cast<float*>(reinterpret_cast<unsigned short*>(slot) + pp * ROWF) : slot + pp * ROWF;
    float* auxp = aux + pp * M;
    [[maybe_unused]] float* stat_cur = aux + (NSTAT > 1 ? s * STAT_F : 0);
    if (!(FUSED && BWD && it == 0)) mbar_wait(&bars[s], parity);
    if constexpr (FUSED && BWD && ST) {
      if (it == 0) {  // the resident tile was not loaded by this pass: fetch its (S, SW) pairs by hand
        if (lane < rows) reinterpret_cast<float2*>(aux)[lane] = a.pix_stats[t * PPT + lane];
        __syncwarp();
      }
    }
    if constexpr (PD == 1) {
      if (((rows * ROWF * 2) & 15) == 0) widen_bf16_inplace(slot, rows * ROWF, lane);  // (a ragged tile was widened by its loads)
    }
    // Two slots (BWD): the other slot's gradient tile was handed to the TMA engine at the end of the previous iteration; once
    // its shared-memory reads are done that slot is refilled with this warp's next tile.  Those reads take microseconds while
    // the memory system is saturated with writes, so the warp does not wait for them here but before component pair
    // a.tm_refill of this tile (0: here, as in round 1) -- the later, the less it waits, as long as the tile still lands in time.
    [[maybe_unused]] auto refill_other = [&]() {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      issue(t + t_dir, s ^ 1);
    };
    [[maybe_unused]] const bool refill = BWD && NSLOT > 1 && it + 1 < t_cnt;
    if constexpr (BWD && NSLOT > 1) {
      if (refill && a.tm_refill <= 0) refill_other();
    }

    // W_m = exp(logit_m - max logit)
    float lmax;
    if constexpr (AL) {  // 64-bit loads in the rotated pair order (no bank conflicts)
      lmax = -INFINITY;
#pragma unroll
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const f2 t = ld_row<AL, DIRECT>(rowp, m0 + 2 * prr, false);
        lmax = fmaxf(lmax, fmaxf(lo(t), hi(t)));
      }
    } else {
      lmax = rowp[m0];
#pragma unroll
      for (int m = 1; m < MC; ++m) lmax = fmaxf(lmax, rowp[m0 + m]);
    }
    lmax = group_max<LPP>(lmax, lane);

    // ST backward: the pixel's sums come from the forward pass
    [[maybe_unused]] float st_rS = 0.f, st_rSW = 0.f, st_lt = 0.f, st_ll = 0.f;
    [[maybe_unused]] bool st_tiny = false;
    if constexpr (BWD && ST) {
      const float2 stp = reinterpret_cast<const float2*>(stat_cur)[pp];  // arrived with the tile
      (void)st;
      st_rS = rcpa(stp.x);
      st_rSW = rcpa(stp.y);
      st_tiny = !(stp.x > kTinySum);  // the linear-domain sum left the float32 range (or is NaN): log-domain responsibilities
      if (st_tiny) modl_pixel_logdomain(param_row(a, i, ROWF), M, px, a.plain != 0, st_lt, st_ll, PD != 0);
    }
    f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      if constexpr (BWD && NSLOT > 1) {
        if (refill && pr > 0 && pr == a.tm_refill) refill_other();
      }
      const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
      const int m = m0 + 2 * prr;
      const bool single = (MC % 2 == 1) && (prr == NPAIR - 1);
      f2 lg = ld_row<AL, DIRECT>(rowp, m, single);
      if (single) lg = pk(lo(lg), -INFINITY);  // the padding half gets zero weight
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mu[c] = ld_row<AL, DIRECT>(rowp, (1 + 3 * c) * M + m, single);
        sc[c] = ld_row<AL, DIRECT>(rowp, (2 + 3 * c) * M + m, single);
        kp[c] = ld_row<AL, DIRECT>(rowp, (3 + 3 * c) * M + m, single);
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
      if constexpr (BWD && ST) {
        f2 r = (W * P) * st_rS;   // posterior responsibility of the component
        f2 pi = W * st_rSW;       // softmax(logits)
        if (st_tiny) {
          const float* grow_t = param_row(a, i, ROWF);
          r = pk(expf(modl_logt(grow_t, M, m, px, a.plain != 0, PD != 0) - st_lt),
                 single ? 0.0f : expf(modl_logt(grow_t, M, m + 1, px, a.plain != 0, PD != 0) - st_lt));
          pi = pk(expf(ld_param(grow_t, m, PD != 0) - st_ll), single ? 0.0f : expf(ld_param(grow_t, m + 1, PD != 0) - st_ll));
        }
        const f2 gr = r * g;
        if (active) {  // final gradients overwrite the component's parameters in place
          st_row<AL, DIRECT>(rowp, m, single, (r - pi) * g);
#pragma unroll
          for (int j = 0; j < 9; ++j) st_row<AL, DIRECT>(rowp, (1 + j) * M + m, single, u[j] * gr);
        }
      } else {
      sumW2 = sumW2 + W;
      sumWP2 = fma2(W, P, sumWP2);
      }
      if constexpr (BWD && !ST) {
        // unscaled gradients overwrite the component's parameters in place; W*P goes to the aux strip.
        // Only lanes that own a real pixel write: shadow lanes (ragged tile, or lanes 30/31 when 3 lanes share a
        // pixel) would otherwise race with the owner of pixel 0.
        if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          st_pair<AL>(rowp, (1 + 3 * c) * M + m, single, u[3 * c + 0]);
          st_pair<AL>(rowp, (2 + 3 * c) * M + m, single, u[3 * c + 1]);
          st_pair<AL>(rowp, (3 + 3 * c) * M + m, single, u[3 * c + 2]);
        }
        st_pair<AL>(auxp, m, single, W * P);
        }
      }
    }
    float S = 1.0f, SW = 1.0f;
    if constexpr (!(BWD && ST)) {
      S = group_sum<LPP>(lo(sumWP2) + hi(sumWP2), lane);
      SW = group_sum<LPP>(lo(sumW2) + hi(sumW2), lane);
    }
    [[maybe_unused]] const bool tiny = !(S > kTinySum);  // also catches NaN
    const float* grow = param_row(a, i, ROWF);

    if constexpr (!BWD) {
      if constexpr (PD == 1) fence_async_smem();  // the widening wrote the slot through the generic proxy
      __syncwarp();
      {  // every lane has read its row: re-arm the slot for this warp's tile NSLOT iterations ahead
        if (it + NSLOT < t_cnt) issue(t + NSLOT * t_dir, s);
      }
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;  // utils/mdl.py:78-89 in one step
      if (tiny) {
        float lt, ll;
        modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
        lp = lt - ll;
      }
      const bool owner = active && sub == 0;
      if (a.pix_stats && owner) a.pix_stats[i] = make_float2(S, SW);
      if (a.lp_pixel && owner) a.lp_pixel[i] = lp;
      const float val = owner ? lp : 0.0f;
      if (a.partial) {
        // float64 from here on: the per-image sums (~ -2e4 nats) feed a softmax over importance samples.
        // A tile holds pixels of at most two images (HW >= PPT on this route): n_first and n_first + 1.
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
    } else {
      if constexpr (!ST) {
      if constexpr (LATE_G) {
        if (it == 0) {  // the first tile's derivatives are formed: from here on the upstream gradient is needed
          pdl_wait();
          g = fetch_g(n, pix_this);
          if (t_cnt > 1) fetch(t + t_dir, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur, st_cur);
        }
      }
      const float rS = rcpa(S), rSW = rcpa(SW);
      float lt = 0.f, ll = 0.f;
      if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        const bool single = (MC % 2 == 1) && (prr == NPAIR - 1);
        f2 lg = ld_pair<AL>(rowp, m, single);
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        const f2 wp = ld_pair<AL>(auxp, m, single);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, m, px, a.plain != 0, PD != 0) - lt),
                 single ? 0.0f : expf(modl_logt(grow, M, m + 1, px, a.plain != 0, PD != 0) - lt));
          pi = pk(expf(ld_param(grow, m, PD != 0) - ll), single ? 0.0f : expf(ld_param(grow, m + 1, PD != 0) - ll));
        }
        const f2 gr = r * g;
        if (active) {
          st_pair<AL>(rowp, m, single, (r - pi) * g);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<AL>(rowp, j * M + m, single, ld_pair<AL>(rowp, j * M + m, single) * gr);
        }
      }
      }
      // hand the gradient tile to the TMA engine
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
      char* dst = reinterpret_cast<char*>(a.dparams) + t * TILE_F * (PD ? 2 : 4);
      if ((bytes & 15u) == 0) {
        if constexpr (PD == 1) {
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
          if constexpr (DIRECT)
            reinterpret_cast<unsigned short*>(dst)[q] = reinterpret_cast<const unsigned short*>(slot)[q];
          else if (PD)
            reinterpret_cast<unsigned short*>(dst)[q] = f32_to_bf16_bits(slot[q]);
          else
            reinterpret_cast<float*>(dst)[q] = slot[q];
        }
        __syncwarp();
      }
      if constexpr (NSLOT == 1) {
        if (it + 1 < t_cnt) {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          issue(t + t_dir, 0);
        }
      }
    }
  }
  if constexpr (BWD) {
    if (lane == 0) bulk_wait_all<0>();
  } else {
    if (a.partial && t_begin < t_end) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;  // image of the warp's last pixel-sample
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

template <int MC, int LPP, bool BWD, int NSLOT, int MAXT, int AR, int PD = 0, bool ST = false>
__global__ void __launch_bounds__(MAXT, 1) modl_tile_kernel(const ModlArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  tile_body<MC, LPP, BWD, NSLOT, AR, false, PD, ST>(a, smem_raw);
}

// ---- one step in one launch: forward -> grid barrier -> per-image sums, log-mean-exp, softmax weights -> grid barrier ->
// backward.  For training shapes small enough that launch boundaries and pipeline ramps dominate (BASELINE configs[0]:
// 131 MB of parameters, ~4 tiles per warp): no launch gaps, one ramp instead of three, and each warp's last forward tile
// is still in shared memory when its reversed backward run starts.  Cooperative launch (all CTAs co-resident).
struct StepArgs {
  ModlArgs a;
  StepFinish f;
};

template <int MC, int LPP, int AR>
__global__ void __launch_bounds__(512, 1) modl_step_kernel(const StepArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int lane = threadIdx.x & 31;
  const long long nwarps = blockDim.x >> 5;
  const long long gw = static_cast<long long>(blockIdx.x) * nwarps + (threadIdx.x >> 5);
  const long long total_warps = static_cast<long long>(gridDim.x) * nwarps;
  tile_body<MC, LPP, false, 1, AR, true, 0, true>(sa.a, smem_raw);   // writes a.pix_stats
  __threadfence();
  grid.sync();
  step_finish(sa.f, gw, total_warps, lane);
  __threadfence();
  grid.sync();
  asm volatile("fence.proxy.async;" ::: "memory");  // the forward pass's (S, SW) pairs are about to be read by bulk copies
  if (sa.f.elbo && gw == total_warps - 1) {  // batch mean, fixed order (the last warp owns the shortest run)
    double t = 0.0;
    for (long long b = lane; b < sa.f.B; b += 32) t += sa.f.lme64[b];
    t = warp_sum(t);
    if (lane == 0) {
      const float e = static_cast<float>(t / static_cast<double>(sa.f.b_norm));  // models/loss.py:37
      sa.f.elbo[0] = e;
      peer_publish(sa.f.peer, e);  // N > 1: the share goes straight into every rank's exchange buffer (NVLink P2P stores)
    }
  }
  tile_body<MC, LPP, true, 1, AR, true, 0, true>(sa.a, smem_raw);    // one-pass gradient from a.pix_stats
}

static __global__ void cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

}  // namespace vaemdl
