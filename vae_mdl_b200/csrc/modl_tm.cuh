#pragma once
// modl_tm.cuh -- the two-pass gradient kernel with the tile it works on held in TENSOR MEMORY.
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
//
// modl_tile_kernel<.., BWD> keeps ONE float32 tile per warp in shared memory (a second does not fit next to 16 warps): while a
// warp computes, none of its bytes are in flight, and while its gradient tile drains and the next tile lands, it waits (15 % of
// the warp samples of the headline backward kernel sit on the mbarrier / bulk-group waits).  Blackwell's 256 KB of tensor
// memory per SM are idle on this path (no contraction anywhere), and their geometry is exactly a warp's working set: a warp
// may address 32 TMEM lanes (its quarter) -- one lane per thread, i.e. per pixel row of the tile -- and, with twelve warps per
// CTA (three per quarter), 168 of the 512 columns of 32 bits: room for the 10 * MC floats a lane owns plus the W*P and W pairs
// the second pass takes over from the first.  So here
//   * the shared-memory slot X is only the landing / staging area of the TMA engine,
//   * the tile being worked on lives in tensor memory C (tcgen05.st / tcgen05.ld, shape 32x32b: thread i <-> TMEM lane i),
//     laid out as one 24-column block per component pair: [W*P | logits | mu s k (R) | mu s k (G) | mu s k (B) | W],
//     so a pair's twenty values arrive with three 8-column loads instead of ten 64-bit shared-memory loads,
//   * pass 1 (unscaled derivatives, mixture sums) runs entirely on C while X receives the NEXT tile,
//   * pass 2 swaps column block by column block: it takes the unscaled derivatives of tile k out of C, moves the parameters
//     of tile k+1 from X into the block of C it just emptied, and puts the final gradient of tile k into the positions of X
//     it just read; then X goes to global memory as one bulk store and is refilled late in pass 1 of tile k+1 (before its
//     last component pair: the store drains slowly while the memory system is saturated with writes, and a warp that asks
//     earlier blocks on it -- measured 309 ... 271 us for "before pair 0 ... 4").
// The arithmetic and its order are those of tile_body<.., BWD, 1, ..>: the gradients are bit-identical to that kernel's.
#include "modl_tile.cuh"

namespace vaemdl {

// ---- tensor memory: allocation and the 32x32b load / store shapes ---------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_dst) {  // one whole warp; writes the base address to smem_dst
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {  // one whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 24-column block of the calling thread's TMEM lane <-> 24 registers.  The load and its wait are ONE statement: the
// registers of a tcgen05.ld are only defined after tcgen05.wait::ld, and nothing else keeps the compiler from scheduling a
// consumer in between.
struct Blk {
  uint32_t r[24];
};
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, Blk& b) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%24+8];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%24+16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]),
        "=r"(b.r[8]), "=r"(b.r[9]), "=r"(b.r[10]), "=r"(b.r[11]), "=r"(b.r[12]), "=r"(b.r[13]), "=r"(b.r[14]), "=r"(b.r[15]),
        "=r"(b.r[16]), "=r"(b.r[17]), "=r"(b.r[18]), "=r"(b.r[19]), "=r"(b.r[20]), "=r"(b.r[21]), "=r"(b.r[22]), "=r"(b.r[23])
      : "r"(taddr)
      : "memory");
}
// The same in two statements, for a load issued one loop iteration ahead of its use: the wait names every register of the
// block as read-and-written, so no consumer can be scheduled between the load and the wait.
__device__ __forceinline__ void tmem_ld24_issue(uint32_t taddr, Blk& b) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%24+8];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%24+16];\n"
      : "=r"(b.r[0]), "=r"(b.r[1]), "=r"(b.r[2]), "=r"(b.r[3]), "=r"(b.r[4]), "=r"(b.r[5]), "=r"(b.r[6]), "=r"(b.r[7]),
        "=r"(b.r[8]), "=r"(b.r[9]), "=r"(b.r[10]), "=r"(b.r[11]), "=r"(b.r[12]), "=r"(b.r[13]), "=r"(b.r[14]), "=r"(b.r[15]),
        "=r"(b.r[16]), "=r"(b.r[17]), "=r"(b.r[18]), "=r"(b.r[19]), "=r"(b.r[20]), "=r"(b.r[21]), "=r"(b.r[22]), "=r"(b.r[23])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld24_wait(Blk& b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7]),
                 "+r"(b.r[8]), "+r"(b.r[9]), "+r"(b.r[10]), "+r"(b.r[11]), "+r"(b.r[12]), "+r"(b.r[13]), "+r"(b.r[14]),
                 "+r"(b.r[15]), "+r"(b.r[16]), "+r"(b.r[17]), "+r"(b.r[18]), "+r"(b.r[19]), "+r"(b.r[20]), "+r"(b.r[21]),
                 "+r"(b.r[22]), "+r"(b.r[23])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st24(uint32_t taddr, const Blk& b) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%24], {%0,%1,%2,%3,%4,%5,%6,%7};\n"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%24+8], {%8,%9,%10,%11,%12,%13,%14,%15};\n"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%24+16], {%16,%17,%18,%19,%20,%21,%22,%23};\n" ::"r"(b.r[0]),
      "r"(b.r[1]), "r"(b.r[2]), "r"(b.r[3]), "r"(b.r[4]), "r"(b.r[5]), "r"(b.r[6]), "r"(b.r[7]), "r"(b.r[8]), "r"(b.r[9]),
      "r"(b.r[10]), "r"(b.r[11]), "r"(b.r[12]), "r"(b.r[13]), "r"(b.r[14]), "r"(b.r[15]), "r"(b.r[16]), "r"(b.r[17]),
      "r"(b.r[18]), "r"(b.r[19]), "r"(b.r[20]), "r"(b.r[21]), "r"(b.r[22]), "r"(b.r[23]), "r"(taddr)
      : "memory");
}
// a register whose value does not matter (columns of a block nobody reads): defined for the compiler, no instruction
__device__ __forceinline__ uint32_t any_reg() {
  uint32_t r;
  asm("" : "=r"(r));
  return r;
}
// value pair q of a block (q = 0: W*P, 1: logits, 2 + j: parameter group 1 + j of the row, 11: W = exp(logit - max logit))
__device__ __forceinline__ f2 blk_get(const Blk& b, int q) { return pk(__uint_as_float(b.r[2 * q]), __uint_as_float(b.r[2 * q + 1])); }
__device__ __forceinline__ void blk_set(Blk& b, int q, f2 v) {
  b.r[2 * q] = __float_as_uint(lo(v));
  b.r[2 * q + 1] = __float_as_uint(hi(v));
}

// Twelve warps per CTA: three per TMEM lane quarter (and per scheduler, whose 16,384 registers then allow 168 per thread: the
// pipelined block loads need them), 168 of the 512 columns each -- room for seven 24-column blocks (MC <= 14).
constexpr int kTmWarps = 12;
constexpr int kTmColsPerWarp = 168;
template <int MC, int LPP>
constexpr bool tm_supported() {
  return (MC % 2 == 0) && (MC / 2) * 24 <= kTmColsPerWarp && Tile<MC, LPP>::ALIGNED;
}

// float32 parameters, one slot per warp, <= kTmWarps warps.  Rows are 40 M bytes with M even: any number of rows is a
// multiple of 16 bytes, so the ragged last tile of a problem travels by bulk copy like every other.
template <int MC, int LPP, int AR>
__device__ __forceinline__ void tile_body_tm(const ModlArgs& a, unsigned char* smem_raw) {
  using T = Tile<MC, LPP>;
  constexpr int M = T::M, PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F, NPAIR = T::NPAIR;
  static_assert(tm_supported<MC, LPP>(), "tensor-memory tile: even MC, at most seven component pairs per lane");
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
  // thread i of a warp <-> TMEM lane 32 * (warp % 4) + i; warps that share a lane quarter take kTmColsPerWarp columns each
  // (taken through a lane-0 shuffle: the block address feeds the uniform datapath, and this tells the compiler that it is warp-uniform)
  const uint32_t tm = __shfl_sync(kFull, *tmem_base_p + (static_cast<uint32_t>(32 * (warp & 3)) << 16) +
                                             static_cast<uint32_t>(kTmColsPerWarp * (warp >> 2)), 0);

  const long long gw = run_index(a, warp, nwarps);
  const bool lane_used = (lane / LPP) < PPT;
  const int p = lane_used ? (lane / LPP) : 0;  // idle lanes (LPP=3: lanes 30,31) shadow pixel 0 and never write X
  const int sub = lane % LPP;
  const int m0 = sub * MC;
  const int rot = (T::ROT && a.pair_rot) ? T::pair_rot(lane) : 0;
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const int t_cnt = static_cast<int>(t_end - t_begin);  // (a warp's run is far below 2^31 tiles)
  const bool rev = a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first();
  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  auto issue = [&](long long t) {  // tile t -> X
    if (lane == 0) {
      const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * 4;
      const uint32_t bytes = static_cast<uint32_t>(tile_rows(t)) * ROWF * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      if (a.bwd_hint & 1)
        bulk_g2s_hint(slot, src, bytes, bar, pol_first);
      else
        bulk_g2s(slot, src, bytes, bar);
    }
  };

  if (t_cnt > 0) {
    issue(t_first);

    // (image, pixel-in-image) of this lane's pixel-sample, advanced incrementally
    const long long step_n = PPT / a.HW;
    const int step_pix = static_cast<int>(PPT - step_n * a.HW);
    long long n_own = (t_first * PPT + p) / a.HW;
    int pix_own = static_cast<int>((t_first * PPT + p) - n_own * a.HW);
    // pixel and upstream gradient of this lane's pixel-sample in tile t; lanes past a ragged last tile shadow its first pixel
    auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, int& pix_out, PixRaw& raw, float& g_out,
                     bool want_g) {
      const long long n_first = __shfl_sync(kFull, n_lane, 0);
      const int pix_first = __shfl_sync(kFull, pix_lane, 0);
      const bool in = p < tile_rows(t);
      const long long n = in ? n_lane : n_first;
      const int pix = in ? pix_lane : pix_first;
      raw = load_pixel_raw(a, n, pix);
      g_out = 0.0f;
      if (want_g) {
        if (a.g_image) g_out = a.g_image[n];
        if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
      }
      n_out = n;
      pix_out = pix;
    };
    long long n_cur = 0;
    int pix_cur = 0;
    PixRaw raw_cur{};
    float g_cur = 0.0f;
    // the kernel is launched programmatically behind the finish kernel: the upstream gradient of the FIRST tile is read after
    // griddepcontrol.wait, which comes after that tile's first pass (as in tile_body, LATE_G)
    fetch(t_first, n_own, pix_own, n_cur, pix_cur, raw_cur, g_cur, false);

    int pp = p < tile_rows(t_first) ? p : 0;  // this lane's row of the tile in C (and, below, of the tile in X)
    float* rowp = slot + pp * ROWF;

    // first tile: X -> C, row by row (thread i moves its own row), and the maximum logit on the way
    float lmax = -INFINITY;
    mbar_wait(bar, 0u);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
      const int m = m0 + 2 * prr;
      Blk b;
      b.r[0] = b.r[1] = b.r[22] = b.r[23] = any_reg();
#pragma unroll
      for (int g = 0; g < 10; ++g) blk_set(b, 1 + g, ld_pair<true>(rowp, g * M + m, false));
      const f2 lg = blk_get(b, 1);
      lmax = fmaxf(lmax, fmaxf(lo(lg), hi(lg)));
      tmem_st24(tm + 24u * pr, b);
    }
    tmem_wait_st();
    lmax = group_max<LPP>(lmax, lane);
    __syncwarp();
    if (t_cnt > 1) issue(t_first + t_dir);  // every lane has read its row: X takes the second tile

    for (int it = 0; it < t_cnt; ++it) {
      const long long t = t_first + it * t_dir;
      const bool has_next = it + 1 < t_cnt;
      const bool active = lane_used && p < tile_rows(t);  // (shadow lanes never write X)
      const long long i = t * PPT + pp;  // this lane's pixel-sample
      const long long n = n_cur;
      float g = g_cur;
      Pixel px;
      decode_pixel<AR>(a, raw_cur, px);
      const int pix_this = pix_cur;
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
      if (it > 0 && has_next) fetch(t + t_dir, n_own, pix_own, n_cur, pix_cur, raw_cur, g_cur, true);
      // X holds (or will hold) the next tile: this lane's row of it
      const int pp_next = has_next ? (p < tile_rows(t + t_dir) ? p : 0) : 0;
      float* rowp_next = slot + pp_next * ROWF;
      rowp = slot + pp * ROWF;
      // X still holds the previous gradient tile, handed to the TMA engine at the end of the last iteration.  Its shared-memory
      // reads take a while when the memory system is saturated with writes, so the warp does not wait for them here: it asks
      // again after a.tm_refill component pairs of the first pass, and X then takes the next tile, which lands during the rest
      // of that pass
      const bool refill = it > 0 && has_next;

      // ---- pass 1, on tensor memory: unscaled derivatives in place, W*P into the block's first pair
      f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
      // (software pipeline: block pr + 1 is requested as soon as block pr's values are in the arithmetic, ahead of the store
      // of block pr's results: neither the load's latency nor the store's hold on its source registers is waited for)
      Blk b;
      tmem_ld24_issue(tm, b);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        if (refill && pr == a.tm_refill) {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          issue(t + t_dir);
        }
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
          P = pair_eval<true, true, Pixel, AR>(px, mu, sc, kp, u);
        else
          P = pair_eval<false, true, Pixel, AR>(px, mu, sc, kp, u);
        sumW2 = sumW2 + W;
        sumWP2 = fma2(W, P, sumWP2);
        Blk o;
        blk_set(o, 11, W);  // (the second pass takes W from here instead of forming it again)
        blk_set(o, 0, W * P);
        blk_set(o, 1, lg);
#pragma unroll
        for (int j = 0; j < 9; ++j) blk_set(o, 2 + j, u[j]);
        __syncwarp();
        if (pr + 1 < NPAIR) tmem_ld24_issue(tm + 24u * (pr + 1), b);
        tmem_st24(tm + 24u * pr, o);
      }
      if (refill && a.tm_refill >= NPAIR) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        issue(t + t_dir);
      }
      tmem_wait_st();
      const float S = group_sum<LPP>(lo(sumWP2) + hi(sumWP2), lane);
      const float SW = group_sum<LPP>(lo(sumW2) + hi(sumW2), lane);
      const bool tiny = !(S > kTinySum);  // also catches NaN
      const float* grow = param_row(a, i, ROWF);
      if (it == 0) {  // from here on the upstream gradient is needed
        pdl_wait();
        g = 0.0f;
        if (a.g_image) g = a.g_image[n];
        if (a.g_pixel) g += a.g_pixel[n * a.HW + pix_this];
        if (has_next) fetch(t + t_dir, n_own, pix_own, n_cur, pix_cur, raw_cur, g_cur, true);
      }
      const float rS = rcpa(S), rSW = rcpa(SW);
      float lt = 0.f, ll = 0.f;
      if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, false);
      __syncwarp();

      // ---- pass 2: C (derivatives of tile k) -> final gradient into X; X (parameters of tile k+1) -> C
      if (has_next) mbar_wait(bar, static_cast<uint32_t>(it + 1) & 1u);
      float lmax_next = -INFINITY;
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = (pr + rot >= NPAIR) ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        Blk b;
        tmem_ld24(tm + 24u * pr, b);
        if (has_next) {
          Blk nb;
          nb.r[0] = nb.r[1] = nb.r[22] = nb.r[23] = any_reg();
#pragma unroll
          for (int q = 0; q < 10; ++q) blk_set(nb, 1 + q, ld_pair<true>(rowp_next, q * M + m, false));
          const f2 nlg = blk_get(nb, 1);
          lmax_next = fmaxf(lmax_next, fmaxf(lo(nlg), hi(nlg)));
          tmem_st24(tm + 24u * pr, nb);
        }
        const f2 W = blk_get(b, 11);
        const f2 wp = blk_get(b, 0);
        f2 r = wp * rS;   // posterior responsibility of the component
        f2 pi = W * rSW;  // softmax(logits)
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, m, px, a.plain != 0, false) - lt), expf(modl_logt(grow, M, m + 1, px, a.plain != 0, false) - lt));
          pi = pk(expf(ld_param(grow, m, false) - ll), expf(ld_param(grow, m + 1, false) - ll));
        }
        const f2 gr = r * g;
        if (active) {
          st_pair<true>(rowp, m, false, (r - pi) * g);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<true>(rowp, j * M + m, false, blk_get(b, 1 + j) * gr);
        }
        __syncwarp();
      }
      tmem_wait_st();
      lmax = group_max<LPP>(lmax_next, lane);
      pp = pp_next;

      // hand the gradient tile to the TMA engine
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        char* dst = reinterpret_cast<char*>(a.dparams) + t * TILE_F * 4;
        const uint32_t bytes = static_cast<uint32_t>(tile_rows(t)) * ROWF * 4u;
        if (a.bwd_hint & 2)
          bulk_s2g_hint(dst, slot, bytes, pol_first);
        else
          bulk_s2g(dst, slot, bytes);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait_all<0>();
  } else {
    pdl_wait();
  }
  tmem_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc_512(*tmem_base_p);
}

template <int MC, int LPP, int MAXT, int AR>
__global__ void __launch_bounds__(MAXT, 1) modl_tile_tm_kernel(const ModlArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  tile_body_tm<MC, LPP, AR>(a, smem_raw);
}

}  // namespace vaemdl
