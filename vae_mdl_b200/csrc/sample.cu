// sample.cu -- explicit-noise MoDL sampler.
//
// Replaces sample_from_discretized_mix_logistic (utils/mdl_openai.py:160-193, explicit-noise lines :167 and :185-186)
// and MixtureDiscretizedLogistic._sample_n (utils/mdl.py:209-252).  The logistic draw, the clipping chain and the
// quantiser run in float64 so that the quantised pixel value is reproducible against the float64 oracle on identical
// uniforms.  The Gumbel-argmax runs in float32 (accurate logf, |error| < 4e-6 per Gumbel value) while the winner leads by
// more than 1e-4; a lane whose two best values are closer than that repeats the search in float64, so the selected
// index is the float64 one in every case -- at a fraction of the float64 work (2M of the 2M + 12 transcendentals/pixel).
//
// One warp per tile of 32 pixels: the tile's parameter rows arrive through a TMA bulk copy into shared memory (same
// scheme as modl_kernels.cu); every lane then owns one pixel.
#include <mutex>

#include "common.cuh"

namespace vaemdl {

struct SampleArgs {
  const float* params;
  const float* u_mix;
  const float* u_log;
  float* x_out;
  uint8_t* x_q;
  uint8_t* idx;
  long long n_px;   // pixels of ONE repetition (n_img * H * W)
  long long n_rep;  // the parameters are re-used for n_rep consecutive blocks of noise / output
  int M;
  int variant_mdl;    // u_log carries a draw for every mixture: [.., 3, M]
  int variant_plain;  // utils/mdl_plain.py: means chained on the means, no sequential dependence between the channels
  int out_unit;
  double clip_lo, clip_hi;  // plain variant: DiscretizedLogistic.sample clips to the class's [low, high]
                            // (utils/discretized_logistic.py:83); its mean() to [-1, 1] (utils/mdl_plain.py:115)
};

// Per-warp shared-memory slot: [ parameter tile: 32 rows x 10M | u_mix tile: 32 x M | u_log tile: 32 x 3 (x M) ], each
// brought in by its own 1-D TMA bulk copy on one mbarrier.  The slot is handed back to the TMA engine (next tile) as
// soon as the lane has picked its component and copied the twelve numbers it still needs into registers, so the
// float64 part overlaps the next tile's loads.
__global__ void __launch_bounds__(512) modl_sample_kernel(const SampleArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int M = a.M, ROWF = 10 * M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int TILE_F = 32 * ROWF;
  const int UM_F = 32 * M;
  const int UL_PER = a.variant_mdl ? 3 * M : 3;  // floats of logistic noise per pixel
  const int UL_F = a.u_log ? 32 * UL_PER : 0;
  const int WARP_F = TILE_F + UM_F + UL_F;      // all three pieces are multiples of 4 floats
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* um_s = slot + TILE_F;
  float* ul_s = um_s + UM_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const long long total_warps = static_cast<long long>(gridDim.x) * nwarps;
  const long long tiles_per_rep = (a.n_px + 31) / 32;
  const long long num_tiles = tiles_per_rep * a.n_rep;

  struct TileId {
    long long rep, tt, i0;
    int rows;
  };
  auto locate = [&](long long t) {
    TileId d;
    d.rep = t / tiles_per_rep;
    d.tt = t - d.rep * tiles_per_rep;  // tiles never straddle two repetitions
    const long long rem = a.n_px - d.tt * 32;
    d.rows = rem < 32 ? static_cast<int>(rem) : 32;
    d.i0 = d.rep * a.n_px + d.tt * 32;  // first pixel of the tile in the noise / output tensors
    return d;
  };
  auto issue = [&](const TileId& d) {
    const uint32_t pb = static_cast<uint32_t>(d.rows) * ROWF * 4u;
    const uint32_t mb = static_cast<uint32_t>(d.rows) * M * 4u;
    const uint32_t lb = a.u_log ? static_cast<uint32_t>(d.rows) * UL_PER * 4u : 0u;
    const float* psrc = a.params + d.tt * TILE_F;
    const float* msrc = a.u_mix + d.i0 * M;
    const float* lsrc = a.u_log ? a.u_log + d.i0 * UL_PER : nullptr;
    const bool bulk = ((pb | mb | lb) & 15u) == 0 && ((reinterpret_cast<uintptr_t>(msrc) | reinterpret_cast<uintptr_t>(lsrc)) & 15u) == 0;
    if (bulk) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, pb + mb + lb);
        bulk_g2s(slot, psrc, pb, bar);
        bulk_g2s(um_s, msrc, mb, bar);
        if (lb) bulk_g2s(ul_s, lsrc, lb, bar);
      }
    } else {  // ragged last tile / unaligned noise: plain loads
      for (int k = lane; k < d.rows * ROWF; k += 32) slot[k] = psrc[k];
      for (int k = lane; k < d.rows * M; k += 32) um_s[k] = msrc[k];
      if (a.u_log)
        for (int k = lane; k < d.rows * UL_PER; k += 32) ul_s[k] = lsrc[k];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };

  long long t = static_cast<long long>(blockIdx.x) * nwarps + warp;
  if (t < num_tiles) issue(locate(t));
  uint32_t parity = 0;
  for (; t < num_tiles; t += total_warps) {
    const TileId d = locate(t);
    const bool active = lane < d.rows;
    const int lr = active ? lane : 0;
    const long long i = d.i0 + lr;
    mbar_wait(bar, parity);
    parity ^= 1;
    const float* row = slot + lr * ROWF;
    // Gumbel-argmax over the mixture logits (utils/mdl_openai.py:167); first maximum wins
    int sel = 0;
    const float* um = um_s + lr * M;
    {
      float best = -INFINITY, second = -INFINITY;
      for (int m = 0; m < M; ++m) {
        const float gmb = row[m] - logf(-logf(um[m]));
        if (gmb > best) {
          second = best;
          best = gmb;
          sel = m;
        } else if (gmb > second) {
          second = gmb;
        }
      }
      if (!(best - second > 1e-4f)) {  // too close to call in float32 (or NaN): decide in float64
        double best64 = -INFINITY;
        sel = 0;
        for (int m = 0; m < M; ++m) {
          const double gmb = static_cast<double>(row[m]) - log(-log(static_cast<double>(um[m])));
          if (gmb > best64) {
            best64 = gmb;
            sel = m;
          }
        }
      }
    }
    // the twelve numbers the float64 part needs, out of shared memory
    float p_mu[3], p_ls[3], p_k[3], p_u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p_mu[c] = row[M + c * 3 * M + sel];                                                           // :177
      p_ls[c] = row[M + c * 3 * M + M + sel];                                                       // :178-180
      p_k[c] = row[M + c * 3 * M + 2 * M + sel];                                                    // :181
      p_u[c] = a.u_log ? (a.variant_mdl ? ul_s[(lr * 3 + c) * M + sel] : ul_s[lr * 3 + c]) : 0.5f;
    }
    __syncwarp();  // every lane is done with the slot: the next tile may stream in while the float64 math runs
    if (t + total_warps < num_tiles) issue(locate(t + total_warps));

    double xs[3];
    double coef[3];
    double noise_keep[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double mu = static_cast<double>(p_mu[c]);
      const double ls = fmax(static_cast<double>(p_ls[c]), -7.0);
      coef[c] = tanh(static_cast<double>(p_k[c]));
      double noise = 0.0;  // u_log == NULL (plain variant only): the selected location itself (utils/mdl_plain.py:104-121)
      if (a.u_log) {
        const double u = static_cast<double>(p_u[c]);
        noise = exp(ls) * log(u / (1.0 - u));                                                       // :185-186 (log u - log(1-u))
      }
      xs[c] = a.variant_plain ? mu : mu + noise;  // plain variant: the noise goes on top of the chained means below
      noise_keep[c] = noise;
    }
    double xo[3];
    if (!a.variant_plain) {
      const double x0 = fmin(fmax(xs[0], -1.0), 1.0);                                               // :190
      const double x1 = fmin(fmax(xs[1] + coef[0] * x0, -1.0), 1.0);                                // :191
      const double x2 = fmin(fmax(xs[2] + coef[1] * x0 + coef[2] * x1, -1.0), 1.0);                 // :192
      xo[0] = x0;
      xo[1] = x1;
      xo[2] = x2;
    } else {
      const double l0 = xs[0];                                                                      // utils/mdl_plain.py:160
      const double l1 = xs[1] + coef[0] * l0;                                                       // :161
      const double l2 = xs[2] + coef[1] * l0 + coef[2] * l1;                                        // :162
      xo[0] = fmin(fmax(l0 + noise_keep[0], a.clip_lo), a.clip_hi);                                 // discretized_logistic.py:80-85
      xo[1] = fmin(fmax(l1 + noise_keep[1], a.clip_lo), a.clip_hi);
      xo[2] = fmin(fmax(l2 + noise_keep[2], a.clip_lo), a.clip_hi);
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double x01 = xo[c] * 0.5 + 0.5;  // utils/mdl.py:250, utils/mdl_openai_iwae.py:99
        if (a.x_out) a.x_out[i * 3 + c] = static_cast<float>(a.out_unit ? x01 : xo[c]);
        if (a.x_q) a.x_q[i * 3 + c] = static_cast<uint8_t>(rint(255.0 * fmin(fmax(x01, 0.0), 1.0)));
      }
      if (a.idx) a.idx[i] = static_cast<uint8_t>(sel);
    }
  }
}

}  // namespace vaemdl

using namespace vaemdl;

static int modl_sample_impl(const float* params, const float* u_mix, const float* u_log, int variant, int out_range,
                            long long n_rep, long long n_img, int H, int W, int M, float* x_out, uint8_t* x_q, uint8_t* idx,
                            void* stream, double clip_lo, double clip_hi) {
  if (!params || !u_mix || n_rep <= 0 || n_img <= 0 || H <= 0 || W <= 0) return VAEMDL_EINVAL;
  if (!u_log && variant != VAEMDL_SAMPLE_PLAIN) return VAEMDL_EINVAL;
  if (!x_out && !x_q && !idx) return VAEMDL_EINVAL;
  if (variant != VAEMDL_SAMPLE_OPENAI && variant != VAEMDL_SAMPLE_MDL && variant != VAEMDL_SAMPLE_PLAIN) return VAEMDL_EINVAL;
  if (out_range != VAEMDL_RANGE_UNIT && out_range != VAEMDL_RANGE_SYM) return VAEMDL_EINVAL;
  if (M < 1 || M > VAEMDL_MAX_MIX) return VAEMDL_EUNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(params) & 15u) return VAEMDL_EALIGN;
  SampleArgs a{};
  a.params = params;
  a.u_mix = u_mix;
  a.u_log = u_log;
  a.x_out = x_out;
  a.x_q = x_q;
  a.idx = idx;
  a.n_px = n_img * H * W;
  a.n_rep = n_rep;
  a.M = M;
  a.variant_mdl = variant != VAEMDL_SAMPLE_OPENAI;
  a.variant_plain = variant == VAEMDL_SAMPLE_PLAIN;
  a.out_unit = out_range == VAEMDL_RANGE_UNIT;
  a.clip_lo = clip_lo;
  a.clip_hi = clip_hi;
  const DeviceInfo& di = device_info();
  const size_t ul_per = u_log ? (a.variant_mdl ? 3 * M : 3) : 0;
  const size_t per_warp = (static_cast<size_t>(32) * 10 * M + 32 * M + 32 * ul_per) * 4 + 8;
  int warps = 16;  // one CTA per SM, as many warps as shared memory allows
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  const size_t smem = warps * per_warp;
  if (smem > static_cast<size_t>(di.max_smem_optin)) return VAEMDL_EUNSUPPORTED;
  {
    static std::mutex mu;
    static size_t c_smem = 0;
    static int c_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || c_smem < smem) {
      cudaError_t e = cudaFuncSetAttribute(modl_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long num_tiles = ((a.n_px + 31) / 32) * n_rep;
  long long grid = (num_tiles + warps - 1) / warps;
  const long long cap = static_cast<long long>(di.sm_count);
  if (grid > cap) grid = cap;
  modl_sample_kernel<<<static_cast<unsigned>(grid), warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
  return cuda_rc(cudaGetLastError());
}

extern "C" int vaemdl_modl_sample(const float* params, const float* u_mix, const float* u_log, int variant,
                                  int out_range, long long n_rep, long long n_img, int H, int W, int M, float* x_out,
                                  uint8_t* x_q, uint8_t* idx, void* stream) {
  return modl_sample_impl(params, u_mix, u_log, variant, out_range, n_rep, n_img, H, W, M, x_out, x_q, idx, stream, -1.0, 1.0);
}

/* PixelMixtureDiscretizedLogistic.sample / .mean with the class's own low / high (utils/mdl_plain.py:68-121): the logistic
 * draws are clipped to [low, high] (utils/discretized_logistic.py:83), the mean (u_log == NULL) to [-1, 1] (:115). */
extern "C" int vaemdl_modl_plain_sample(const float* params, const float* u_mix, const float* u_log, float low, float high,
                                        int out_range, long long n_rep, long long n_img, int H, int W, int M, float* x_out,
                                        uint8_t* x_q, uint8_t* idx, void* stream) {
  if (!(high > low)) return VAEMDL_EINVAL;
  const bool mean = u_log == nullptr;
  return modl_sample_impl(params, u_mix, u_log, VAEMDL_SAMPLE_PLAIN, out_range, n_rep, n_img, H, W, M, x_out, x_q, idx, stream,
                          mean ? -1.0 : static_cast<double>(low), mean ? 1.0 : static_cast<double>(high));
}
