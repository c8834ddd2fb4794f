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
};

__global__ void __launch_bounds__(256) modl_sample_kernel(const SampleArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int M = a.M, ROWF = 10 * M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int TILE_F = 32 * ROWF;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * TILE_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * TILE_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const long long total_warps = static_cast<long long>(gridDim.x) * nwarps;
  const long long tiles_per_rep = (a.n_px + 31) / 32;
  const long long num_tiles = tiles_per_rep * a.n_rep;
  uint32_t parity = 0;
  for (long long t = static_cast<long long>(blockIdx.x) * nwarps + warp; t < num_tiles; t += total_warps) {
    const long long rep = t / tiles_per_rep;
    const long long tt = t - rep * tiles_per_rep;  // tiles never straddle two repetitions
    const long long rem = a.n_px - tt * 32;
    const int rows = rem < 32 ? static_cast<int>(rem) : 32;
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
    const float* src = a.params + tt * TILE_F;
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(slot, src, bytes, bar);
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32) slot[i] = src[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
    const bool active = lane < rows;
    const long long i = rep * a.n_px + tt * 32 + (active ? lane : 0);  // index into the noise / output tensors
    mbar_wait(bar, parity);
    parity ^= 1;
    const float* row = slot + (active ? lane : 0) * ROWF;
    // Gumbel-argmax over the mixture logits (utils/mdl_openai.py:167); first maximum wins
    int sel = 0;
    const float* um = a.u_mix + i * M;
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
    double xs[3];
    double coef[3];
    double noise_keep[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double mu = static_cast<double>(row[M + c * 3 * M + sel]);                              // :177
      const double ls = fmax(static_cast<double>(row[M + c * 3 * M + M + sel]), -7.0);              // :178-180
      coef[c] = tanh(static_cast<double>(row[M + c * 3 * M + 2 * M + sel]));                        // :181
      double noise = 0.0;  // u_log == NULL (plain variant only): the selected location itself (utils/mdl_plain.py:104-121)
      if (a.u_log) {
        const double u = static_cast<double>(a.variant_mdl ? a.u_log[(i * 3 + c) * M + sel] : a.u_log[i * 3 + c]);
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
      xo[0] = fmin(fmax(l0 + noise_keep[0], -1.0), 1.0);                                            // discretized_logistic.py:80-85
      xo[1] = fmin(fmax(l1 + noise_keep[1], -1.0), 1.0);
      xo[2] = fmin(fmax(l2 + noise_keep[2], -1.0), 1.0);
    }
    __syncwarp();  // every lane is done with the slot before the next bulk copy overwrites it
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

extern "C" int vaemdl_modl_sample(const float* params, const float* u_mix, const float* u_log, int variant,
                                  int out_range, long long n_rep, long long n_img, int H, int W, int M, float* x_out,
                                  uint8_t* x_q, uint8_t* idx, void* stream) {
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
  const DeviceInfo& di = device_info();
  const size_t tile_b = static_cast<size_t>(32) * 10 * M * 4;
  int warps = 8;
  while (warps > 1 && warps * tile_b + warps * 8 > static_cast<size_t>(di.max_smem_optin) / 2) --warps;
  const size_t smem = warps * tile_b + warps * 8;
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
  const long long cap = static_cast<long long>(di.sm_count) * 2;
  if (grid > cap) grid = cap;
  modl_sample_kernel<<<static_cast<unsigned>(grid), warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
  return cuda_rc(cudaGetLastError());
}
