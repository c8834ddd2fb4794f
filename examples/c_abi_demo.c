/* c_abi_demo.c -- the C ABI of include/vaemdl.h from plain C: no Python, no PyTorch.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o examples/c_abi_demo \
 *       -L vae_mdl_b200 -lvaemdl_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/vae_mdl_b200
 *   ./examples/c_abi_demo [S B H W n_mix]
 *
 * One IWAE step of the observation model (models/loss.py:26-55 + the gradient of models/model05.py:141-145) on
 * deterministic pseudo-random parameters: prints the loss, two per-image log-likelihoods, a gradient checksum and the
 * number of kernel launches.  tests/test_c_abi_demo_gpu.py runs it and compares with the Python layer. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "vaemdl.h"

#define CK(e)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (e);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

static uint64_t rng_state = 88172645463325252ull;
static double uniform01(void) { /* xorshift64*: the test regenerates the same stream in Python */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  const int S = argc > 5 ? atoi(argv[1]) : 5, B = argc > 5 ? atoi(argv[2]) : 4, H = argc > 5 ? atoi(argv[3]) : 32,
            W = argc > 5 ? atoi(argv[4]) : 32, M = argc > 5 ? atoi(argv[5]) : 10;
  const long long n_img = (long long)S * B;
  const size_t n_param = (size_t)n_img * H * W * 10 * M, n_x = (size_t)B * H * W * 3;
  float* params = (float*)malloc(n_param * sizeof(float));
  uint8_t* x = (uint8_t*)malloc(n_x);
  float* extra = (float*)malloc((size_t)n_img * sizeof(float));
  if (!params || !x || !extra) return 1;
  for (size_t i = 0; i < n_param; ++i) params[i] = (float)(4.0 * uniform01() - 2.0);
  for (size_t i = 0; i < n_x; ++i) x[i] = (uint8_t)(256.0 * uniform01());
  for (long long i = 0; i < n_img; ++i) extra[i] = (float)(2.0 * uniform01() - 1.0);

  float *d_params, *d_extra, *d_lme, *d_elbo, *d_gll, *d_dparams;
  double* d_ll64;
  uint8_t* d_x;
  void* d_ws;
  const size_t ws_bytes = vaemdl_modl_step_workspace_bytes(n_img, H, W);
  CK(cudaMalloc((void**)&d_params, n_param * sizeof(float)));
  CK(cudaMalloc((void**)&d_dparams, n_param * sizeof(float)));
  CK(cudaMalloc((void**)&d_x, n_x));
  CK(cudaMalloc((void**)&d_extra, (size_t)n_img * sizeof(float)));
  CK(cudaMalloc((void**)&d_ll64, (size_t)n_img * sizeof(double)));
  CK(cudaMalloc((void**)&d_gll, (size_t)n_img * sizeof(float)));
  CK(cudaMalloc((void**)&d_lme, (size_t)B * sizeof(float)));
  CK(cudaMalloc((void**)&d_elbo, sizeof(float)));
  CK(cudaMalloc(&d_ws, ws_bytes));
  CK(cudaMemcpy(d_params, params, n_param * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, x, n_x, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_extra, extra, (size_t)n_img * sizeof(float), cudaMemcpyHostToDevice));

  int launches = 0;
  const int rc = vaemdl_modl_iwae_step(d_params, d_x, VAEMDL_X_U8, VAEMDL_RANGE_UNIT, VAEMDL_EDGE_MDL, S, B, 0, B, H, W, M,
                                       d_extra, NULL, d_ll64, NULL, d_lme, d_elbo, d_gll, d_dparams, d_ws, ws_bytes,
                                       NULL /* default stream */, &launches);
  if (rc != 0) {
    fprintf(stderr, "vaemdl_modl_iwae_step: %s (%d)\n", vaemdl_strerror(rc), rc);
    return 3;
  }
  CK(cudaDeviceSynchronize());
  float elbo = 0.f;
  double ll_first = 0.0, ll_last = 0.0;
  float* grads = (float*)malloc(n_param * sizeof(float));
  CK(cudaMemcpy(&elbo, d_elbo, sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ll_first, d_ll64, sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ll_last, d_ll64 + (n_img - 1), sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(grads, d_dparams, n_param * sizeof(float), cudaMemcpyDeviceToHost));
  double gsum = 0.0, gabs = 0.0;
  for (size_t i = 0; i < n_param; ++i) {
    gsum += grads[i];
    gabs += fabs((double)grads[i]);
  }
  printf("abi %d S %d B %d H %d W %d M %d launches %d\n", vaemdl_version(), S, B, H, W, M, launches);
  printf("loss %.9g\nll_first %.12g\nll_last %.12g\ngrad_sum %.9g\ngrad_abs %.9g\n", -(double)elbo, ll_first, ll_last, gsum, gabs);
  return 0;
}
