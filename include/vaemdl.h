/*
 * vaemdl.h -- C ABI of libvaemdl_b200.so: the B200 (sm_100a) observation-model
 * kernels that stand in for the hot path of nbip/vae-mdl.
 *
 * The reference has NO native / FFI boundary for this path (it is a TensorFlow
 * op graph built by five Python classes and two loss functions); every entry
 * point below cites the reference code (file:line, relative to the reference
 * root) whose computation it replaces.  The reference-side binding a maintainer
 * would add (a ctypes stub inside utils/mdl.py etc.) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - Plain C: raw pointers + sizes, no torch / C++ types.
 *   - Every *device* entry point is an asynchronous enqueue on `stream`
 *     (a cudaStream_t passed as void*); nothing is allocated inside; all
 *     buffers are caller-owned DEVICE memory unless the name ends in _host.
 *   - Returns 0 on success, a negative VAEMDL_E* code for argument errors, or a
 *     positive cudaError_t.  Never throws, never aborts.  Re-entrant.
 *   - Tensors are dense row-major float32 unless stated.
 *   - "image" n = one (importance-sample, batch) element of the flattened
 *     leading dims (s-major, b-minor: n = s*B + b, utils/mdl_openai_iwae.py:38-46).
 *     Image n is scored against x[n % x_batch]; x_batch = B for x [B,H,W,3],
 *     1 for the broadcast x [H,W,3] of models/model05.py:173.
 *   - MoDL parameter row per pixel (utils/mdl.py:98-108, utils/mdl_openai.py:90-94):
 *       [ logit(M) | muR(M) sR(M) kR(M) | muG(M) sG(M) kG(M) | muB(M) sB(M) kB(M) ]
 */
#ifndef VAEMDL_H_
#define VAEMDL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAEMDL_ABI_VERSION 1

#if defined(__GNUC__)
#define VAEMDL_API __attribute__((visibility("default")))
#else
#define VAEMDL_API
#endif

/* error codes (negative; positive values are cudaError_t) */
#define VAEMDL_OK 0
#define VAEMDL_EINVAL (-1)      /* bad argument (null pointer, non-positive size, unknown enum) */
#define VAEMDL_EALIGN (-2)      /* parameter / gradient pointer not 16-byte aligned */
#define VAEMDL_EUNSUPPORTED (-3) /* n_mix outside [1, VAEMDL_MAX_MIX] */
#define VAEMDL_EWORKSPACE (-4)  /* workspace too small */

#define VAEMDL_MAX_MIX 64

/* dtype of the observed image x */
#define VAEMDL_X_F32 0 /* float32 */
#define VAEMDL_X_U8 1  /* raw uint8 bytes k; the kernel forms k/255.f exactly as utils/data.py:15-16 */

/* value range of x */
#define VAEMDL_RANGE_UNIT 0 /* x in [0,1]; kernel applies x*2-1 (utils/mdl.py:65, utils/mdl_openai_iwae.py:35) */
#define VAEMDL_RANGE_SYM 1  /* x already in [-1,1] (utils/mdl_openai.py:31-32); only valid with VAEMDL_X_F32 */

/* which edge test selects the x=0 / x=255 branches */
#define VAEMDL_EDGE_MDL 0    /* x <= -1 , x >= 1        (utils/mdl.py:200-205)        */
#define VAEMDL_EDGE_OPENAI 1 /* x < -0.999 , x > 0.999  (utils/mdl_openai.py:139,:142) */

/* sampler variants */
#define VAEMDL_SAMPLE_OPENAI 0 /* select mixture first, one logistic draw per sub-pixel (utils/mdl_openai.py:160-193); u_log [n,H,W,3] */
#define VAEMDL_SAMPLE_MDL 1    /* a draw for every mixture, then select (utils/mdl.py:209-252); u_log [n,H,W,3,M] */
#define VAEMDL_SAMPLE_PLAIN 2  /* utils/mdl_plain.py:68-102: means chained on the means, channels drawn independently, clip
                                  to [-1,1]; u_log [n,H,W,3,M], or NULL for .mean() (:104-121: the selected locations) */

VAEMDL_API int vaemdl_version(void);
VAEMDL_API const char* vaemdl_strerror(int code);

/* ------------------------------------------------------------------------ *
 * Mixture of discretized logistics -- log-likelihood
 * replaces: MixtureDiscretizedLogistic._log_prob            utils/mdl.py:56-207
 *           discretized_mix_logistic_loss(sum_all=False)    utils/mdl_openai.py:83-157
 *           MixtureDiscretizedLogisticOpenaiIWAE._log_prob  utils/mdl_openai_iwae.py:33-67
 *           + the caller's reduce_sum over [-1,-2,-3]       models/loss.py:32
 *
 * params   [n_img, H, W, 10*M]
 * x        [x_batch, H, W, 3]  (float32 or uint8, see x_dtype / x_range)
 * lp_pixel [n_img, H, W]   nullable -- per-pixel log-prob (what log_prob() returns, minus the trailing 1)
 * ll_image [n_img]         nullable -- sum over H,W of lp_pixel
 * ll_image_f64 [n_img]     nullable -- the same sum in float64.  The per-pixel values are float32, but they are
 *                          accumulated in float64 (fixed order whenever an image holds at least one tile of pixels --
 *                          64 for n_mix 1..9, at most 32 otherwise -- float64 atomics for smaller images):
 *                          |ll| ~ 2e4 nats has a float32 ulp of 2e-3, which would go straight into the
 *                          softmax over importance samples of the IWAE gradient.  Feed this to vaemdl_iwae_tail.
 * workspace: 8-byte aligned, at least vaemdl_modl_workspace_bytes(n_img, H, W) bytes (used when a sum is requested)
 * ------------------------------------------------------------------------ */
VAEMDL_API size_t vaemdl_modl_workspace_bytes(long long n_img, int H, int W);

VAEMDL_API int vaemdl_modl_fwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    long long n_img, int x_batch, int H, int W, int M,
                    float* lp_pixel, float* ll_image, double* ll_image_f64,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ *
 * MoDL forward fused with the IWAE tail: TWO launches (forward, finish) when S <= 512 and an image holds at least one
 * tile of pixels, otherwise forward + per-image reduce + IWAE tail, for
 *     lpxz = reduce_sum(pxz.log_prob(x), [-1,-2,-3])                    models/loss.py:32
 *     log_w = lpxz + extra ; lme_b = logmeanexp(log_w, axis=0)          models/loss.py:34-37, utils/utils.py:9-11
 *     elbo = sum_b lme_b / B_total ; g_ll = d(-elbo)/d lpxz = -softmax_s(log_w) / B_total
 * params [S,B,H,W,10M]; extra [S,B] nullable (= beta*(lpz-lqzx)); B_total: whole-batch size when B is one rank's shard
 * (0 = B).  Outputs (all nullable): ll_image [S,B] float32, ll_image_f64 [S,B], log_w [S,B], lme_b [B], elbo [1]
 * (needs lme_b), g_ll [S,B] -- feed g_ll to vaemdl_modl_bwd(g_image=...).  Sums are float64 and in a fixed order.
 * workspace: as for vaemdl_modl_fwd with n_img = S*B.
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_iwae_fwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                    const float* extra,
                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ *
 * One IWAE step of the observation model in ONE call: forward, per-image sums, log-mean-exp over importance
 * samples, batch mean and the parameter gradient of the loss -elbo.
 * replaces: iwae_loss + tape.gradient restricted to pxz    models/loss.py:26-55 ; models/model05.py:141-145
 * Same arguments and outputs as vaemdl_modl_iwae_fwd (lme_b and g_ll required here) followed by
 * vaemdl_modl_bwd(g_image = g_ll) -> dparams [S, B, H, W, 10*M].
 * For S <= 32, n_mix in {5, 10, 20, 30} and problems of up to a few dozen tiles per warp (the training shapes of
 * models/model05.py: 5 x 64..128 x 32 x 32) this is ONE cooperative kernel launch -- forward pass, grid barrier, IWAE
 * finish, grid barrier, backward pass starting on the tile still resident in shared memory; anything else runs the
 * three launches of the calls above.  *launches (nullable) receives the number of kernel launches that were enqueued.
 * ------------------------------------------------------------------------ */
/* workspace for vaemdl_modl_iwae_step: the forward workspace plus one (mixture sum, logit normaliser) float pair per
 * pixel-sample, which the forward pass leaves for the backward pass so that the gradient of a component is scaled in the
 * same pass that forms it (n_mix in {5, 10, 20, 30}; with a buffer of only vaemdl_modl_workspace_bytes the step still
 * runs, with the two-pass gradient kernel). */
VAEMDL_API size_t vaemdl_modl_step_workspace_bytes(long long n_img, int H, int W);

VAEMDL_API int vaemdl_modl_iwae_step(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                    const float* extra,
                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                    float* dparams, void* workspace, size_t workspace_bytes, void* stream, int* launches);

/* ------------------------------------------------------------------------ *
 * Mixture of discretized logistics -- gradient w.r.t. params
 * replaces: tf.GradientTape over the ops above              models/model05.py:141-145
 * The upstream gradient on lp_pixel[n,h,w] is
 *     (g_image ? g_image[n] : 0) + (g_pixel ? g_pixel[n,h,w] : 0)
 * g_image [n_img] nullable, g_pixel [n_img,H,W] nullable (at least one non-null)
 * dparams [n_img, H, W, 10*M]  (fully overwritten)
 * The log-scale gradient is masked with (raw_log_scale >= -7) (tf.maximum, utils/mdl.py:109).
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_bwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    long long n_img, int x_batch, int H, int W, int M,
                    const float* g_image, const float* g_pixel,
                    float* dparams, void* stream);

/* The two halves of vaemdl_modl_iwae_step as separate calls: pix_stats [n_img*H*W*2] floats (8-byte aligned) is written
 * by the forward call and read by the backward call on the SAME params / x (n_mix in {5, 10, 20, 30}: one-pass gradient;
 * any other n_mix ignores it and forms the sums again).  pix_stats = NULL: exactly vaemdl_modl_iwae_fwd / vaemdl_modl_bwd. */
VAEMDL_API int vaemdl_modl_iwae_fwd_stats(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                    const float* extra,
                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                    float* pix_stats, void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_bwd_stats(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                    long long n_img, int x_batch, int H, int W, int M,
                    const float* g_image, const float* g_pixel, const float* pix_stats,
                    float* dparams, void* stream);

/* ------------------------------------------------------------------------ *
 * bfloat16 parameters (the decoder's last conv often runs in bf16: models/model05.py:76-90 under mixed precision)
 * Same calls as vaemdl_modl_fwd / _iwae_fwd / _bwd with params (and dparams) as bfloat16 [n_img, H, W, 10*M]:
 * half the DRAM bytes per px-sample.  The tile is widened to float32 in shared memory; every operation, the per-image
 * float64 sums and the outputs other than dparams are exactly those of the float32 entry points evaluated on the
 * widened parameters; dparams is rounded to nearest-even once, at the store.  params / dparams 16-byte aligned.
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_fwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                    long long n_img, int x_batch, int H, int W, int M,
                    float* lp_pixel, float* ll_image, double* ll_image_f64,
                    void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_iwae_fwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                    int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                    const float* extra,
                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                    void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_bwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                    long long n_img, int x_batch, int H, int W, int M,
                    const float* g_image, const float* g_pixel,
                    void* dparams_bf16, void* stream);
/* The same with the per-pixel mixture sums handed from the forward to the backward call (pix_stats [n_img*H*W*2] float32,
 * written by the forward kernel): for n_mix 10 / 20 / 30 the backward kernel then keeps the tile in bfloat16 in shared
 * memory (two slots per warp), widens every component pair as it reads it and writes the final gradient rounded ONCE --
 * the route that makes bfloat16 parameters faster than float32 ones.  Other n_mix ignore pix_stats. */
VAEMDL_API int vaemdl_modl_iwae_fwd_stats_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                         int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                         const float* extra,
                         float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                         float* pix_stats, void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_bwd_stats_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                          long long n_img, int x_batch, int H, int W, int M,
                          const float* g_image, const float* g_pixel, const float* pix_stats,
                          void* dparams_bf16, void* stream);

/* ------------------------------------------------------------------------ *
 * Pixel mixture of discretized logistics WITHOUT conditioning on the observed x
 * replaces: PixelMixtureDiscretizedLogistic.log_prob + get_mixture_params   utils/mdl_plain.py:36-66, :124-168
 * Same parameter row, same outputs and workspace as the entry points above; the only difference is the chain of
 * means: loc_g = mu_g + tanh(kR)*loc_r, loc_b = mu_b + tanh(kG)*loc_r + tanh(kB)*loc_g (utils/mdl_plain.py:160-162)
 * instead of the observed x_r, x_g.  x in [0,1] (the class rescales to [-1,1], :45); low / high / levels are the class's
 * constructor arguments (utils/mdl_plain.py:18; defaults -1, 1, 256): edge tests x <= low / x >= high, bin width
 * (high - low) / (levels - 1) (utils/discretized_logistic.py:18-21, :71-76).  Supported: bin width <= 0.049 (levels >= 42
 * on [-1,1]; the defaults give 0.0078) -- coarser grids return VAEMDL_EUNSUPPORTED: with log-scales clamped at -7 the
 * linear-domain product of the three sub-pixel terms would leave the float32 range.
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_plain_fwd(const float* params, const void* x, int x_dtype,
                    long long n_img, int x_batch, int H, int W, int M, float low, float high, float levels,
                    float* lp_pixel, float* ll_image, double* ll_image_f64,
                    void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_plain_iwae_fwd(const float* params, const void* x, int x_dtype,
                    int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                    float low, float high, float levels, const float* extra,
                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                    void* workspace, size_t workspace_bytes, void* stream);
VAEMDL_API int vaemdl_modl_plain_iwae_step(const float* params, const void* x, int x_dtype,
                         int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                         float low, float high, float levels, const float* extra,
                         float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                         float* dparams, void* workspace, size_t workspace_bytes, void* stream, int* launches);
VAEMDL_API int vaemdl_modl_plain_bwd(const float* params, const void* x, int x_dtype,
                    long long n_img, int x_batch, int H, int W, int M, float low, float high, float levels,
                    const float* g_image, const float* g_pixel,
                    float* dparams, void* stream);

/* ------------------------------------------------------------------------ *
 * Plain discretized logistic
 * replaces: DiscretizedLogistic.log_prob   utils/discretized_logistic.py:35-78
 *           + reduce_sum over [-1,-2,-3]   models/loss.py:32, models/model06.py:45
 *
 * loc, logscale: element e = (n, i) with i < D = H*W*C lives at  ptr[(n*D + i)/C*ld + (n*D+i)%C]
 *                i.e. a [.., C] tensor with channel-row stride `ld` floats
 *                (ld = C for separate tensors; ld = 2*C with logscale = loc + C for the
 *                 un-split [..,6] conv output of models/model03.py:88-91).
 * x  [x_batch, D] float32 or uint8 (uint8 => x = k/255.f); used as is (no rescale, :37).
 * lp_elem [n_img, D] nullable ; ll_image [n_img] nullable ; ll_image_f64 [n_img] nullable (float64 accumulation).
 * ------------------------------------------------------------------------ */
VAEMDL_API size_t vaemdl_dlogistic_workspace_bytes(long long n_img, long long D);

VAEMDL_API int vaemdl_dlogistic_fwd(const float* loc, const float* logscale, int C, int ld,
                         const void* x, int x_dtype, long long n_img, int x_batch, long long D,
                         float low, float high, float levels,
                         float* lp_elem, float* ll_image, double* ll_image_f64,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Plain discretized logistic forward fused with the IWAE tail (model03/04/06 loss: models/loss.py:32-37,
 * models/model06.py:45-50): same outputs and launch count as vaemdl_modl_iwae_fwd. loc/logscale [S,B,..] addressed as
 * above with n_img = S*B; extra [S,B] nullable (= every other term of log_w). */
VAEMDL_API int vaemdl_dlogistic_iwae_fwd(const float* loc, const float* logscale, int C, int ld,
                         const void* x, int x_dtype, int S, long long B, long long B_total, int x_batch, long long D,
                         float low, float high, float levels, const float* extra,
                         float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                         void* workspace, size_t workspace_bytes, void* stream);

/* One IWAE step of the plain discretized logistic in ONE call: vaemdl_dlogistic_iwae_fwd + vaemdl_dlogistic_bwd(g_image =
 * g_ll) (lme_b, g_ll, dloc, dlogscale required).  For image tensors (C = 3, dense [..,3] pairs or the un-split [..,6]
 * layout), S <= 32 and at most two 64-pixel tiles per resident warp (BASELINE configs[1]: 5 x 128 x 32 x 32 x 3) this is a
 * single cooperative launch that reads the parameters once and keeps the unscaled derivatives in registers across the
 * grid barriers; three launches otherwise (the two routes agree to float32 round-off).  *launches: nullable. */
VAEMDL_API int vaemdl_dlogistic_iwae_step(const float* loc, const float* logscale, int C, int ld,
                         const void* x, int x_dtype, int S, long long B, long long B_total, int x_batch, long long D,
                         float low, float high, float levels, const float* extra,
                         float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                         float* dloc, float* dlogscale, int ld_out,
                         void* workspace, size_t workspace_bytes, void* stream, int* launches);

/* dloc / dlogscale use the same (C, ld_out) addressing as loc / logscale. */
VAEMDL_API int vaemdl_dlogistic_bwd(const float* loc, const float* logscale, int C, int ld,
                         const void* x, int x_dtype, long long n_img, int x_batch, long long D,
                         float low, float high, float levels,
                         const float* g_image, const float* g_elem,
                         float* dloc, float* dlogscale, int ld_out, void* stream);

/* ------------------------------------------------------------------------ *
 * IWAE log-mean-exp over the importance-sample axis
 * replaces: logmeanexp(log_w, axis=0)   utils/utils.py:9-11
 *           iwae_loss tail              models/loss.py:34-43 ; models/model06.py:47-55
 * log_w [S, B] ; out_b [B]
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_logmeanexp_fwd(const float* log_w, int S, long long B, float* out_b, void* stream);
VAEMDL_API int vaemdl_logmeanexp_fwd_f64(const double* log_w, int S, long long B, float* out_b, void* stream);
/* dlog_w[s,b] = g_out[b] * softmax_s(log_w[:,b])  (gradient of utils/utils.py:9-11; no stop_gradient there) */
VAEMDL_API int vaemdl_logmeanexp_bwd(const float* log_w, const float* g_out, int S, long long B, float* dlog_w, void* stream);
VAEMDL_API int vaemdl_logmeanexp_bwd_f64(const double* log_w, const float* g_out, int S, long long B, float* dlog_w, void* stream);

/* Fused IWAE tail: log_w = ll + (extra ? extra : 0); lme_b = logmeanexp_s; elbo = sum_b lme_b / B_total;
 * g_ll[s,b] = d(-elbo)/d ll[s,b] = -softmax_s(log_w)[s,b] / B_total.       (models/loss.py:34-37)
 * B_total: the batch size the mean is taken over; 0 means B.  A rank holding a shard of B images of a batch of
 * B_total passes both, and the per-rank elbo values then simply add up to the global one.
 * ll [S,B] float32 or ll_f64 [S,B] float64 (exactly one may be NULL; ll_f64 wins when both are given);
 * extra [S,B] nullable (= beta*(lpz-lqzx)); outputs nullable: log_w [S,B], lme_b [B], elbo [1], g_ll [S,B].
 * Fixed summation order (bitwise reproducible). */
VAEMDL_API int vaemdl_iwae_tail(const float* ll, const double* ll_f64, const float* extra, int S, long long B,
                     long long B_total,
                     float* log_w, float* lme_b, float* elbo, float* g_ll, void* stream);

/* ------------------------------------------------------------------------ *
 * Latent-side terms of the importance log-weights (sums of Normal log-densities over the latent axis)
 * replaces: lpz / lqzx and beta*(lpz - lqzx)              models/loss.py:28-34
 *           (lpz2 - lqz2z1) + (lpz1z2 - lqz1x)            models/model06.py:40-47
 *           and their gradients w.r.t. z, loc, scale      (tf.GradientTape)
 * term t:  T_t[s,b] = sum_d log N(z[s,b,d]; loc[.,b,d], scale[.,b,d]);  z [S,B,D];  loc / scale [B,D], or [S,B,D] when
 *          params_per_sample != 0;  loc == scale == NULL: the standard normal.
 * fwd:  extra_out[s,b] = (extra_in ? extra_in[s,b] : 0) + sum_t weight_t * T_t[s,b]   -> the `extra` of the *_iwae_fwd
 *       entry points;  term_sums [n_terms,S,B] nullable (the lpz / lqzx metrics of models/loss.py:48-55).
 * bwd:  given g_extra [S,B] = d loss / d extra (= g_ll of the *_iwae_fwd entry points), per term (each nullable):
 *       dz [S,B,D], dloc / dscale (shape of loc / scale).  Terms that pass the SAME dz pointer (two densities of one z)
 *       accumulate into it.  One launch each; sums run in a fixed order.
 * ------------------------------------------------------------------------ */
#define VAEMDL_MAX_LATENT_TERMS 4
typedef struct {
  const float* z;
  const float* loc;   /* NULL: standard normal */
  const float* scale; /* NULL iff loc is NULL  */
  int D;
  int params_per_sample;
  float weight;
} vaemdl_latent_term;

VAEMDL_API int vaemdl_latent_terms_fwd(const vaemdl_latent_term* terms, int n_terms, int S, long long B,
                            const float* extra_in, float* extra_out, float* term_sums, void* stream);
VAEMDL_API int vaemdl_latent_terms_bwd(const vaemdl_latent_term* terms, int n_terms, int S, long long B,
                            const float* g_extra, float* const* dz, float* const* dloc, float* const* dscale,
                            void* stream);

/* Importance samples split across ranks (batch smaller than the rank count, or to balance): this rank holds
 * S_local of the S_total samples of every image.  replaces: logmeanexp over the FULL sample axis (utils/utils.py:9-11)
 * and the tail of iwae_loss (models/loss.py:34-37) when axis 0 is spread over a process group.
 *   vaemdl_iwae_split_local   : pair_out [2,B] float64 = (max_s log_w, sum_s exp(log_w - max)) over the local samples,
 *                               log_w = ll_f64 + (extra ? extra : 0)
 *   -- the caller all-gathers the pairs in rank order: pairs_all [world,2,B] (the path's ONE collective, 16*B bytes/rank) --
 *   vaemdl_iwae_split_combine : lme_b [B] = log-mean-exp over all S_total samples, elbo [1] = sum_b lme_b / B_total,
 *                               g_ll [S_local,B] = d(-elbo)/d ll of the LOCAL samples (-softmax over all samples / B_total),
 *                               log_w [S_local,B]; all outputs nullable.  Bit-identical lme / elbo on every rank. */
VAEMDL_API int vaemdl_iwae_split_local(const double* ll_f64, const float* extra, int S_local, long long B, double* pair_out,
                            void* stream);
VAEMDL_API int vaemdl_iwae_split_combine(const double* ll_f64, const float* extra, int S_local, long long B,
                              const double* pairs_all, int world, int S_total, long long B_total,
                              float* log_w, float* lme_b, float* elbo, float* g_ll, void* stream);

/* ------------------------------------------------------------------------ *
 * The step's one collective at N > 1, over peer memory
 * replaces: the batch mean of models/loss.py:37 when the batch is split over the GPUs of one box (each rank's step ends
 *           with its additive ELBO share sum_b lme_b / B_total; the loss is the sum of the shares).
 * Instead of a collective launch after the step, the kernel that forms the share stores it into EVERY rank's exchange buffer
 * (NVLink P2P stores through CUDA-IPC mappings): one 8-byte word {step sequence number << 32 | float bits} at
 * [seq % ring][rank] of a [ring][n_ranks] array of 64-bit words.
 *   vaemdl_peer_next     : attaches an exchange to the NEXT call of this host thread that produces an `elbo` output
 *                          (vaemdl_*_iwae_fwd*, vaemdl_*_iwae_step); that call publishes its share and detaches it.
 *                          slots[r] = rank r's buffer as mapped in this process (slots[rank] = the local buffer).
 *   vaemdl_peer_elbo_sum : waits (device-side, bounded) until the n_ranks words of step `seq` have arrived in THIS rank's
 *                          buffer and writes their sum, added in rank order, to out[0] (NaN on time-out or overrun).
 * ------------------------------------------------------------------------ */
#define VAEMDL_MAX_PEERS 8
typedef struct VaemdlPeer {
  unsigned long long* slots[VAEMDL_MAX_PEERS];
  int n_ranks, rank, ring;
  unsigned seq;
} VaemdlPeer;
VAEMDL_API int vaemdl_peer_next(const VaemdlPeer* peer);
VAEMDL_API int vaemdl_peer_elbo_sum(const unsigned long long* my_slots, int n_ranks, int ring, unsigned seq, float* out,
                         void* stream);

/* ------------------------------------------------------------------------ *
 * Samplers (explicit uniform noise; float64 internal arithmetic)
 * replaces: sample_from_discretized_mix_logistic   utils/mdl_openai.py:160-193 (explicit-noise lines :167, :185-186)
 *           MixtureDiscretizedLogistic._sample_n   utils/mdl.py:209-252
 *           DiscretizedLogistic.sample             utils/discretized_logistic.py:80-85
 * params [n_img,H,W,10M], re-used for n_rep consecutive blocks of noise (the reference tiles the parameter tensor n
 *        times instead: utils/mdl_openai.py:39-45, utils/mdl_openai_iwae.py:78-84; tfd sample(n));
 * u_mix [n_rep,n_img,H,W,M]; u_log [n_rep,n_img,H,W,3] (OPENAI) or [n_rep,n_img,H,W,3,M] (MDL); uniforms in (0,1)
 * outputs carry the leading [n_rep, n_img]:
 * x_out  [..,H,W,3] float32 nullable; out_range: VAEMDL_RANGE_SYM -> [-1,1], VAEMDL_RANGE_UNIT -> x*0.5+0.5
 * x_q    [..,H,W,3] uint8 nullable: rint(255*clip(x01,0,1))  (new-build definition, the reference never quantises)
 * idx    [..,H,W]   uint8 nullable: selected mixture
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_sample(const float* params, const float* u_mix, const float* u_log, int variant, int out_range,
                       long long n_rep, long long n_img, int H, int W, int M,
                       float* x_out, uint8_t* x_q, uint8_t* idx, void* stream);

/* PixelMixtureDiscretizedLogistic.sample / .mean (utils/mdl_plain.py:68-121) with the class's own low / high: the
 * VAEMDL_SAMPLE_PLAIN variant of vaemdl_modl_sample, the logistic draws clipped to [low, high]
 * (utils/discretized_logistic.py:83); u_log == NULL gives mean(): the selected locations clipped to [-1, 1] (:115). */
VAEMDL_API int vaemdl_modl_plain_sample(const float* params, const float* u_mix, const float* u_log, float low, float high,
                       int out_range, long long n_rep, long long n_img, int H, int W, int M,
                       float* x_out, uint8_t* x_q, uint8_t* idx, void* stream);

/* Plain discretized-logistic sampler (utils/discretized_logistic.py:80-85; models/model06.py:166 calls it on every
 * forward pass): x_out[e] = clip(loc + exp(logscale) * (log u - log(1 - u)), low, high), float32 arithmetic with the
 * accurate logf / log1pf / expf (|error| < ~3e-7 on unclipped values).  loc / logscale addressed as base[(e / C) * ld + e % C]. */
VAEMDL_API int vaemdl_dlogistic_sample(const float* loc, const float* logscale, int C, int ld, const float* u,
                            long long n_elem, float low, float high, float* x_out, void* stream);

/* ------------------------------------------------------------------------ *
 * Host-buffer step: the whole IWAE observation-model step through HOST memory.
 * params_host [S,B,H,W,10M] (pinned recommended), x_host uint8 [B,H,W,3],
 * extra_host [S,B] nullable, outputs: dparams_host [S,B,H,W,10M] nullable (forward only if NULL),
 * ll_host [S,B], lme_host [B], elbo_host [1].
 * Splits B into chunks, pipelines H2D copy / fwd / IWAE tail / bwd / D2H copy over internal
 * streams and device staging buffers (allocated on first use, cached per device, freed by
 * vaemdl_host_release).  Synchronous: returns when all outputs are in host memory.
 * replaces: one train_step's loss+grad on the observation model, models/model05.py:139-148.
 * ------------------------------------------------------------------------ */
VAEMDL_API int vaemdl_modl_iwae_step_host(const float* params_host, const uint8_t* x_host, const float* extra_host,
                               int S, int B, int H, int W, int M,
                               float* dparams_host, float* ll_host, float* lme_host, float* elbo_host,
                               int chunk_b /*0 = auto*/);
VAEMDL_API void vaemdl_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* VAEMDL_H_ */
