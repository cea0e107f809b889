/*
 * pbnn_b200 -- C ABI of the B200-native (sm_100a) implementation of pbnn's data-parallel hot path:
 * many independent SG-MCMC chains / ensemble members advancing over small MLPs.
 *
 * Every entry point replaces a Python-level interface of the reference (bstaber/pbnn); the
 * reference has no native boundary of its own (it is pure Python over JAX/BlackJAX), so each
 * function cites the reference call it stands in for.  This is the boundary a `jax.ffi` /
 * ctypes / cffi binding attaches to.
 *
 * Conventions
 *  - all data pointers are DEVICE pointers (cudaMalloc'd / torch CUDA tensors), float32 / uint32 /
 *    int32, C-contiguous; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - no allocation, re-entrant across streams, nothing read from the environment; the only process-wide state is
 *    the option table written EXPLICITLY through pbnn_set_option() (engine selection for tests / experiments;
 *    with no option set the library picks engines from the arguments alone) plus per-device caches of immutable
 *    device properties (SM count, kernel shared-memory attribute);
 *  - return 0 on success, a negative PBNN_E* code otherwise; pbnn_last_error() gives the text
 *    (thread-local);
 *  - parameters are FLAT vectors in jax.flatten_util.ravel_pytree order of the Flax tree
 *    {"Dense_k": {"bias": (out,), "kernel": (in,out)}} (bias first, kernel row-major [in][out]);
 *  - rng keys are legacy uint32[2] threefry2x32 keys, jax_threefry_partitionable=True semantics;
 *  - chain c of a call == one reference call with (keys[c], theta[c]); chains never interact.
 */
#ifndef PBNN_B200_H
#define PBNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBNN_MAX_LAYERS 8

#define PBNN_OK 0
#define PBNN_EINVAL (-1)   /* bad argument                                   */
#define PBNN_ETOOBIG (-2)  /* model / batch does not fit the resident engine */
#define PBNN_ECUDA (-3)    /* CUDA runtime error                             */

#define PBNN_ACT_RELU 0
#define PBNN_ACT_TANH 1

/* likelihoods of benchmark/pipelines/logprobs.py:
 *   homoscedastic  (:12-17)  -sum_o 0.5 (y_o - f_o)^2 / noise_level^2, y has dims[n_layers] columns
 *   heteroscedastic (:37-42) network outputs (mu, log sigma^2) (dims[n_layers] == 2), y has ONE column:
 *                            -0.5 log sigma^2 - 0.5 (y - mu)^2 / sigma^2   (noise_level is ignored) */
#define PBNN_LIK_HOMOSCEDASTIC 0
#define PBNN_LIK_HETEROSCEDASTIC 1

/* minibatch semantics of the SG samplers */
#define PBNN_MB_REFERENCE_FIXED 0 /* the draw the reference's traced lax.scan re-uses for all steps */
#define PBNN_MB_PER_STEP 1        /* a fresh draw of the same generator at every step               */

/* Describes `pbnn.models.MLP` (src/pbnn/models.py:9-35, generalised to a list of widths) together with
 * `homoscedastic_loglike_fn` / `logprior_fn` (benchmark/pipelines/logprobs.py:6-17). */
typedef struct pbnn_mlp_spec {
  int32_t n_layers;                  /* number of Dense layers, 1..PBNN_MAX_LAYERS            */
  int32_t dims[PBNN_MAX_LAYERS + 1]; /* dims[0]=input width d, dims[l+1]=features of Dense_l  */
  int32_t activation;                /* PBNN_ACT_RELU | PBNN_ACT_TANH (hidden layers)         */
  float noise_level;                 /* sigma of the Gaussian likelihood                      */
  float prior_scale;                 /* std of the N(0, s^2) prior; the reference uses 1      */
  int32_t likelihood;                /* PBNN_LIK_HOMOSCEDASTIC | PBNN_LIK_HETEROSCEDASTIC       */
} pbnn_mlp_spec;

int pbnn_version(void);
const char* pbnn_last_error(void);
/* Engine-selection / tuning options (process-wide, explicit).  Names: "no_tc" (no tensor-core engines), "no_tc2"
 * (no two-hidden-layer tcgen05 engine), "no_h1" (no fused one-hidden-layer path), "no_resident", "force_gstate",
 * "no_sched" / "seg_len" (balanced HMC schedule), "nw", "bt", "dw_ks", "two_cta", "one_cta", "pred_tpc" (1, 2 or 4),
 * "pred_chunks", "tc2_min_h", "tc2_min_rows", "tc2_ctas", "tc2_nz", "tc2_nohelp", "no_w1" (no warp-per-chain engine),
 * "w1_min_chains" (chain count from which that engine is chosen; default 297, 520 for SGLD on relu nets), "no_w1c" (no
 * four-warps-per-chain variant for few chains), "no_w2" (no four-warp Adam kernel),
 * "debug".  pbnn_get_option returns 1 (and the sentinel INT_MIN in *value) when the option is unset.  Every engine computes the same arithmetic contract; options never change
 * results beyond the stated tolerances. */
int pbnn_set_option(const char* name, int value);
int pbnn_get_option(const char* name, int* value);
void pbnn_reset_options(void);
/* number of parameters P of the flat vector; <0 on invalid spec */
int pbnn_num_params(const pbnn_mlp_spec* spec);
/* launch plan the library would use: out[0]=threads per CTA, out[1]=batch-tile rows,
 * out[2]=dynamic shared memory bytes, out[3]=padded parameter count.  `rows` is the minibatch size
 * (or N for full-batch HMC / M for predict), `n_state` the number of resident parameter-sized arrays. */
int pbnn_plan_query(const pbnn_mlp_spec* spec, int rows, int n_state, int32_t out[4]);
/* gradient engine sampler `op` (PBNN_OP_*) would run for `C` chains with `rows` rows per gradient evaluation:
 * PBNN_ENGINE_* below, < 0 on error.  All engines implement the same arithmetic contract (the value_and_grad of
 * benchmark/pipelines/logprobs.py:6-17); the tcgen05 engine computes the two GEMMs as split-TF32 (hi/lo) tensor-core
 * products with FP32 accumulation (~1e-6 relative to the FP32 engines). */
#define PBNN_ENGINE_TILED 0    /* register-tiled FP32 GEMMs, any depth / width                                   */
#define PBNN_ENGINE_FUSED_H1 1 /* one hidden layer, scalar output: warp-autonomous fused FP32 path                */
#define PBNN_ENGINE_TCGEN05 2  /* HMC, one hidden relu layer, scalar output: tensor-core engine (pbnn_tc.cuh)     */
#define PBNN_ENGINE_TILED_GLOBAL_STATE 3 /* as TILED with the parameter-sized arrays in the global workspace      */
#define PBNN_ENGINE_TCGEN05_MLP2 4 /* SG samplers / Adam, two equal hidden relu layers, scalar output, B <= 128:
                                      bf16x3 tensor-core engine with TMA-streamed weights (pbnn_tc2.cuh)           */
#define PBNN_ENGINE_WARP_H1 5  /* SG samplers / Adam, one hidden layer, scalar output, many chains: one warp per chain,
                                  weights in registers for the whole run (pbnn_warp.cuh)                            */
#define PBNN_ENGINE_CTA4_H1 7   /* SGLD, one hidden relu layer, scalar output, FEW chains: the warp engine with a CTA of four
                                   warps per chain (replicated weights, split forward / backward / noise)            */
#define PBNN_ENGINE_WARP_MLP2 6 /* Adam (deep ensembles, MAP), two equal hidden relu layers, H <= 64, d + 1 <= 16,
                                   B <= 32: one warp per member, weights in its shared-memory block (pbnn_warp2.cuh) */
int pbnn_plan_engine(const pbnn_mlp_spec* spec, int op, int rows, int C);

/* Workspace.  Models whose parameter-sized state does not fit shared memory (e.g. 90-256-256-1, P = 89 345) keep it in
 * a caller-provided, 16-byte aligned DEVICE workspace (L2 resident while the chain is advanced).  Returns the bytes the
 * sampler `op` needs for C chains (0 when everything is resident on chip; pass NULL then), < 0 on error.
 * PBNN_OP_HMC on the tcgen05 engine with more chains than SMs also returns > 0: hand-over buffers of the balanced
 * persistent schedule (C * (P + 4) * 4 + 16 bytes).  That workspace is OPTIONAL -- without it pbnn_hmc_run runs one
 * CTA per chain (same results bit for bit, whole waves instead of balanced SMs). */
#define PBNN_OP_SGLD 0
#define PBNN_OP_PSGLD 1
#define PBNN_OP_SGHMC 2
#define PBNN_OP_ADAM 3
#define PBNN_OP_GRAD 4
#define PBNN_OP_HMC 5
int64_t pbnn_workspace_bytes(const pbnn_mlp_spec* spec, int op, int rows, int C);

/* ---- jax.random primitives (jax/_src/prng.py, random.py); test hooks + key plumbing --------------- */
/* jax.random.split: out[c][i] = threefry(keys[c], (0, i)).  out: [C][n][2] */
int pbnn_threefry_split(const uint32_t* keys, int C, int n, uint32_t* out, void* stream);
/* jax.random.fold_in(key, data) = threefry(key, (0, data)).  out: [C][2] */
int pbnn_threefry_fold_in(const uint32_t* keys, int C, uint32_t data, uint32_t* out, void* stream);
/* jax.random.bits(key, (n,), uint32).  out: [C][n] */
int pbnn_threefry_bits(const uint32_t* keys, int C, int n, uint32_t* out, void* stream);
/* jax.random.uniform / normal float32.  out: [C][n] */
int pbnn_threefry_uniform(const uint32_t* keys, int C, int n, float* out, void* stream);
int pbnn_threefry_normal(const uint32_t* keys, int C, int n, float* out, void* stream);
/* indices of the `draw`-th (1-based) `next(batch_labeled_data(...))` (src/pbnn/utils/data.py:47-78):
 * key <- split(key)[1] `draw` times, then randint(key,(B,),0,N).  out: [C][B] int32 */
int pbnn_minibatch_indices(const uint32_t* keys, int C, int draw, int B, int N, int32_t* out, void* stream);
/* jax.random.permutation(key, N) (stable sort by fresh random bits, ceil(3 ln N / ln(2^32-1)) rounds).
 * out: [C][N] int32 ; workspace: at least 3*C*N uint64 (device). */
int pbnn_permutation(const uint32_t* keys, int C, int N, int32_t* out, void* workspace, void* stream);

/* ---- samplers -------------------------------------------------------------------------------------- */
/* Common arguments of the SG samplers:
 *   X [N][d], y [N][s]           training data (row major)
 *   idx                          optional explicit minibatch indices [C][idx_steps][B] (idx_steps 1 or T);
 *                                NULL = generated on device from keys[c] according to `mb_mode`
 *   keys [C][2]                  the `rng_key` of each chain (split(rng_key,T)[t] is derived on device)
 *   theta [C][P]                 in: init_positions, out: final positions
 *   step_sizes [n_step_sizes]    1 value (constant) or T values (scheduled_sgld: langevin.py:143-204)
 *   t0                           index of the first step (keys[t0+t] are used) -- exact resume
 *   thin, burnin                 trace keeps steps t>=burnin with (t-burnin)%thin==0
 *   trace [C][T_out][P]          optional (NULL = keep only the final state)
 *   flags [C]                    optional; bit 0: the chain became non-finite; bit 1 (HMC): a segment of the balanced
 *                                schedule timed out and was skipped (an error: the host API raises); bit 2: a
 *                                caller-supplied minibatch index was outside [0, N) and has been clamped
 *   cv_grad_center, cv_grad_center_est [C][P]  (sgld, sghmc) optional control variates of blackjax.sgmcmc.gradients
 *                                control_variates, as used by sgld_cv / sghmc_cv / *_svrg (langevin.py:207-354,
 *                                hamiltonian.py:207-359): gradient of the full-data log-posterior at the centring
 *                                position (pbnn_logposterior_grad) and the minibatch estimate at the centring position
 *                                (pbnn_grad_estimator on the loop's fixed minibatch); the step uses
 *                                grad_center + grad_est - grad_center_est.  NULL, NULL = plain estimator.
 *   temperature                  (sgld, sghmc) 1 in pbnn; 0 turns the step into the plain gradient step of the SGD
 *                                phase of cyclical_sgld (langevin.py:364-470)
 */

/* pbnn.mcmc.sgld (src/pbnn/mcmc/langevin.py:21-78) over blackjax.sgld */
int pbnn_sgld_run(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const int32_t* idx,
                  int idx_steps, int mb_mode, int B, const uint32_t* keys, float* theta, const float* step_sizes,
                  int n_step_sizes, int64_t t0, int T, int C, int thin, int burnin, float* trace, uint32_t* flags,
                  const float* cv_grad_center, const float* cv_grad_center_est, float temperature, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* pbnn.mcmc.pSGLD (langevin.py:81-140; sgmcmc/psgld.py:25-48; sgmcmc/diffusions.py:33-60).
 * g,V [C][P]: logprob_grad / square_avg of pSGLDState; if has_state==0 they are initialised on device
 * from the init draw (psgld.init) and written back at the end (resume). */
int pbnn_psgld_run(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const int32_t* idx,
                   int idx_steps, int mb_mode, int B, const uint32_t* keys, float* theta, float* g, float* V,
                   int has_state, float alpha, const float* step_sizes, int n_step_sizes, int64_t t0, int T, int C,
                   int thin, int burnin, float* trace, uint32_t* flags, void* workspace, int64_t workspace_bytes,
                   void* stream);

/* pbnn.mcmc.kernels.psgld(grad_estimator_fn, alpha) -> SamplingAlgorithm(init_fn, step_fn)
 * (src/pbnn/mcmc/kernels.py:80-94; sgmcmc/psgld.py:25-48): BlackJAX-style single transitions on an EXPLICIT
 * minibatch Xb [B][d], yb [B][s] shared by the C positions; `data_size` is the N of the N*mean(loglik) estimator.
 * init: g = grad(theta, batch), V = g^2.  step: keys[c] is used directly as the transition's rng_key. */
int pbnn_psgld_init(const pbnn_mlp_spec* spec, const float* Xb, const float* yb, int B, int data_size,
                    const float* theta, float* g, float* V, int C, void* stream);
int pbnn_psgld_step(const pbnn_mlp_spec* spec, const float* Xb, const float* yb, int B, int data_size,
                    const uint32_t* keys, float* theta, float* g, float* V, float alpha, float step_size, int C,
                    void* stream);

/* pbnn.mcmc.sghmc (src/pbnn/mcmc/hamiltonian.py:80-140) over blackjax.sghmc(grad_fn, L, alpha=0.01, beta=0).
 * Momentum resampling at the start of every iteration (upstream: `generate_gaussian_noise(rng_key, position, ...)` in
 * blackjax/sgmcmc/sghmc.py, not vendored -- SURVEY.md 2b):
 *     p = (momentum_mu + momentum_mu_sqrt_step * sqrt(step)) + (momentum_sigma + momentum_sigma_sqrt_step * sqrt(step)) z
 * (0, 0, 1, 0): p ~ N(0, 1) ("unit", a call without a scale argument); (0, 0, 0, 1): p ~ N(0, step), the SGHMC paper's
 * parameterisation and the only one of the three that yields a usable sampler at the reference's settings;
 * (0, 1, 1, 0): p ~ N(sqrt(step), 1), sqrt(step) passed positionally into the `mu` slot. */
int pbnn_sghmc_run(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const int32_t* idx,
                   int idx_steps, int mb_mode, int B, const uint32_t* keys, float* theta, int L, float alpha,
                   float beta, float momentum_mu, float momentum_mu_sqrt_step, float momentum_sigma,
                   float momentum_sigma_sqrt_step, const float* step_sizes, int n_step_sizes, int64_t t0, int T, int C,
                   int thin, int burnin, float* trace, uint32_t* flags, const float* cv_grad_center,
                   const float* cv_grad_center_est, float temperature, void* workspace, int64_t workspace_bytes,
                   void* stream);

/* pbnn.mcmc.hmc (hamiltonian.py:18-77) over blackjax.hmc with the full-batch target
 * logprior + sum_i loglik_i (benchmark/pipelines/hmc.py:59-62).
 *   inv_mass [P] (flat order), logp [C] out (log density of the final state), accept_count [C] out,
 *   info_delta [C][num_samples] / info_accept [C][num_samples] optional per-iteration diagnostics. */
int pbnn_hmc_run(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const uint32_t* keys,
                 float* theta, const float* inv_mass, float step_size, int L, int64_t t0, int num_samples, int C,
                 int thin, int burnin, float* trace, float* logp, int32_t* accept_count, float* info_delta,
                 int32_t* info_accept, uint32_t* flags, void* workspace, int64_t workspace_bytes, void* stream);

/* value and gradient of the full-batch log-posterior logprior + sum_i loglik_i (benchmark/pipelines/hmc.py:59-62;
 * also the `logdensity_grad_estimator(centering_position, data)` of blackjax control_variates) for C positions.
 * grad [C][P] and/or logp [C]. */
int pbnn_logposterior_grad(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const float* theta, int C,
                           float* grad, float* logp, void* workspace, int64_t workspace_bytes, void* stream);

/* inner loop of pbnn.deep_ensembles.deep_ensembles_fn (src/pbnn/deep_ensembles.py:59-75): n_steps Adam
 * (optax.adam defaults) steps on -logposterior for C members at once.
 *   idx [C][n_steps][B] minibatch rows (permutation-derived, see pbnn_permutation)
 *   m, v [C][P] Adam moments (in/out), count0 = steps already taken */
int pbnn_adam_ensemble_run(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const int32_t* idx,
                           int n_steps, int B, float* theta, float* m, float* v, int count0, float lr, float b1,
                           float b2, float eps, int C, uint32_t* flags, void* workspace, int64_t workspace_bytes,
                           void* stream);

/* predict_fn (langevin.py:75-76) + moments (metrics_prediction_intervals.py:188; variance = ddof 0 extension).
 *   samples [S][P], Xtest [M][d];  preds [S][M][s] optional; sum / sumsq [M][s] optional (sum_s f, sum_s f^2);
 *   workspace: pbnn_predict_workspace_bytes(spec, S, M, want_moments) bytes, 16-byte aligned (partial sums of the
 *   sample chunks, plus padded parameter copies for models that do not fit shared memory); NULL when that is 0. */
int pbnn_predict_chunks(int S);
int64_t pbnn_predict_workspace_bytes(const pbnn_mlp_spec* spec, int S, int M, int want_moments);
int pbnn_predict(const pbnn_mlp_spec* spec, const float* samples, int S, const float* Xtest, int M, float* preds,
                 float* sum, float* sumsq, float* workspace, int64_t workspace_bytes, void* stream);

/* thinning_fn (src/pbnn/utils/misc.py:58-125): greedy energy-distance thinning of `n` flat positions [n][D] down to
 * `size` indices, in the reference's order (picks 1..size-1, then the initial pick).  workspace: 3*n + 2064 floats. */
int pbnn_thinning(const float* positions, int n, int D, int size, int32_t* idx_out, float* workspace, void* stream);

/* gradient of the (scaled) log-posterior estimator for C positions on explicit minibatches -- test hook for
 * blackjax grad_estimator == src/pbnn/utils/misc.py:34-55.  idx [C][B]; grad [C][P]; loglik [C] optional. */
int pbnn_grad_estimator(const pbnn_mlp_spec* spec, const float* X, const float* y, int N, const int32_t* idx, int B,
                        const float* theta, int C, float* grad, float* loglik, void* workspace, int64_t workspace_bytes,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PBNN_B200_H */
