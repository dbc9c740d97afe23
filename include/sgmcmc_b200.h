/* libsgmcmc_b200.so - C ABI of the B200-native SGMCMC hot path.
 *
 * The reference (jeremiecoullon/SGMCMCJax v0.2.13) is pure Python on JAX and has
 * no FFI layer; the interface this library replaces is its closure API:
 *
 *   sgmcmcjax/samplers.py:27-60,98-136   build_*_sampler(...) -> sampler(key, Nsamples, params)
 *   sgmcmcjax/kernels.py:17-124,130-601  build_*_kernel(...)  -> (init_fn, kernel, get_params)
 *   sgmcmcjax/gradient_estimation.py:12-153  the three minibatch gradient estimators
 *   sgmcmcjax/diffusions.py:47-347       the seven diffusion updates
 *   sgmcmcjax/util.py:9-50               build_grad_log_post
 *   sgmcmcjax/ksd.py:43-66               imq_KSD
 *
 * Conventions
 *  - Every pointer documented as "device" is a CUDA device pointer on the
 *    current device, BORROWED for the duration of the call (the model handle
 *    additionally borrows nothing: sgmcmc_set_data copies into its own layout).
 *  - `stream` is a cudaStream_t passed as void*; every call is stream-ordered
 *    and returns without synchronising unless stated.
 *  - Return value: 0 = OK, negative = error (see SGMCMC_E_*); the message of
 *    the last error on the calling thread is returned by sgmcmc_last_error().
 *  - A handle is not thread-safe; distinct handles are independent.
 *  - All floating point is IEEE fp32, indices int32, PRNG words uint32
 *    (jax 0.4.13 threefry2x32, non-partitionable layout).
 */
#ifndef SGMCMC_B200_H
#define SGMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGMCMC_ABI_VERSION 4

enum { SGMCMC_OK = 0, SGMCMC_E_INVALID = -1, SGMCMC_E_UNSUPPORTED = -2, SGMCMC_E_CUDA = -3,
       SGMCMC_E_STATE = -4 };

/* Registered model families (BASELINE.json north_star). */
enum { SGMCMC_FAMILY_GAUSSIAN_MEAN = 0,   /* README.md:22-26, tests/models.py:6-24            */
       SGMCMC_FAMILY_LOGREG = 1,          /* models/logistic_regression.py:65-72               */
       SGMCMC_FAMILY_MLP = 2,             /* models/bayesian_NN/NN_model.py:31-58              */
       SGMCMC_FAMILY_PMF = 3 };           /* SURVEY.md a15 (not in the reference)              */

/* gradient_estimation.py:12-46 / :50-94 / :97-153 */
enum { SGMCMC_EST_STANDARD = 0, SGMCMC_EST_CV = 1, SGMCMC_EST_SVRG = 2 };

/* diffusions.py:47-347 */
enum { SGMCMC_DIFF_SGLD = 0, SGMCMC_DIFF_PSGLD = 1, SGMCMC_DIFF_SGLDADAM = 2, SGMCMC_DIFF_SGHMC = 3,
       SGMCMC_DIFF_BAOAB = 4, SGMCMC_DIFF_SGNHT = 5, SGMCMC_DIFF_BADODAB = 6,
       SGMCMC_DIFF_ADAM_OPT = 7 };   /* not a sampler: the Adam MAP optimiser of optimizer.py:12-62 (sgmcmc_optimize only);
                                        hyper = {b1, b2, eps}, step coefficients as for sgldAdam with h = learning rate */

#define SGMCMC_MAX_LEAVES 8

typedef struct sgmcmc_config {
  int32_t abi_version;        /* SGMCMC_ABI_VERSION */
  int32_t family;             /* SGMCMC_FAMILY_*    */
  int32_t estimator;          /* SGMCMC_EST_*       */
  int32_t diffusion;          /* SGMCMC_DIFF_*      */
  int64_t n_data;             /* rows of data[0] (util.py:41)                                   */
  int32_t batch;              /* minibatch size; == n_data selects the static full-batch bypass
                                 of the standard estimator (gradient_estimation.py:41-42)      */
  int32_t n_leaves;           /* parameter pytree leaves, jax leaf order                        */
  int32_t leaf_size[SGMCMC_MAX_LEAVES];   /* elements per leaf; P = sum                         */
  int32_t data_dim;           /* columns of data[0]                                             */
  int32_t leapfrog;           /* L of build_sghmc*_kernel (kernels.py:471-601); 0 otherwise     */
  int32_t update_rate;        /* SVRG update_rate (gradient_estimation.py:134)                  */
  int32_t flags;              /* SGMCMC_FLAG_*                                                  */
  float lik_coef[SGMCMC_MAX_LEAVES];      /* gaussian_mean: a_l in  -a_l |x - theta_l|^2        */
  float prior_coef[SGMCMC_MAX_LEAVES];    /* gaussian_mean: c_l in  -c_l |theta_l|^2 ;
                                             logreg: prior_coef[0] = prior precision (0.1)      */
  float hyper[4];             /* psgld {alpha, eps}; sgldAdam {beta1, beta2, eps};
                                 sghmc {alpha}; sgnht {a}; badodab {a}                          */
  int32_t dims[8];            /* pmf: {n_users, n_items, rank}, leaves (U[n_users, rank], V[n_items, rank]),
                                 lik_coef[0] = alpha (rating precision), prior_coef[0] = lambda;
                                 mlp: {n_layers, size_0, ..., size_L}, leaves (W_1[size_1, size_0], b_1, ...),
                                 prior_coef[0] = prior precision (1 for NN_model.py:53-58)          */
} sgmcmc_config;

/* kernels.py:457-460: build_badodabCV_kernel binds both palindrome halves to badodab's
 * update2.  With this flag the literal (buggy) wiring is reproduced. */
#define SGMCMC_FLAG_BADODABCV_REFERENCE_BUG 1

/* Per-chain sampler state, caller-owned device buffers, C = n_chains rows each
 * (types.py:14-34: DiffusionState + param_grad + SVRGState; `key` is the carry of
 * the lax.scan in samplers.py:47-57).  Unused members may be NULL. */
typedef struct sgmcmc_state {
  int32_t n_chains;
  int32_t reserved;
  float* theta;        /* [C, P]  x of every leaf, flattened in leaf order              */
  float* aux1;         /* [C, P]  v (psgld/sghmc/baoab/sgnht/badodab) or m (sgldAdam)   */
  float* aux2;         /* [C, P]  v of sgldAdam                                         */
  float* alpha;        /* [C, n_leaves] thermostat scalar per leaf (sgnht/badodab)      */
  float* grad;         /* [C, P]  carried param_grad (kernels.py:54,58)                 */
  float* svrg_center;  /* [C, P]  SVRGState.centering_value                             */
  float* svrg_grad;    /* [C, P]  SVRGState.fb_grad_center                              */
  uint32_t* key;       /* [C, 2]  scan-carry key                                        */
} sgmcmc_state;

/* One entry per step i (or a single entry with stride 0 for a constant dt):
 * h = fp32(dt(i)), hh = fp32(dt(i)/2), ca / cb = diffusion-specific coefficients
 * evaluated by the host with the reference's weak-type rounding:
 *   sgld     ca = sqrt(fp32(2 dt))
 *   sgldAdam ca = 1 - beta1^(i+1), cb = 1 - beta2^(i+1)
 *   sghmc    ca = sqrt(fp32(2 (alpha - beta) dt)), cb = sqrt(fp32(dt))      (resample)
 *   baoab    ca = exp(fp32(-gamma dt)),  cb = fp32(tau) * sqrt(1 - ca^2)
 *   sgnht    ca = sqrt(fp32(2 a dt)),    cb = sqrt(fp32(dt(0)))
 *   badodab  ca = sqrt(fp32(dt))                                            (alpha == 0 branch) */
typedef struct sgmcmc_step_coef { float h, hh, ca, cb; } sgmcmc_step_coef;

typedef struct sgmcmc_handle sgmcmc_handle;

const char* sgmcmc_last_error(void);
int sgmcmc_abi_version(void);

/* Number of CUDA kernels this library has launched in the calling process so far
 * (monotone; bench.py reports the difference over its timed region as gpu_launches). */
int64_t sgmcmc_launch_count(void);

int sgmcmc_create(const sgmcmc_config* cfg, sgmcmc_handle** out);
int sgmcmc_destroy(sgmcmc_handle* h);

/* Launch shape the vector-family sampler uses for n_chains chains (gaussian_mean, logreg): warps per CTA and CTAs per
 * chain.  cluster > 1 means one thread-block cluster per chain: the minibatch of gradient_estimation.py:32-33 is split
 * over the CTAs and the partial sums are exchanged over distributed shared memory (few chains per GPU: config 2 on
 * 8 GPUs, the single chain of samplers.py:45-58).  Environment overrides: SGMCMC_WARPS_PER_CHAIN, SGMCMC_CLUSTER. */
int sgmcmc_launch_shape(const sgmcmc_handle* h, int32_t n_chains, int32_t* warps_per_cta, int32_t* cluster);

/* Ingest the data tuple (gradient_estimation.py:25-29).  x: device [n_data, data_dim] row-major,
 * f32 (gaussian_mean, logreg, mlp) or int32 (pmf: (user, item) pairs, data_dim = 2).
 * y: device, NULL for 1-tuples; logreg: int32 [n_data] in {0,1}; mlp: f32 [n_data, size_L];
 * pmf: f32 [n_data] ratings.  The rows are re-laid-out into the handle's own HBM table
 * (see DESIGN.md "data layout"). */
int sgmcmc_set_data(sgmcmc_handle* h, const void* x, const void* y, void* stream);

/* CV centring value (gradient_estimation.py:50-72): theta_hat device f32 [P]; the full-batch
 * gradient at theta_hat is computed here (gradient_estimation.py:70). */
int sgmcmc_set_centering(sgmcmc_handle* h, const float* theta_hat, void* stream);

/* Full-batch grad_log_post (util.py:43-49 with all rows) at M parameter vectors:
 * thetas device f32 [M, P] -> out device f32 [M, P].  This is also the `grads` producer of imq_KSD (ksd.py:43).
 * Logistic regression with m >= 64 runs as two tcgen05 contractions (operand copies of the table are built on the
 * first such call and owned by the handle); otherwise one streaming pass over the rows per parameter vector. */
int sgmcmc_fullbatch_grad(sgmcmc_handle* h, const float* thetas, int32_t m, float* out, void* stream);

/* init_fn(key, params) of kernels.py:48-51 for every chain: state->theta must hold params.
 * keys: device u32 [C, 2].  sampler_mode != 0 additionally performs `key, subkey =
 * split(key)` (samplers.py:53), uses subkey for init_gradient and stores the carry in
 * state->key; sampler_mode == 0 uses keys[c] directly and leaves state->key alone. */
int sgmcmc_init(sgmcmc_handle* h, sgmcmc_state* st, const uint32_t* keys, int32_t sampler_mode,
                void* stream);

/* K iterations of the scan body of samplers.py:47-51 for every chain, i = i0 .. i0+K-1.
 * coef: device array of K entries (coef_stride 1) or one entry (coef_stride 0).
 * samples: NULL, or device f32 with element (c, s, e) at samples[c * sample_chain_stride + s * P + e]
 * (get_params after every iteration); sample_chain_stride == 0 means the dense [C, K, P]. */
int sgmcmc_run(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k,
               const sgmcmc_step_coef* coef, int32_t coef_stride, float* samples,
               int64_t sample_chain_stride, void* stream);

/* sgmcmc_run with thinning and on-device posterior summaries (the reference keeps every state,
 * samplers.py:57).  Iteration s = 0 .. k-1 of the launch is KEPT when (s + 1) % thin == 0 (thin >= 1; keep launches
 * aligned: i0 % thin == 0 and k % thin == 0 make chunked runs equal one long run).  A kept state goes to sample slot
 * (s + 1) / thin - 1 of `samples` (NULL: nothing is stored; sample_chain_stride == 0 means the dense [C, k / thin, P])
 * and, when `moments` is not NULL, is accumulated around a caller-chosen reference point r (device f32 [C, 3, P]:
 * plane 2 holds r, e.g. the initial parameters; planes 0 and 1 are zeroed by the caller before the first launch):
 * moments[c][0][e] += theta_e - r_e, moments[c][1][e] += (theta_e - r_e)^2.  Mean = r + S0 / n and marginal variance
 * = S1 / n - (S0 / n)^2 follow on the host; the shift keeps the fp32 sums free of cancellation. */
int sgmcmc_run_thinned(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k,
                       const sgmcmc_step_coef* coef, int32_t coef_stride, int32_t thin, float* samples,
                       int64_t sample_chain_stride, float* moments, void* stream);

/* K iterations of the optimiser loop of optimizer.py:41-60 (vector families; the handle is created with
 * SGMCMC_DIFF_ADAM_OPT and the standard estimator): key, subkey = split(key); value and gradient of log_post on the
 * minibatch drawn with subkey; optax.adam ascent step.  state: theta (params), aux1 / aux2 (Adam moments, zero at the
 * start), grad (scratch), key.  lp_out: NULL or device f32 [C, K], the log-posterior values (optimizer.py:47). */
int sgmcmc_optimize(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k, const sgmcmc_step_coef* coef, float* lp_out,
                    void* stream);

/* kernel(i, key, state) of kernels.py:53-64 / :110-122 with caller-supplied per-chain keys
 * (device u32 [C, 2]); state->key is not touched. */
int sgmcmc_kernel_step(sgmcmc_handle* h, sgmcmc_state* st, int32_t i, const uint32_t* keys,
                       const sgmcmc_step_coef* coef, void* stream);

/* ---- optional data-sharded mode (vector families, standard / CV estimators, no SGHMC) ----------
 * The ROWS of the table are split over ranks (for tables that do not fit one GPU); every rank runs every chain with the
 * same keys, draws the same minibatch indices (gradient_estimation.py:32 over ALL n_data rows), processes the ones it
 * holds, and the host all-reduces the [C, DS] minibatch sums once per step (DS = sgmcmc_sum_width(h): the row stride of
 * the table in floats - data_dim rounded up to 4, plus 4 for logreg with the CV estimator, whose rows carry center . x~).
 *   sgmcmc_create(cfg with n_data = TOTAL rows); sgmcmc_set_row_shard(h, row0, n_rows); sgmcmc_set_data(h, x, y) with
 *   n_rows rows; [CV: all-reduce sgmcmc_fullbatch_grad, then sgmcmc_set_centering_grad];
 *   sgmcmc_shard_init -> all-reduce sum -> { sgmcmc_shard_step(flags = STEP [| UPDATE2]) -> all-reduce sum } x K
 *   -> sgmcmc_shard_step(flags = [UPDATE2]) to finish the last gradient. */
#define SGMCMC_SHARD_UPDATE2 1   /* run the palindrome's update2 (baoab / badodab, coef_prev) after finishing the gradient */
#define SGMCMC_SHARD_STEP 2      /* then do one transition (samplers.py:47-51) up to this rank's partial sums */
int sgmcmc_set_row_shard(sgmcmc_handle* h, int64_t row0, int64_t n_rows);
/* DS: floats per row of the device table == width of the sum_in / sum_out blocks below. */
int sgmcmc_sum_width(const sgmcmc_handle* h);
/* CV centre with an externally reduced full-batch gradient (gradient_estimation.py:70): device f32 [P] each. */
int sgmcmc_set_centering_grad(sgmcmc_handle* h, const float* theta_hat, const float* grad, void* stream);
/* init_fn (kernels.py:48-51) up to this rank's partial minibatch sums: sum_out device f32 [C, DS]. */
int sgmcmc_shard_init(sgmcmc_handle* h, sgmcmc_state* st, const uint32_t* keys, int32_t sampler_mode, float* sum_out,
                      void* stream);
/* Finish the gradient of the previous estimate from the all-reduced sums sum_in [C, DS]; then, per `flags`, update2 and
 * one transition i whose sample goes to sample_out [C, P] (may be NULL) and whose partial sums go to sum_out. */
int sgmcmc_shard_step(sgmcmc_handle* h, sgmcmc_state* st, int32_t i, const sgmcmc_step_coef* coef,
                      const sgmcmc_step_coef* coef_prev, int32_t flags, const float* sum_in, float* sum_out, float* sample_out,
                      void* stream);

/* random.choice(key, arange(n), (batch,)) (gradient_estimation.py:32): key is a HOST
 * pointer to 2 words, out a device int32 [batch].  Test hook for bit-exactness. */
int sgmcmc_minibatch_indices(const uint32_t* key, int64_t n, int32_t batch, int32_t* out, void* stream);

/* random.normal(key, (n,)) in fp32 (diffusions.py:67 etc.): key HOST pointer, out device f32 [n]. */
int sgmcmc_normal(const uint32_t* key, int32_t n, float* out, void* stream);

/* random.uniform(key, (n,)) on [0, 1) in fp32 (jax 0.4.13: 23 mantissa bits per word) -> device out[n]. */
int sgmcmc_uniform(const uint32_t* key, int32_t n, float* out, void* stream);
/* vmap(random.bernoulli)(random.split(key, n), p) as models/logistic_regression.py:59-61 draws the labels:
 * out[i] = uniform(split(key, n)[i], ()) < p[i]; p device f32 [n], out device int32 [n]. */
int sgmcmc_bernoulli_rows(const uint32_t* key, int32_t n, const float* p, int32_t* out, void* stream);
/* random.split(key, num): key HOST pointer, out device u32 [num, 2]. */
int sgmcmc_split(const uint32_t* key, int32_t num, uint32_t* out, void* stream);

/* Partial sum of the IMQ Stein kernel k_0 (ksd.py:7-37, c = 1, beta = -1/2) over the ordered
 * pairs (i, j) with row0 <= i < row1 and 0 <= j < n.  x, g: device f32 [n, d] (samples,
 * grads).  partial_sum: device double[1], overwritten.  imq_KSD = sqrt(sum over all row
 * blocks) / n (ksd.py:65-66).  scratch may be NULL (allocated internally). */
int sgmcmc_ksd_imq_partial(const float* x, const float* g, int32_t n, int32_t d, int32_t row0,
                           int32_t row1, double* partial_sum, void* stream);

/* Tensor-core (tcgen05, 3xTF32) variant of the same sum, sharded by Gram TILES instead of rows: the
 * upper-triangular 128 x 128 tile list (k_0 is symmetric; off-diagonal tiles count twice) is dealt
 * round-robin to `nparts` partitions and this call sums partition `part`.  The sum over all partitions
 * is the sum over all n^2 ordered pairs.  workspace: device scratch of sgmcmc_ksd_workspace_bytes(n, d)
 * bytes (centred / tf32-split operands, row terms), borrowed for the call. */
int64_t sgmcmc_ksd_workspace_bytes(int32_t n, int32_t d);
int sgmcmc_ksd_imq_tc(const float* x, const float* g, int32_t n, int32_t d, int32_t part, int32_t nparts,
                      void* workspace, double* partial_sum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGMCMC_B200_H */
