/*
 * pzflow_b200 -- C-ABI of the B200-native evaluator for pzflow's
 * neural-spline-flow hot path (Flow.log_prob / Flow.sample / Flow.posterior).
 *
 * pzflow itself has no FFI: its extension seam is the Bijector protocol
 * (pzflow/bijectors.py:16-118) -- InitFunction(rng, input_dim) ->
 * (params, ForwardFunction, InverseFunction), with
 * Forward/InverseFunction(params, inputs[N,D], conditions=[N,C]) ->
 * (outputs[N,D], log_det[N]) -- which Flow binds as _forward/_inverse
 * (pzflow/flow.py:285-288, 185-188) and funnels through
 * Flow._log_prob(params, inputs, conditions) (pzflow/flow.py:397-407).
 * The entry points below are what a ctypes / XLA-FFI binding for that seam
 * would bind; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer to C-contiguous row-major float32
 *    (rows = samples, exactly the reference's [N, D] layout) unless the
 *    parameter is documented as host memory;
 *  - the caller owns every buffer; a plan owns only its packed weight copy;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - every function returns 0 on success and a negative pzb_status otherwise,
 *    nothing throws across the ABI; pzb_last_error() describes the last
 *    failure on the calling thread;
 *  - a plan is immutable after creation and may be launched concurrently from
 *    several host threads / streams; one plan per device.
 */
#ifndef PZFLOW_B200_H
#define PZFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZB_ABI_VERSION 1

typedef enum {
  PZB_OK = 0,
  PZB_ERR_INVALID = -1,      /* bad argument (maps to ValueError in the Python mirror)  */
  PZB_ERR_UNSUPPORTED = -2,  /* stage / shape outside the hot path                       */
  PZB_ERR_CUDA = -3,         /* CUDA runtime error, see pzb_last_error()                 */
  PZB_ERR_NO_DEVICE = -4     /* no CUDA device: there is no CPU fallback                 */
} pzb_status;

/* Bijector stages of the hot path (a flattened Bijector_Info, nested Chains
 * inlined: pzflow/bijectors.py:12-13, pzflow/utils.py:11-19). */
typedef enum {
  PZB_STAGE_SHIFT_BOUNDS = 1,     /* pzflow/bijectors.py:717-777 */
  PZB_STAGE_STANDARD_SCALER = 2,  /* pzflow/bijectors.py:817-861 */
  PZB_STAGE_ROLL = 3,             /* pzflow/bijectors.py:559-598 */
  PZB_STAGE_NSC = 4,              /* pzflow/bijectors.py:375-522 (NeuralSplineCoupling) */
  /* Chain nesting markers (pzflow/bijectors.py:155-160): every Chain sums its
   * stages' log-determinants from zeros before its parent adds the total, so
   * the flattening keeps the brackets to reproduce the fp32 summation order. */
  PZB_STAGE_CHAIN_BEGIN = 5,
  PZB_STAGE_CHAIN_END = 6,
  /* parameter-free bijectors, evaluated as steps of the same fused chain */
  PZB_STAGE_REVERSE = 7,          /* pzflow/bijectors.py:525-556: columns reversed (folded like Roll)          */
  PZB_STAGE_SCALE = 8,            /* pzflow/bijectors.py:668-714: B = scale                                     */
  PZB_STAGE_INV_SOFTPLUS = 9,     /* pzflow/bijectors.py:314-372: vec_a = column indices, vec_b = sharpness(es) */
  PZB_STAGE_COLOR_TRANSFORM = 10, /* pzflow/bijectors.py:177-311: shift = ref_idx, vec_a = mag_idx, n_vec       */
  PZB_STAGE_SHUFFLE = 11,         /* pzflow/bijectors.py:780-814: out[:, j] = in[:, perm[j]], vec_a = perm (as floats),
                                     n_vec = D; the caller derives perm from the jax key schedule (folded like Roll) */
  PZB_STAGE_DEQUANTIZE = 12       /* pzflow/bijectors.py:865-911 (UniformDequantizer) AS THE REFERENCE EXECUTES IT: the
                                     forward map is the identity (its .at[].set() result is discarded, line 897), the
                                     inverse floors the columns vec_a[0, n_vec); log-det 0 both ways */
} pzb_stage_kind;

/* Latent distributions of the hot path. */
typedef enum {
  PZB_LATENT_UNIFORM = 1,    /* pzflow/distributions.py:395-470 */
  PZB_LATENT_CENTBETA13 = 2, /* pzflow/distributions.py:137-229 */
  PZB_LATENT_NORMAL = 3,     /* pzflow/distributions.py:232-303 (B unused) */
  PZB_LATENT_CENTBETA = 4,   /* pzflow/distributions.py:45-135: alpha = beta = 1 until pzb_plan_set_latent     */
  PZB_LATENT_TDIST = 5,      /* pzflow/distributions.py:306-392: nu = 30 until pzb_plan_set_latent (B unused)  */
  PZB_LATENT_JOINT = 6       /* pzflow/distributions.py:473-562: segments via pzb_plan_set_latent (required)   */
} pzb_latent_kind;

/* One latent distribution over `dim` consecutive columns.  params: HOST floats in the reference's pytree order --
 * CentBeta ((log a_1, log b_1), ...) flattened to 2 * dim values, Tdist log(nu) -- or NULL for the defaults. */
typedef struct {
  int32_t kind; /* pzb_latent_kind, not JOINT */
  int32_t dim;
  float B;
  const float *params;
  int32_t n_params;
} pzb_latent_desc;

/* Conditioner evaluation engines. */
typedef enum {
  PZB_ENGINE_AUTO = 0,  /* tensor-core engine when the plan's shapes allow it, else SIMT */
  PZB_ENGINE_SIMT = 1,  /* fp32 FFMA conditioner (any shape)                              */
  PZB_ENGINE_TC = 2,    /* tcgen05 conditioner, fp16x2 split-accumulate in fp32 TMEM      */
  PZB_ENGINE_WIDE = 3   /* tcgen05 conditioner for wide / deep shapes (hidden <= 256, K <= 32, any number of hidden
                           layers >= 1, <= 63 conditioner inputs): weights streamed L2 -> shared memory in slabs */
} pzb_engine;

typedef struct {
  int32_t kind; /* pzb_stage_kind */

  /* ROLL: jnp.roll(x, shift, axis=-1). */
  int32_t shift;

  /* SHIFT_BOUNDS: vec_a = min, vec_b = max, n_vec = 1 or D, B = output half-range.
   * STANDARD_SCALER: vec_a = means, vec_b = stds, n_vec = D.  HOST pointers.
   * INV_SOFTPLUS: vec_a = column indices (as floats), vec_b = sharpness per index (or one value), n_vec = #columns.
   * COLOR_TRANSFORM: vec_a = mag_idx (as floats), n_vec = #magnitudes, shift = ref_idx. */
  const float *vec_a;
  const float *vec_b;
  int32_t n_vec;

  /* NSC: the Bijector_Info argument tuple
   * (K, B, hidden_layers, hidden_dim, transformed_dim, n_conditions, periodic);
   * transformed_dim < 0 means "None" (lower = D - D/2). B is shared with SHIFT_BOUNDS. */
  float B;
  int32_t K;
  int32_t hidden_layers;
  int32_t hidden_dim;
  int32_t transformed_dim;
  int32_t n_conditions;
  int32_t periodic;
  /* NSC parameters in stax order: 2*(hidden_layers+1) HOST pointers
   * W0, b0, W1, b1, ... with W[in, out] row-major float32
   * (pzflow/utils.py:45-48; stax Dense = x @ W + b). */
  const float *const *weights;
} pzb_stage_desc;

typedef struct pzb_plan pzb_plan;

/* ABI version of the loaded library (== PZB_ABI_VERSION). */
int pzb_abi_version(void);

/* Last error message of the calling thread (never NULL). */
const char *pzb_last_error(void);

/* Number of visible CUDA devices (0 when there is none; never an error). */
int pzb_device_count(void);

/* Compile a flattened bijector chain + latent into a device plan.
 * Replaces: Flow.set_bijector / the load path's build_bijector_from_info +
 * init_fun (pzflow/flow.py:183-189, 285-294) as the owner of the evaluation
 * functions.  Weights are copied and packed once; rebuild the plan when
 * flow._params changes identity.  `device` is the CUDA ordinal. */
int pzb_plan_create(const pzb_stage_desc *stages, int32_t n_stages, int32_t input_dim,
                    int32_t latent_kind, float latent_B, int32_t device, pzb_plan **out);

/* Latent distributions with parameters, and Joint: the latent as a list of segments over consecutive columns
 * (one segment for CentBeta / Tdist with learned parameters, several for Joint; the dims must add up to
 * input_dim).  Part of plan construction: call right after pzb_plan_create, before the first evaluation. */
int pzb_plan_set_latent(pzb_plan *plan, const pzb_latent_desc *segs, int32_t n_segs);

void pzb_plan_destroy(pzb_plan *plan);

/* Introspection used by the host mirror, the tests and bench.py. */
int pzb_plan_input_dim(const pzb_plan *plan);
int pzb_plan_n_conditions(const pzb_plan *plan);
/* total number of spline-transformed columns over all NSC stages
 * (the row width of the optional `bins` debug output). */
int pzb_plan_n_spline_dims(const pzb_plan *plan);
/* conditioner MLP FLOPs per row (SURVEY.md §8d formula), for rooflines. */
double pzb_plan_flops_per_row(const pzb_plan *plan);
/* 1 if the tensor-core engine supports this plan's shapes: every coupling stage has 2 hidden layers of at most 128
 * units, at most 16 bins (fewer than 16 only when not periodic), at most 15 conditioner inputs (columns + conditions)
 * and at most 16 transformed columns; at most 31 data columns.  Everything else runs on the SIMT engine. */
int pzb_plan_tc_supported(const pzb_plan *plan);
/* 1 if the wide tensor-core engine (PZB_ENGINE_WIDE) supports this plan's shapes: every coupling stage has at least one
 * hidden layer of at most 256 units, at most 32 bins (periodic: exactly 16 or 32) and at most 63 conditioner inputs.
 * Replaces the same reference lines as the other engines (pzflow/bijectors.py:462-520, pzflow/utils.py:22-49, 64-227). */
int pzb_plan_wide_supported(const pzb_plan *plan);
/* Rows of FINITE inputs (cumulative over this plan's launches on a tensor-core engine) whose spline logits came out non-finite: the
 * conditioner's activations ride fp16 hi parts, so a pre-activation beyond +-65504 overflows where the reference's
 * fp32 (pzflow/utils.py:45-48) stays finite; such rows land in the zero-probability class.  Synchronises the device;
 * `reset` != 0 clears the counter.  The host-buffer entry points of an AUTO-engine plan re-evaluate a call on the
 * SIMT engine when it raised this counter. */
int pzb_plan_overflow_rows(const pzb_plan *plan, int32_t reset, int64_t *count);
/* select the conditioner engine for subsequent launches (default AUTO). */
int pzb_plan_set_engine(pzb_plan *plan, int32_t engine);
/* engine a launch of this plan uses right now (PZB_ENGINE_SIMT, _TC or _WIDE; AUTO prefers TC, then WIDE). */
int pzb_plan_engine(const pzb_plan *plan);
/* how many kernels of this library were launched by this process so far. */
int64_t pzb_launch_count(void);

/* Flow._log_prob (pzflow/flow.py:397-407): u, ld = forward(x, cond);
 * out = latent.log_prob(u) + ld; NaN -> zero probability.
 * X [N, D]; cond [N, C] already standard-scaled (pzflow/flow.py:334-336) or
 * NULL when C == 0; out [N].  bins (optional, may be NULL) receives the
 * spline bin index of every transformed column, int32 [N, n_spline_dims],
 * NSC stages in chain order. */
int pzb_log_prob(const pzb_plan *plan, const float *X, const float *cond, int64_t N,
                 float *out, int32_t *bins, void *stream);

/* Bijector ForwardFunction of the whole chain (pzflow/bijectors.py:162-164):
 * U [N, D], log_det [N]. */
int pzb_forward(const pzb_plan *plan, const float *X, const float *cond, int64_t N,
                float *U, float *log_det, int32_t *bins, void *stream);

/* Bijector InverseFunction of the whole chain (pzflow/bijectors.py:166-170),
 * the map Flow.sample applies to latent draws (pzflow/flow.py:795):
 * X [N, D], log_det [N] (may be NULL: Flow.sample discards it). */
int pzb_inverse(const pzb_plan *plan, const float *U, const float *cond, int64_t N,
                float *X, float *log_det, int32_t *bins, void *stream);

/* Flow.posterior main loop (pzflow/flow.py:666-738, err_samples=None,
 * marg_rules=None): rest [Ng, D-1] = data columns with `col_idx` removed,
 * cond [Ng, C] or NULL, grid [G]; pdfs [Ng, G] = exp(log_prob) on the
 * galaxy-major / grid-minor expansion, never materialised; then, per galaxy,
 * pdfs /= trapezoid(pdfs, grid) if normalize, NaN -> 0 if nan_to_zero. */
int pzb_posterior(const pzb_plan *plan, const float *rest, const float *cond, int64_t Ng,
                  int32_t col_idx, const float *grid, int32_t G, int32_t normalize,
                  int32_t nan_to_zero, float *pdfs, void *stream);

/* FlowEnsemble.log_prob reduction (pzflow/flowEnsemble.py:236-243):
 * lp [n_flows, N] (one pzb_log_prob result per row) -> out [N] =
 * log(mean_f(exp(lp[f, i]))), the naive form the reference uses. */
int pzb_ensemble_logmeanexp(const float *lp, int32_t n_flows, int64_t N, float *out,
                            void *stream);

/* FlowEnsemble.posterior reduction (pzflow/flowEnsemble.py:346-351):
 * pdfs [n_flows, Ng, G] un-normalised -> out [Ng, G] = mean over flows, then
 * trapezoid-normalised per galaxy if normalize.  (NaNs were already zeroed
 * per flow, as in the reference.) */
int pzb_ensemble_posterior_mean(const float *pdfs, int32_t n_flows, int64_t Ng, int32_t G,
                                const float *grid, int32_t normalize, float *out,
                                void *stream);

/* Per-galaxy trapezoid normalisation + NaN->0 of pdfs [Ng, G] in place
 * (pzflow/flow.py:732-737); also used for returnEnsemble=True
 * (pzflow/flowEnsemble.py:335-344). */
int pzb_normalize_pdfs(float *pdfs, int64_t Ng, int32_t G, const float *grid,
                       int32_t normalize, int32_t nan_to_zero, void *stream);

/* Latent draws for Flow.sample speed runs (pzflow/distributions.py:443-470):
 * U [N, D] ~ Uniform(-B, B) from a counter-based Philox stream.  This is NOT
 * jax.random's threefry stream (pzb_sample_uniform_threefry below is); kept for draws beyond one threefry call.
 * row_offset (here and in the three samplers below): index of U's first row in the caller's whole array, so that a
 * call sharded over rows (several GPUs, several chunks) draws exactly what one big call would; row_offset * D must be
 * a multiple of 4 for the uniform and the normal sampler (their draws come in blocks of four). */
int pzb_sample_uniform(float *U, int64_t N, int32_t D, float B, uint64_t seed, int64_t row_offset, void *stream);

/* Uniform.sample (pzflow/distributions.py:464-469) in jax's OWN random stream: rows [row_offset, row_offset + N) of
 * jax.random.uniform(key, shape=(total_rows, D), minval, maxval), key = (key0, key1) -- PRNGKey(seed) with x64
 * disabled is (0, seed mod 2^32).  Threefry-2x32/20 in the layout of jax's default PRNG implementation: the count
 * vector iota(total_rows * D) rides through the block function as two halves (partitionable = 0: jax < 0.5's default;
 * reproduces the samples printed from real JAX at docs/tutorials/marginalization.ipynb:133-135), or one block per
 * element with its two words xor-ed (partitionable = 1: jax_threefry_partitionable=True, the default of the jax >= 0.5.3
 * the reference requires; reproduces random.uniform(key(0)) = 0.947667 of jax's documentation).  total_rows
 * is needed because the original layout pairs element g with g + ceil(total / 2).  A request of 2^32 - 1 words or more
 * in the original layout returns PZB_ERR_UNSUPPORTED (jax cuts those into separately keyed blocks). */
int pzb_sample_uniform_threefry(float *U, int64_t N, int32_t D, float minval, float maxval, uint32_t key0, uint32_t key1,
                                int64_t total_rows, int64_t row_offset, int32_t partitionable, void *stream);

/* The other latents' draws (pzflow/distributions.py:101-135 CentBeta, 196-229 CentBeta13, 277-303 Normal, 360-392
 * Tdist), each from its own counter-based Philox stream keyed by (seed, element index):
 *   centbeta: U[N, D] = 2B (Beta(a[j], b[j]) - 0.5), Beta as a ratio of two Marsaglia-Tsang gammas; a, b are HOST
 *             arrays of D positive shapes (CentBeta13: all 13);
 *   normal:   U[N, D] standard normal (Box-Muller);
 *   tdist:    U[N, D] = z / sqrt(chi2(nu) / nu), one chi2 per row (unit scale matrix). */
int pzb_sample_centbeta(float *U, int64_t N, int32_t D, const float *a, const float *b, float B, uint64_t seed,
                        int64_t row_offset, void *stream);
int pzb_sample_normal(float *U, int64_t N, int32_t D, uint64_t seed, int64_t row_offset, void *stream);
int pzb_sample_tdist(float *U, int64_t N, int32_t D, float nu, uint64_t seed, int64_t row_offset, void *stream);

/* RationalQuadraticSpline as the stand-alone function the reference exports (pzflow/utils.py:64-227; the
 * `_RationalQuadraticSpline` of north_star): inputs [N, L], bin widths W and heights H [N, L, K] (each summing to 2B),
 * knot derivatives D [N, L, K - 1] (K if periodic) -> outputs [N, L], log_det [N] (negated for the inverse).  Same bin
 * rule (`sum(knots <= x) - 1`), out-of-bounds identity and NaN for a bin index of K as inside the fused chain kernels,
 * which build W, H, D from the conditioner themselves and never call this. */
int pzb_rational_quadratic_spline(const float *inputs, const float *W, const float *H, const float *D, int64_t N,
                                  int32_t L, int32_t K, float B, int32_t periodic, int32_t inverse, float *outputs,
                                  float *log_det, void *stream);

/* ---- Gaussian error convolution (err_samples) ---------------------------
 * Flow.log_prob(inputs, err_samples=S) (pzflow/flow.py:448-464) with the default
 * gaussian_error_model (pzflow/utils.py:52-61): every row is expanded into S
 * samples x + eps * x_err (and cond + eps' * cond_err), evaluated, and reduced as
 * out[i] = log(mean_s exp(lp[i, s])).  The expansion is generated inside the
 * chain kernel and never stored.  Xerr [N, D] / cond_err [N, C] are the stds
 * (cond_err in standard-scaled units, i.e. divided by condition_stds); either
 * may be NULL = zero error (the reference fills missing *_err columns with
 * zeros, flow.py:366-369).  eps comes from a counter-based Philox stream keyed
 * by `seed` (see pzb_error_eps) -- or, after pzb_set_error_stream (below), from
 * jax.random.normal(PRNGKey(seed), ...) itself, which is what the Python mirror
 * asks for by default.  scratch: caller-owned device buffer of N * S floats.
 * row_offset: index of X's first row in the caller's whole array, so that a
 * call cut into row chunks draws exactly the eps of one big call. */
int pzb_log_prob_err(const pzb_plan *plan, const float *X, const float *Xerr, const float *cond,
                     const float *cond_err, int64_t N, int32_t err_samples, uint64_t seed,
                     int64_t row_offset, float *out, float *scratch, void *stream);

/* Flow.posterior(..., err_samples=S) main loop (pzflow/flow.py:675-737): the S
 * samples of every galaxy's other columns (and conditions) are each expanded
 * over the grid, exp(log_prob) is averaged over the samples (flow.py:722-724),
 * then normalised like pzb_posterior.  rest_err [Ng, D-1] may be NULL.
 * scratch: caller-owned device buffer of Ng * S * G floats. */
int pzb_posterior_err(const pzb_plan *plan, const float *rest, const float *rest_err,
                      const float *cond, const float *cond_err, int64_t Ng, int32_t col_idx,
                      const float *grid, int32_t G, int32_t err_samples, uint64_t seed,
                      int64_t row_offset, int32_t normalize, int32_t nan_to_zero, float *pdfs,
                      float *scratch, void *stream);

/* Flow.posterior(..., marg_rules=...) for galaxies that miss the SAME n_mcols (1..4) data columns
 * (pzflow/flow.py:560-664): every galaxy is evaluated in V variants -- its row with the missing columns replaced by
 * mvals[galaxy, v, q] (q-th listed column; mcols[q] = index of that column among the flow's data columns, not
 * col_idx) -- as a second expansion axis generated inside the chain kernel (the reference re-enters posterior() on a
 * host-side repeat of the rows, flow.py:610-643).  With err_samples = S > 0 every variant is additionally
 * averaged over S Gaussian samples of the NON-marginalised columns / conditions (the reference drops the marginalised
 * column's error column, flow.py:625-628); pass 0 for none.  pdf_sums [Ng, G] receives
 *   sum_v nan_to_num?( mean_s exp(log_prob(variant v, sample s, grid point)) )
 * un-normalised, as the reference's inner calls return it (normalize=False, flow.py:640); the caller normalises the
 * assembled array with pzb_normalize_pdfs.  scratch: caller-owned device buffer of Ng * V * max(S, 1) * G floats. */
int pzb_posterior_marg(const pzb_plan *plan, const float *rest, const float *rest_err, const float *cond,
                       const float *cond_err, int64_t Ng, int32_t col_idx, const float *grid, int32_t G,
                       int32_t n_mcols, const int32_t *mcols, const float *mvals, int32_t V, int32_t err_samples,
                       uint64_t seed, int64_t row_offset, int32_t nan_to_zero, float *pdf_sums, float *scratch,
                       void *stream);

/* The standard normals the two calls above use: eps [n, W] for sample index
 * es in [0, n) (log_prob: es = row * S + s; posterior: es = galaxy * S + s) and
 * column j in [0, W); stream_id 0 = data columns (posterior: the D-1 remaining
 * columns in order), 1 = conditions.  Lets a caller (and the parity tests)
 * reproduce the expansion exactly. */
int pzb_error_eps(float *eps, int64_t n, int32_t W, int32_t stream_id, uint64_t seed, void *stream);

/* The error samples in jax's OWN stream.  pzb_set_error_stream(layout, total_units) tells the NEXT pzb_log_prob_err /
 * pzb_posterior_err calls of this host thread to draw eps as the reference does (pzflow/flow.py:455-464, 552-553, 675-693
 * with pzflow/utils.py:52-61): key = PRNGKey(seed) = (0, seed mod 2^32) for data AND conditions, eps of (unit i, sample s,
 * column j) = element ((i * S + s) * W + j) of jax.random.normal(key, (total_units, S, W)), W = D (the posterior's evaluated
 * column rides along, flow.py:370-373) or C.  layout 1 = jax's original threefry layout, 2 = the partitionable layout (jax >=
 * 0.5's default) -- each checked against the values jax's documentation prints for random.normal --, 0 = back to the
 * library's Philox streams.  total_units = rows / galaxies of the caller's WHOLE call (a chunked call passes row_offset as before).  The
 * normals are sqrt(2) * erf_inv(u) with XLA's float32 erf_inv (Giles' polynomials), so they are jax's own values up to an
 * ulp of log1pf; the values jax's documentation prints for random.normal are reproduced digit for digit.
 * pzb_posterior_marg keeps the Philox streams.  pzb_sample_normal_threefry materialises rows [row_offset, row_offset + N) of
 * jax.random.normal(key, (total_rows, W)): Normal.sample (pzflow/distributions.py:297-303) and gaussian_error_model's eps. */
int pzb_set_error_stream(int32_t layout, int64_t total_units);
int pzb_sample_normal_threefry(float *out, int64_t N, int32_t W, uint32_t key0, uint32_t key1, int64_t total_rows,
                               int64_t row_offset, int32_t partitionable, void *stream);

/* ---- training support (the optional data-parallel training step) ---------
 * The spline of a coupling stage for the loss -mean(w * log_prob) (pzflow/flow.py:961-965): softmax / softplus
 * knot construction, bin search, rational-quadratic spline and log-determinant (pzflow/bijectors.py:488-506,
 * pzflow/utils.py:64-227; forward direction) as one kernel, and its hand-derived backward.
 * logits [N, L, 3K-1] (3K when periodic; the conditioner's output), x [N, L] (the transformed columns); forward writes y [N, L] and
 * log_det [N, L] (per element; the caller sums over L).  backward takes grad_y [N, L] and grad_log_det [N, L]
 * and writes grad_logits [N, L, 3K-1] and grad_x [N, L].  The dense layers around it are plain GEMMs (cuBLAS). */
int pzb_spline_train_forward(const float *logits, const float *x, int64_t N, int32_t L, int32_t K, float B,
                             int32_t periodic, float *y, float *log_det, void *stream);
int pzb_spline_train_backward(const float *logits, const float *x, const float *grad_y,
                              const float *grad_log_det, int64_t N, int32_t L, int32_t K, float B,
                              int32_t periodic, float *grad_logits, float *grad_x, void *stream);

/* ---- one optimisation step's gradient as ONE kernel (pzflow/flow.py:961-982, 1053-1084) ------------------------
 * loss = -sum_rows(w * log_prob) / n_global of this rank's N rows and its gradient w.r.t. every conditioner weight:
 * forward and backward through the whole chain in one launch (a CTA per 16 rows; hand-derived backward of the spline
 * and of the dense layers; per-CTA partial gradients summed in a fixed order by a second kernel).  Parameters are
 * read from the caller's FLAT buffer -- the stax pytree's leaves in traversal order, W[in, out] row-major then b[out]
 * per dense layer (pzflow/utils.py:45-48) -- and flat_grad has the same layout, ready for one all-reduce over the
 * data-parallel ranks and pzb_adam_step.
 *   pzb_plan_train_supported: 1 for chains of ShiftBounds / StandardScaler / Scale / Roll / Reverse /
 *     NeuralSplineCoupling (hidden_dim <= 256, (3K-1) * transformed columns <= 1024, K <= 64) with a Uniform,
 *     CentBeta13 or Normal latent; everything else trains through the autograd route of pzflow_b200/train.py.
 *   pzb_plan_set_train_layout: w_offsets / b_offsets [n_layers_total] = offset (in floats) of every dense layer's W
 *     and b inside the flat buffer, coupling stages in chain order, layers in network order; the layers must cover
 *     the n_params floats exactly.
 *   pzb_train_workspace_floats: size of the device workspace a step of N rows needs (saved activations + partial
 *     gradients).
 *   pzb_train_step: writes flat_grad [n_params] (non-finite entries zeroed, as jnp.nan_to_num does for the reference's
 *     update) and loss [1] (this rank's share of the loss; may be NULL).  Rows whose log_prob is not finite carry no
 *     gradient.  row_idx (may be NULL): the N batch rows are X[row_idx[i]] etc. -- the epoch's permutation
 *     (flow.py:1063-1066) applied inside the kernel, so X / cond / weights are the WHOLE training arrays.  Xerr /
 *     cond_err (may be NULL): per-row Gaussian sigmas of convolve_errs (flow.py:1068-1079); batch row i is then
 *     evaluated at x + eps * sigma with eps from the library's Philox stream (seed, noise_base + i). */
int pzb_plan_train_supported(const pzb_plan *plan);
int pzb_plan_set_train_layout(pzb_plan *plan, const int64_t *w_offsets, const int64_t *b_offsets, int32_t n_layers_total,
                              int64_t n_params);
int64_t pzb_train_workspace_floats(const pzb_plan *plan, int64_t N);
int pzb_train_step(const pzb_plan *plan, const float *params, const float *X, const float *cond, const float *weights,
                   int64_t N, int64_t n_global, const int64_t *row_idx, const float *Xerr, const float *cond_err,
                   uint64_t seed, uint64_t noise_base, float *workspace, float *flat_grad, float *loss, void *stream);

/* Adam over one flat fp32 parameter buffer (optax.adam(1e-3) as pzflow/flow.py:968-981 uses it; eps_root = 0).
 * step_dev is a device int32 holding the number of steps taken so far; the call increments it on the stream
 * before the update, so the whole call can be captured in a CUDA graph. */
int pzb_adam_step(float *params, const float *grads, float *m, float *v, int64_t n, float lr, float b1,
                  float b2, float eps, int32_t *step_dev, void *stream);

/* Host threads (casts, gathers, staging copies) the host-buffer entry points below may use when called from THIS
 * host thread; 0 restores the default (the process's CPU affinity, shared between torchrun's local ranks, at most 16).
 * A caller that drives several devices from several host threads gives each its share. */
int pzb_set_host_threads(int32_t n);

/* (no reference counterpart) Content check of host arrays for the Python mirror's plan cache: sums[2i] = wrapping sum of
 * leaf i's 8-byte words, sums[2i + 1] = sum of its tail bytes; all n leaves in one call on the host cores.  A cached
 * plan is reused only while the parameter arrays it was built from are the same objects AND still hold the same bytes
 * (in-place edits of a weight array must not evaluate stale weights).  No device work. */
int pzb_host_digest(const void *const *leaves, const int64_t *nbytes, int32_t n, uint64_t *sums);

/* ---- host-buffer forms --------------------------------------------------
 * The same evaluations for callers whose arrays live in HOST memory, which is
 * where the reference's callers hold them (NumPy arrays cut from a DataFrame,
 * pzflow/flow.py:440-446, 697-737, 783-795).  All data pointers below are
 * HOST pointers.  Rows are cut into chunks that move through a three-deep ring
 * of device staging buffers on plan-owned streams, so H2D copy, chain kernel
 * and D2H copy of consecutive chunks overlap and HBM holds three chunks at
 * most (the 4 GB posterior of config 3 never sits on the device as a whole).
 * Blocking: the results are in the output arrays on return.  Page-locked host
 * arrays are copied from / into directly; pageable arrays (plain NumPy arrays)
 * of 8 MB or more go through plan-owned page-locked staging filled / emptied
 * by all host cores inside the same pipeline, at the same speed.  Calls on one
 * plan serialise. */
int pzb_log_prob_host(const pzb_plan *plan, const float *X, const float *cond, int64_t N,
                      float *out);
/* Flow.log_prob straight from DataFrame columns (pzflow/flow.py:440-446): data_cols [D] and cond_cols [C]
 * are pointers to contiguous float64 HOST columns of N values (what `df[c].to_numpy()` is for a pandas
 * column); the float64 -> float32 down-cast, the row gather and the standard scaling of the conditions
 * ((float(c) - mean) / std, flow.py:330-336) happen on the host cores chunk by chunk into page-locked staging
 * inside the same three-stage pipeline, so they overlap with the copies and the kernel.  out [N] HOST. */
int pzb_log_prob_host_columns(const pzb_plan *plan, const double *const *data_cols,
                              const double *const *cond_cols, const float *cond_means,
                              const float *cond_stds, int64_t N, float *out);
/* The same for float32 columns (a DataFrame read from float32 data). */
int pzb_log_prob_host_columns_f32(const pzb_plan *plan, const float *const *data_cols,
                                  const float *const *cond_cols, const float *cond_means,
                                  const float *cond_stds, int64_t N, float *out);
int pzb_forward_host(const pzb_plan *plan, const float *X, const float *cond, int64_t N,
                     float *U, float *log_det);
int pzb_inverse_host(const pzb_plan *plan, const float *U, const float *cond, int64_t N,
                     float *X, float *log_det);
/* FlowEnsemble.log_prob (pzflow/flowEnsemble.py:189-243, returnEnsemble=False) from DataFrame columns: every chunk
 * of rows is staged ONCE, evaluated by all n_flows plans (same device, input_dim and n_conditions) and reduced with
 * log(mean(exp(lp), axis=flows)) on the device, inside the pipeline of flows[0].  Columns are float64, or float32
 * when cols_are_f32 != 0; the conditions are standard-scaled with ONE mean / std pair (the flows of an ensemble are
 * trained on the same data; a caller whose flows disagree evaluates them one by one).  out [N] HOST. */
int pzb_ensemble_log_prob_host_columns(const pzb_plan *const *flows, int32_t n_flows,
                                       const void *const *data_cols, int32_t cols_are_f32,
                                       const void *const *cond_cols, const float *cond_means,
                                       const float *cond_stds, int64_t N, float *out);

/* FlowEnsemble.posterior (pzflow/flowEnsemble.py:245-351, returnEnsemble=False) on HOST arrays: every chunk of
 * galaxies is staged once, each flow writes its un-normalised pdfs (NaN -> 0 per flow if nan_to_zero), the mean over
 * flows is taken and trapezoid-normalised (if normalize) on the device, and only the [Ng, G] result travels back.
 * Arguments as pzb_posterior_host; the flows share device, input_dim and n_conditions, and `cond` is the (one)
 * standard-scaled conditions array. */
int pzb_ensemble_posterior_host(const pzb_plan *const *flows, int32_t n_flows, const float *rest,
                                const float *cond, int64_t Ng, int32_t col_idx, const float *grid, int32_t G,
                                int32_t normalize, int32_t nan_to_zero, float *pdfs);

/* Flow.sample (pzflow/flow.py:783-795) for latent draws that were made ON THE DEVICE (pzb_sample_uniform, or any
 * device array U_dev [N, D]; cond_dev [N, C] device or NULL): x = inverse(u) chunk by chunk, results in the HOST
 * arrays X [N, D] and log_det [N] (or NULL) through the same pipeline (kernel | D2H | host copy overlap, pageable
 * destinations staged).  after_stream: the stream the draws were produced on; it is synchronised first. */
int pzb_inverse_to_host(const pzb_plan *plan, const float *U_dev, const float *cond_dev, int64_t N, float *X,
                        float *log_det, void *after_stream);
/* The same with the samples delivered as D separate HOST column arrays X_cols[j][N] -- the layout a pandas
 * DataFrame keeps its values in, so that Flow.sample's `pd.DataFrame(x, columns=...)` (flow.py:797-817) wraps the
 * result without the transposing copy a row-major [N, D] array costs it (160 ms for 10 M rows of 7 columns). */
int pzb_inverse_to_host_columns(const pzb_plan *plan, const float *U_dev, const float *cond_dev, int64_t N,
                                float *const *X_cols, void *after_stream);
int pzb_posterior_host(const pzb_plan *plan, const float *rest, const float *cond, int64_t Ng,
                       int32_t col_idx, const float *grid, int32_t G, int32_t normalize,
                       int32_t nan_to_zero, float *pdfs);

#ifdef __cplusplus
}
#endif
#endif /* PZFLOW_B200_H */
