// Device-side plan layout shared by the engines (SIMT FFMA and tcgen05).
//
// A plan is the flattened bijector chain of one Flow (pzflow/bijectors.py:121-174):
// affine stages (ShiftBounds / StandardScaler), NeuralSplineCoupling stages and
// the Chain nesting markers that fix the order log-determinants are summed in.
// Roll stages never move data: they are folded into per-stage column
// permutations at plan time (pzflow/bijectors.py:585-593).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pzb {

constexpr int TM = 64;          // rows per CTA tile (SIMT engine)
constexpr int NTHREADS = 256;   // threads per CTA (SIMT engine)
constexpr int KC = 32;          // k-rows per streamed weight chunk
constexpr int MAX_NC = 192;     // widest GEMM column chunk
constexpr int MAX_D = 64;       // data columns
constexpr int MAX_KEY = 128;    // conditioner inputs (upper + conditions)
constexpr int MAX_DENSE = 8;    // dense layers per conditioner (hidden_layers + 1)
constexpr int MAX_NSC = 128;
constexpr int MAX_AFFINE = 16;
constexpr int MAX_STEPS = 320;
constexpr int MAX_CHAIN_DEPTH = 4;

constexpr int MAX_COLOR = 8;
enum StepKind : int { STEP_AFFINE = 0, STEP_NSC = 1, STEP_CHAIN_BEGIN = 2, STEP_CHAIN_END = 3, STEP_COLOR = 4 };
enum AffKind : int { AFF_SHIFT_BOUNDS = 1, AFF_STANDARD_SCALER = 2, AFF_SCALE = 3, AFF_INV_SOFTPLUS = 4, AFF_DEQUANTIZE = 5 };
enum Mode : int { MODE_LOGPROB = 0, MODE_FORWARD = 1, MODE_INVERSE = 2, MODE_POSTERIOR = 3 };

// One dense layer, packed chunk-major: W[n_chunks][in_dim][NC], b[n_chunks][NC]
// (zero padded). Hidden layers use NC = 128; the output layer's chunks hold
// whole spline dimensions (dpc dims x cpd logits).
struct DenseDev {
  const float* W;
  const float* b;
  int in_dim, out_dim, NC, n_chunks;
};

// Extra packing for the tcgen05 engine (H = 128 conditioners): fp16 hi/lo
// split weights in the UMMA canonical K-major SWIZZLE_128B layout.
struct TcLayerDev {
  const void* blob;   // [W2 hi | W2 lo | W3 hi | W3 lo | W1 hi | W1 lo | par] for one NSC stage (nsf_tc.cu)
  int blob_bytes;
};

constexpr int TC_SUB_MAX = 4;  // tensor-core engine: a coupling stage's spline dimensions are taken four at a time

struct NscDev {
  int U, L, C, K, cpd, dpc, periodic, n_dense, H, Hpad, bins_off;
  float B;
  DenseDev dense[MAX_DENSE];
  int n_sub;                     // tensor-core sub-stages of this stage: ceil(L / 4)
  TcLayerDev tc[TC_SUB_MAX];
  const void* wide_blob;         // wide tensor-core engine: [parameter block | W1 slabs | hidden slabs | output slabs] (nsf_wide.cu)
  int wide_bytes;
  int16_t perm[MAX_D];  // logical column j lives in physical column perm[j]
};

struct AffineDev {
  int kind;
  float B;
  float a[MAX_D];  // physical index: SB mean / SS means
  float b[MAX_D];  // physical index: SB half_range / SS stds
  float logdet_fwd, logdet_inv;
};

// ColorTransform (pzflow/bijectors.py:177-311): magnitudes -> (reference magnitude, adjacent colours).
// slot[i] = physical column of magnitude i (mag_idx order) before the stage; after it slot[n_mag - 1] holds the
// reference magnitude and slot[i], i < n_mag - 1, the colour mag[i] - mag[i + 1].
struct ColorDev {
  int n_mag, ref;  // ref: position of ref_idx inside mag_idx
  int16_t slot[MAX_D];
};

struct StepDev {
  int kind, idx;
};

constexpr int MAX_LAT = 8;
// kind: 1 Uniform, 2 CentBeta13, 3 Normal, 4 CentBeta, 5 Tdist.  c0/c1/c2 are per-kind constants:
//   Uniform: c0 = -n log(2B); CentBeta13: c0 = betaln(13, 13); Tdist: c0 = gammaln(t) - gammaln(nu/2) - n/2 log(nu pi),
//   c1 = t = (nu + n)/2, c2 = 1/nu (pzflow/distributions.py:330-358 with cov = I).
struct LatentSeg {
  int kind, start, n;
  float B, c0, c1, c2;
};

struct PlanDev {
  int D, C, n_steps, n_nsc, n_affine, latent_kind, n_spline_dims;
  float latent_B;
  float latent_const;  // Uniform: -D * log(2B); CentBeta13: betaln(13, 13) in f32 arithmetic
  int16_t perm_final[MAX_D];
  // CentBeta (pzflow/distributions.py:45-135): per logical dimension alpha, beta and betaln(alpha, beta)
  float latent_a[MAX_D], latent_b[MAX_D], latent_betaln[MAX_D];
  // The latent as a list of segments over consecutive logical dimensions: one segment for the plain latents,
  // several for Joint (pzflow/distributions.py:473-562).
  int n_lat;
  LatentSeg lat[MAX_LAT];
  int n_color;
  ColorDev color[MAX_COLOR];
  StepDev steps[MAX_STEPS];
  AffineDev affine[MAX_AFFINE];
  NscDev nsc[MAX_NSC];
};

struct LaunchArgs {
  int mode;
  const float* X;     // MODE_LOGPROB/FORWARD: X[N,D]; INVERSE: U[N,D]; POSTERIOR: rest[Ng,D-1]
  const float* cond;  // [N,C] (POSTERIOR: [Ng,C]) or nullptr
  const float* grid;  // POSTERIOR: [G]
  int G, col_idx;
  long long N;        // rows to evaluate (POSTERIOR: Ng*G)
  float* out0;        // lp[N] | U[N,D] | X[N,D] | pdfs[Ng*G]
  float* out1;        // log_det[N] (FORWARD/INVERSE, may be null)
  int32_t* bins;      // optional [N, n_spline_dims]
  // Gaussian error convolution (pzflow/flow.py:339-395, pzflow/utils.py:52-61), S = err_samples > 0:
  // every unit (row / galaxy) is expanded into S samples x + eps * x_err generated here, never stored.
  // N then counts the EXPANDED rows: LOGPROB units*S (sample-minor), POSTERIOR units*S*G (grid-minor).
  const float* Xerr;      // same shape as X (stds), or nullptr (no data errors)
  const float* cond_err;  // [units, C] stds in standard-scaled condition units, or nullptr
  int S;
  unsigned long long seed;
  long long unit0;        // index of X's first row in the caller's whole array: chunked calls draw the same eps
  // Marginalisation axis (pzflow/flow.py:560-664), V > 0: every unit (galaxy) is evaluated in V variants that differ in
  // n_mcols data columns (mcol[q] = logical column index, -1 unused) whose values come from mvals[unit, v, q].  N then
  // counts units * V * max(S, 1) * G rows: unit-major, then variant, then error sample, then grid point.
  const float* mvals;
  int V, n_mcols;
  int mcol[4];
  // Fused posterior reduction (tensor-core engine, MODE_POSTERIOR): out0 is the FINAL pdfs [units, G] -- the mean over
  // the error samples, the sum over the variants, the per-galaxy trapezoid normalisation and NaN -> 0
  // (pzflow/flow.py:720-737) happen in shared memory before the only HBM write.
  int fuse, normalize, nan_to_zero;
  // set to 1 by a tensor-core launch in which some row of finite inputs got non-finite logits (an fp16 overflow of a
  // conditioner pre-activation): the host pipeline's per-chunk guard (pzb_api.cu: run_host); may be null
  unsigned int* ovf_flag;
  // The error samples in jax's OWN stream (eps_layout 1: jax's original threefry layout, 2: the partitionable one;
  // 0: this library's Philox streams): eps of (unit i, sample s, column j) is element ((i * S + s) * W + j) of
  // jax.random.normal(key, (eps_units, S, W)) -- W = D for data (the posterior's evaluated column rides along as in
  // pzflow/flow.py:370-373), C for conditions; key = (eps_k0, eps_k1); eps_units = units of the WHOLE call.
  int eps_layout;
  uint32_t eps_k0, eps_k1;
  long long eps_units;
};

// ---- counter-based normals for the error samples: Philox-4x32-10 keyed by the seed, counter =
// (sample index, 4-column block, stream: 0 data / 1 conditions), Box-Muller on the four words ----
__device__ __forceinline__ void pz_philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}

__device__ __forceinline__ void pz_philox4(unsigned long long seed, unsigned long long ctr, uint32_t block,
                                           uint32_t stream, uint32_t (&c)[4]) {
  c[0] = (uint32_t)ctr;
  c[1] = (uint32_t)(ctr >> 32);
  c[2] = block;
  c[3] = stream;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    pz_philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Threefry-2x32, 20 rounds (Salmon et al., SC'11): the block function of jax's default PRNG.
__device__ __forceinline__ void pz_threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  constexpr int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = __funnelshift_l(x1, x1, R[i & 1][r]) ^ x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
}

// The 32 random bits jax's random_bits gives element g of an array of n_total elements.  Original layout: the count
// vector iota(n_total) rides through the block function as two halves (g < h pairs with g + h, a zero pads an odd
// count); partitionable layout: one block per element, (hi32(g), lo32(g)), its two output words xor-ed.
__device__ __forceinline__ uint32_t pz_jax_bits(uint32_t k0, uint32_t k1, long long g, long long n_total, bool partitionable) {
  uint32_t x0, x1;
  if (partitionable) {
    x0 = (uint32_t)((unsigned long long)g >> 32);
    x1 = (uint32_t)g;
    pz_threefry2x32(k0, k1, x0, x1);
    return x0 ^ x1;
  }
  const long long h = (n_total + 1) / 2;
  if (g < h) {
    x0 = (uint32_t)g;
    x1 = (g + h < n_total) ? (uint32_t)(g + h) : 0u;
    pz_threefry2x32(k0, k1, x0, x1);
    return x0;
  }
  x0 = (uint32_t)(g - h);
  x1 = (uint32_t)g;
  pz_threefry2x32(k0, k1, x0, x1);
  return x1;
}

// XLA's float32 erf_inv (M. Giles, "Approximating the erfinv function", GPU Computing Gems 2011: two degree-8
// polynomials in w = -log1p(-x * x), evaluated by Horner without contraction) -- what jax.random.normal runs through.
__device__ __forceinline__ float pz_xla_erfinv(float x) {
  float w = -log1pf(__fmul_rn(-x, x));
  const bool lt = w < 5.0f;
  w = lt ? __fsub_rn(w, 2.5f) : __fsub_rn(sqrtf(w), 3.0f);
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
  p = __fadd_rn(lt ? 3.43273939e-07f : 0.000100950558f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -3.5233877e-06f : 0.00134934322f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -4.39150654e-06f : -0.00367342844f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 0.00021858087f : 0.00573950773f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -0.00125372503f : -0.0076224613f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? -0.00417768164f : 0.00943887047f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 0.246640727f : 1.00167406f, __fmul_rn(p, w));
  p = __fadd_rn(lt ? 1.50140941f : 2.83297682f, __fmul_rn(p, w));
  return __fmul_rn(p, x);
}

// jax.random.normal from those bits: u = uniform in (-1, 1) -- 23 mantissa bits | 1.0f, minus 1, times (1 - lo) = 2,
// plus lo = nextafter(-1, 0), clamped below at lo -- then sqrt(2) * erf_inv(u)  (jax/_src/random.py: _normal_real).
// With the polynomial above this reproduces the values jax's documentation prints for random.normal digit for digit
// (tests/test_oracle.py: jax_normal_known_answers); log1pf is CUDA's, within an ulp of XLA's.
__device__ __forceinline__ float pz_jax_normal(uint32_t bits) {
  const float lo = -0.99999994f;
  const float f = __fsub_rn(__uint_as_float((bits >> 9) | 0x3f800000u), 1.0f);
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
  return __fmul_rn(1.4142135623730951f, pz_xla_erfinv(u));
}

// One error sample in jax's stream.  Not inlined: the chain kernels' input stage only pays a call for it, and only on
// error-sample launches, so the register allocation of their hot loops stays what it was.
static __device__ __noinline__ float pz_jax_eps(uint32_t k0, uint32_t k1, long long g, long long n_total, int layout) {
  return pz_jax_normal(pz_jax_bits(k0, k1, g, n_total, layout == 2));
}

// eps for (sample index es, column j) of a stream: standard normal, float32
__device__ __forceinline__ float pz_err_normal(unsigned long long seed, uint32_t stream, unsigned long long es, int j) {
  uint32_t c[4];
  pz_philox4(seed, es, (uint32_t)(j >> 2), stream, c);
  const int pair = (j >> 1) & 1;
  // u1 in (0, 1], u2 in [0, 1): 24 random bits each
  const float u1 = ((float)(c[2 * pair] >> 8) + 1.0f) * 5.9604644775390625e-8f;
  const float u2 = (float)(c[2 * pair + 1] >> 8) * 5.9604644775390625e-8f;
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return (j & 1) ? r * sn : r * cs;
}

// Where evaluation row `row` sits on the expansion axes (posterior grid point, error sample, marginal variant, unit):
// computed ONCE per row -- the divisions are 64-bit and there is one per axis.
struct RowIndex {
  long long es;    // error-sample index (== variant == unit without those axes): keys the eps stream
  long long vs;    // variant index: row of mvals
  long long unit;  // row of X / cond
  int g;           // posterior grid point
};
__device__ __forceinline__ RowIndex pz_row_index(const LaunchArgs& a, long long row) {
  RowIndex r;
  if (a.mode == MODE_POSTERIOR) {
    // rows of one launch stay below 2^31 in every caller of this library (chunked host pipeline, bounded scratch):
    // take the cheap 32-bit division when they do
    if (row < 0x7fffffffLL) {
      const unsigned q = (unsigned)row / (unsigned)a.G;
      r.es = q;
      r.g = (int)((unsigned)row - q * (unsigned)a.G);
    } else {
      r.es = row / a.G;
      r.g = (int)(row - r.es * a.G);
    }
  } else {
    r.es = row;
    r.g = 0;
  }
  r.vs = a.S > 0 ? r.es / a.S : r.es;
  r.unit = a.V > 0 ? r.vs / a.V : r.vs;
  return r;
}

// Value of data column j (logical order) and of condition j for an evaluation row: the posterior's grid expansion
// (pzflow/flow.py:697-714), the marginal variants and the Gaussian error samples are generated here.
__device__ __forceinline__ float pz_input_value(const LaunchArgs& a, const RowIndex& r, int j, int D) {
  int jj = j, W = D;
  if (a.mode == MODE_POSTERIOR) {
    if (j == a.col_idx) return a.grid[r.g];
    jj = j < a.col_idx ? j : j - 1;
    W = D - 1;
  }
  if (a.V > 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < a.n_mcols && a.mcol[q] == j) return a.mvals[r.vs * a.n_mcols + q];  // a marginalised column: no error model
  }
  float v = a.X[r.unit * W + jj];
  if (a.S > 0 && a.Xerr != nullptr) {
    const long long es = r.es + a.unit0 * a.S * (a.V > 0 ? a.V : 1);
    const float eps = a.eps_layout == 0
                          ? pz_err_normal(a.seed, 0u, (unsigned long long)es, jj)
                          : pz_jax_eps(a.eps_k0, a.eps_k1, es * D + j, a.eps_units * a.S * D, a.eps_layout);
    v = v + eps * a.Xerr[r.unit * W + jj];
  }
  return v;
}

__device__ __forceinline__ float pz_cond_value(const LaunchArgs& a, const RowIndex& r, int j, int C) {
  if (a.cond == nullptr) return 0.0f;
  float v = a.cond[r.unit * C + j];
  if (a.S > 0 && a.cond_err != nullptr) {
    const long long es = r.es + a.unit0 * a.S * (a.V > 0 ? a.V : 1);
    const float eps = a.eps_layout == 0
                          ? pz_err_normal(a.seed, 1u, (unsigned long long)es, j)
                          : pz_jax_eps(a.eps_k0, a.eps_k1, es * C + j, a.eps_units * a.S * C, a.eps_layout);
    v = v + eps * a.cond_err[r.unit * C + j];
  }
  return v;
}

// One element of an "affine" (elementwise) stage; returns the element's log-det contribution where the stage has
// a data-dependent one (InvSoftplus), 0 otherwise (constant log-dets are added once per row by the caller).
//   ShiftBounds   fwd B*(x-mean)/half_range ; inv x*half_range/B + mean          (bijectors.py:760-773)
//   StandardScaler fwd (x-means)/stds ; inv x*stds + means                        (bijectors.py:848-857)
//   Scale         fwd scale*x ; inv (1/scale)*x                                   (bijectors.py:689-711)
//   UniformDequantizer fwd x (as the reference runs it) ; inv floor(x) on the listed columns (bijectors.py:892-907)
//   InvSoftplus   fwd log(-1+exp(k x))/k, ld log(1+exp(-k y)) ; inv log(1+exp(k x))/k, ld -log(1+exp(-k x))
//                 on the listed columns (k = b[j] != 0)                             (bijectors.py:350-369)
__device__ __forceinline__ float pz_affine_elem(const AffineDev& af, bool inverse, int j, float x, float& ld) {
  ld = 0.0f;
  switch (af.kind) {
    case AFF_SHIFT_BOUNDS:
      return inverse ? (x * af.b[j]) / af.B + af.a[j] : (af.B * (x - af.a[j])) / af.b[j];
    case AFF_STANDARD_SCALER:
      return inverse ? x * af.b[j] + af.a[j] : (x - af.a[j]) / af.b[j];
    case AFF_SCALE:
      return inverse ? (1.0f / af.B) * x : af.B * x;
    case AFF_DEQUANTIZE:  // forward identity, inverse floor of the flagged columns (bijectors.py:892-907); log-det 0
      return (inverse && af.b[j] != 0.0f) ? floorf(x) : x;
    default: {
      const float k = af.b[j];
      if (k == 0.0f) return x;
      if (inverse) {
        ld = -logf(1.0f + expf((-k) * x));
        return logf(1.0f + expf(k * x)) / k;
      }
      const float y = logf(-1.0f + expf(k * x)) / k;
      ld = logf(1.0f + expf((-k) * y));
      return y;
    }
  }
}

// ColorTransform on one row whose column c lives at st[c * stride] (log-det 0)
__device__ __forceinline__ void pz_color_row(const ColorDev& cd, bool inverse, float* st, int stride) {
  const int n = cd.n_mag;
  if (!inverse) {
    float prev = st[cd.slot[0] * stride];
    const float refv = st[cd.slot[cd.ref] * stride];
    for (int i = 0; i + 1 < n; ++i) {
      const float nxt = st[cd.slot[i + 1] * stride];
      st[cd.slot[i] * stride] = -(nxt - prev);  // -jnp.diff(mags)
      prev = nxt;
    }
    st[cd.slot[n - 1] * stride] = refv;
  } else {
    const float refv = st[cd.slot[n - 1] * stride];
    // cumsum of the colours: mag[0] - mag[i + 1]
    float run = 0.0f, at_ref = 0.0f;
    for (int i = 0; i + 1 < n; ++i) {
      run = (i == 0) ? st[cd.slot[0] * stride] : run + st[cd.slot[i] * stride];
      if (i + 1 == cd.ref) at_ref = run;
    }
    const float mag0 = cd.ref == 0 ? refv : refv + at_ref;
    run = 0.0f;
    float carry = mag0;  // value to store into slot[i]: mag[i]
    for (int i = 0; i + 1 < n; ++i) {
      const float c = st[cd.slot[i] * stride];
      run = (i == 0) ? c : run + c;
      st[cd.slot[i] * stride] = carry;
      carry = mag0 - run;
    }
    st[cd.slot[n - 1] * stride] = carry;
  }
}

// ---- small numerics shared by both engines (reference op order, no FMA contraction:
// the files are compiled with -fmad=false and GEMM loops call fmaf explicitly) ----

__device__ __forceinline__ float pz_softplus(float x) {
  // jax.nn.softplus = logaddexp(x, 0)
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float pz_leaky(float x) {
  // stax.LeakyRelu: where(x >= 0, x, 0.01 * x) == max(x, 0.01 * x), bit for bit (FMUL + FMNMX)
  return fmaxf(x, 0.01f * x);
}

__device__ __forceinline__ float pz_nan_to_num_lp(float lp) {
  // jnp.nan_to_num(lp, nan=-inf): NaN -> -inf, then +/-inf -> finfo.max/min (pzflow/flow.py:406)
  if (isnan(lp) || lp == -INFINITY) return -3.4028234663852886e38f;
  if (lp == INFINITY) return 3.4028234663852886e38f;
  return lp;
}

struct SplineKnotSel {
  float wk, hk, xk, yk, dk, dkp1;
  int idx;
};

__device__ __forceinline__ float pz_lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// pzflow/utils.py:163-227 once the bin's seven values are known.  FASTLOG evaluates the three
// logarithms of the log-determinant with lg2.approx (absolute error ~2e-7 each, far below the
// float32 noise floor of the path); the SIMT engine keeps logf.
template <bool FASTLOG = false>
__device__ __forceinline__ void pz_rqs_eval(const SplineKnotSel& s, float x, float mx, bool oob,
                                            bool inverse, bool periodic, float& out, float& ld) {
  const float sk = s.hk / s.wk;
  const float dsum = (s.dkp1 + s.dk) - 2.0f * sk;
  float relx;
  if (inverse) {
    const float t = mx - s.yk;
    const float a = s.hk * (sk - s.dk) + t * dsum;
    const float b = s.hk * s.dk - t * dsum;
    const float c = (-sk) * t;
    relx = (2.0f * c) / ((-b) - sqrtf(b * b - (4.0f * a) * c));
    out = relx * s.wk + s.xk;
  } else {
    relx = (mx - s.xk) / s.wk;
    const float num = s.hk * (sk * (relx * relx) + (s.dk * relx) * (1.0f - relx));
    const float den = sk + (dsum * relx) * (1.0f - relx);
    out = s.yk + num / den;
  }
  const float omr = 1.0f - relx;
  const float dnum = (s.dkp1 * (relx * relx) + ((2.0f * sk) * relx) * omr) + s.dk * (omr * omr);
  const float dden = sk + (dsum * relx) * omr;
  if (FASTLOG)
    ld = 0.6931471805599453f * ((2.0f * pz_lg2_approx(sk) + pz_lg2_approx(dnum)) - 2.0f * pz_lg2_approx(dden));
  else
    ld = (2.0f * logf(sk) + logf(dnum)) - 2.0f * logf(dden);
  if (!periodic && oob) {
    out = x;
    ld = 0.0f;
  }
}

// Latent log-density of one row given u in LOGICAL column order through get(j): the sum over the latent's
// segments (Joint: pzflow/distributions.py:528-541; every other latent is a single segment).
template <typename GetU>
__device__ __forceinline__ float pz_latent_logprob(const PlanDev* plan, GetU get) {
  float total = 0.0f;
  for (int si = 0; si < plan->n_lat; ++si) {
    const LatentSeg sg = plan->lat[si];
    const float B = sg.B;
    float lp;
    if (sg.kind == 1) {  // Uniform, pzflow/distributions.py:414-441
      bool inside = true;
      for (int j = sg.start; j < sg.start + sg.n; ++j) {
        const float u = get(j);
        inside = inside && (u >= -B) && (u <= B);
      }
      lp = inside ? sg.c0 : -INFINITY;
    } else if (sg.kind == 3) {  // Normal, pzflow/distributions.py:255-275: multivariate_normal.logpdf(u, 0, I)
      float ss = 0.0f;
      for (int j = sg.start; j < sg.start + sg.n; ++j) {
        const float u = get(j);
        ss += u * u;
      }
      lp = -0.5f * ((float)sg.n * 1.8378770664093453f + ss);
    } else if (sg.kind == 5) {  // Tdist, pzflow/distributions.py:330-358 (cov = I: maha = |u|^2, log_det = 0)
      float ss = 0.0f;
      for (int j = sg.start; j < sg.start + sg.n; ++j) {
        const float u = get(j);
        ss += u * u;
      }
      lp = sg.c0 + (-sg.c1) * logf(1.0f + sg.c2 * ss);
    } else {
      // CentBeta13 (distributions.py:165-194) / CentBeta (45-99): sum_i beta.logpdf(u_i; a_i, b_i, loc=-B, scale=2B)
      //   = -betaln(a, b) + xlogy(a - 1, y) + xlog1py(b - 1, -y) - log(scale),  y = (u + B) / 2B
      const bool general = sg.kind == 4;
      const float loc = -B, scale = 2.0f * B;
      const float lscale = logf(scale);
      lp = 0.0f;
      for (int j = sg.start; j < sg.start + sg.n; ++j) {
        const float u = get(j);
        const float y = (u - loc) / scale;
        const float am1 = general ? plan->latent_a[j] - 1.0f : 12.0f, bm1 = general ? plan->latent_b[j] - 1.0f : 12.0f;
        const float betaln = general ? plan->latent_betaln[j] : sg.c0;
        // xlogy(0, .) = 0 (scipy semantics), only reachable for the general form
        const float t1 = (general && am1 == 0.0f) ? 0.0f : am1 * logf(y);
        const float t2 = (general && bm1 == 0.0f) ? 0.0f : bm1 * log1pf(-y);
        float l1 = ((-betaln) + (t1 + t2)) - lscale;
        if (u > loc + scale || u < loc) l1 = -INFINITY;
        lp += l1;
      }
    }
    total = (si == 0) ? lp : total + lp;
  }
  return total;
}

}  // namespace pzb
