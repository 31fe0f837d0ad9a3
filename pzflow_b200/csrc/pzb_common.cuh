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

enum StepKind : int { STEP_AFFINE = 0, STEP_NSC = 1, STEP_CHAIN_BEGIN = 2, STEP_CHAIN_END = 3 };
enum AffKind : int { AFF_SHIFT_BOUNDS = 1, AFF_STANDARD_SCALER = 2 };
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
  const void* blob;   // [W2 hi | W2 lo | W3 hi | W3 lo | W1 | b1 | b2 | b3] for one NSC stage
  int blob_bytes;
};

struct NscDev {
  int U, L, C, K, cpd, dpc, periodic, n_dense, H, Hpad, bins_off;
  float B;
  DenseDev dense[MAX_DENSE];
  TcLayerDev tc;
  int16_t perm[MAX_D];  // logical column j lives in physical column perm[j]
};

struct AffineDev {
  int kind;
  float B;
  float a[MAX_D];  // physical index: SB mean / SS means
  float b[MAX_D];  // physical index: SB half_range / SS stds
  float logdet_fwd, logdet_inv;
};

struct StepDev {
  int kind, idx;
};

struct PlanDev {
  int D, C, n_steps, n_nsc, n_affine, latent_kind, n_spline_dims;
  float latent_B;
  float latent_const;  // Uniform: -D * log(2B); CentBeta13: betaln(13, 13) in f32 arithmetic
  int16_t perm_final[MAX_D];
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
};

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

// Latent log-density of one row given u in LOGICAL column order through get(j).
template <typename GetU>
__device__ __forceinline__ float pz_latent_logprob(int latent_kind, int D, float B, float lconst,
                                                   GetU get) {
  if (latent_kind == 1) {  // Uniform, pzflow/distributions.py:414-441
    bool inside = true;
    for (int j = 0; j < D; ++j) {
      const float u = get(j);
      inside = inside && (u >= -B) && (u <= B);
    }
    return inside ? lconst : -INFINITY;
  }
  // CentBeta13, pzflow/distributions.py:165-194: sum_i beta.logpdf(u_i; 13, 13, -B, 2B)
  const float betaln = lconst;  // gammaln(13) + gammaln(13) - gammaln(26), f32 arithmetic
  const float loc = -B, scale = 2.0f * B;
  const float lscale = logf(scale);
  float acc = 0.0f;
  for (int j = 0; j < D; ++j) {
    const float u = get(j);
    const float y = (u - loc) / scale;
    const float lin = 12.0f * logf(y) + 12.0f * log1pf(-y);
    float lp = ((-betaln) + lin) - lscale;
    if (u > loc + scale || u < loc) lp = -INFINITY;
    acc += lp;
  }
  return acc;
}

}  // namespace pzb
