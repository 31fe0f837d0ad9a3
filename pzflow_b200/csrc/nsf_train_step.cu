// One optimisation step's gradient as ONE kernel (pzflow/flow.py:961-982, 1053-1084):
// the loss -sum(w * log_prob) / n of a batch and its gradient w.r.t. every conditioner weight, forward and backward
// through the whole bijector chain, hand-derived, fp32 FFMA.  A CTA owns 16 rows of the batch: it walks the chain
// forward (saving each coupling stage's inputs, hidden activations and logits in an L2-resident scratch), evaluates
// the latent log-density, then walks the chain backward: spline backward (analytic partials, as nsf_train.cu), and per
// dense layer the weight gradient h^T g (written to this CTA's slice of a partial-gradient buffer -- no atomics, so
// the sum over CTAs taken by train_reduce_kernel is deterministic), the bias gradient and the input gradient g W^T.
// The caller all-reduces the flat gradient over the data-parallel ranks (NCCL) and applies Adam (adam_step_kernel):
// one step is this kernel, the reduction, the all-reduce and Adam -- no library GEMM, no autograd graph.
//
// Parameters are read from (and gradients laid out like) the caller's FLAT parameter buffer: the stax pytree's leaves
// in traversal order, W[in, out] row-major then b[out] per dense layer (pzflow/utils.py:45-48).
//
// Supported (everything else keeps the torch autograd route of pzflow_b200/train.py): chains of ShiftBounds /
// StandardScaler / Scale / Roll / Reverse / NeuralSplineCoupling (hidden_dim <= 256, (3K-1) * transformed columns
// <= 1024, K <= 64), latents Uniform / CentBeta13 / Normal.
#include <math.h>

#include "pzb_common.cuh"

namespace pzb {

constexpr int TS_ROWS = 16;
constexpr int TS_THREADS = 256;
constexpr int TS_HMAX = 256;
constexpr int TS_OUT_MAX = 1024;
constexpr int TS_KMAX = 64;
constexpr int TS_KC = 32;    // weight rows per streamed chunk (forward GEMM, 256 output columns at a time), double buffered
constexpr int TS_NC = 32;    // weight columns per streamed chunk (input-gradient GEMM), double buffered
constexpr float TS_SLOPE = 0.01f;

struct TrainOffsets {  // flat-buffer offsets (in floats) of a coupling stage's dense layers
  long long w[MAX_DENSE], b[MAX_DENSE];
};

struct TrainBatch {  // how a step's rows are picked out of the caller's whole arrays
  const long long* row_idx;       // [N] indices into X / cond / weights (the epoch's permutation, this rank's share), or null
  const float* Xerr;              // [rows of X, D] Gaussian sigmas (convolve_errs), or null
  const float* cond_err;          // [rows of X, C] sigmas in standard-scaled condition units, or null
  unsigned long long seed, noise_base;   // eps of batch row i: Philox(seed, noise_base + i, column)
};

// [col][16 rows] with the four 16-byte row segments of a column XOR-swizzled by the column, so that float4 accesses of
// eight consecutive columns hit eight different bank groups
__device__ __forceinline__ int ts_swz(int col, int row) { return col * TS_ROWS + (row ^ (((col >> 1) & 3) << 2)); }

// first coupling stage of the chain: the stages before it hold no parameters, so the backward walk ends there
__host__ __device__ inline int ts_first_nsc(const PlanDev& hp) {
  for (int si = 0; si < hp.n_steps; ++si)
    if (hp.steps[si].kind == STEP_NSC) return si;
  return hp.n_steps;
}

bool train_step_supported(const PlanDev& hp) {
  if (hp.n_nsc < 1) return false;
  if (hp.n_lat != 1 || (hp.lat[0].kind != 1 && hp.lat[0].kind != 2 && hp.lat[0].kind != 3)) return false;
  // ColorTransform / InvSoftplus (the photo-z chains of the reference's tutorials put them in FRONT of the coupling
  // stages) run forward only: behind a coupling stage they would need a backward of their own
  const int first = ts_first_nsc(hp);
  for (int si = first; si < hp.n_steps; ++si) {
    const StepDev st = hp.steps[si];
    if (st.kind == STEP_COLOR) return false;
    if (st.kind == STEP_AFFINE && hp.affine[st.idx].kind == AFF_INV_SOFTPLUS) return false;
  }
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    if (st.n_dense < 1 || st.H > TS_HMAX || st.K > TS_KMAX || st.cpd * st.L > TS_OUT_MAX || st.U + st.C > TS_HMAX) return false;
  }
  return true;
}

// floats of scratch one ROW needs: per coupling stage the state before it, every hidden activation and the logits
long long train_row_scratch(const PlanDev& hp) {
  long long n = 0;
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    n += hp.D + (long long)(st.n_dense - 1) * st.H + (long long)st.cpd * st.L;
  }
  return n;
}

struct TsSmem {  // float offsets
  int xs, gx, cs, gld, ain, g0, g1, lg, wc, red, total;
};
__host__ __device__ inline TsSmem ts_layout(int D, int C) {
  TsSmem s;
  int o = 0;
  s.xs = o;  o += D * TS_ROWS;
  s.gx = o;  o += D * TS_ROWS;
  s.cs = o;  o += (C > 0 ? C : 1) * TS_ROWS;
  s.gld = o; o += 2 * TS_ROWS;                  // d loss / d log-det per row, and the running log-det
  s.ain = o; o += TS_HMAX * TS_ROWS;
  s.g0 = o;  o += TS_HMAX * TS_ROWS;
  s.g1 = o;  o += TS_HMAX * TS_ROWS;
  s.lg = o;  o += TS_OUT_MAX * TS_ROWS;
  s.wc = o;  o += 2 * (TS_HMAX * (TS_NC + 1) > TS_KC * 256 ? TS_HMAX * (TS_NC + 1) : TS_KC * 256);   // two chunk buffers
  s.red = o; o += 32;
  s.total = o;
  return s;
}

constexpr int TS_WC_HALF = TS_HMAX * (TS_NC + 1) > TS_KC * 256 ? TS_HMAX * (TS_NC + 1) : TS_KC * 256;

__device__ __forceinline__ void ts_cp4(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void ts_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_>
__device__ __forceinline__ void ts_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_)); }

// out[n][r] = act(b[n] + sum_k in[k][r] W[k][n]), n < N, k < K; W row-major [K][N] in global memory
template <bool LEAKY>
__device__ void ts_gemm_fwd(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ b,
                            int N, float* __restrict__ out, float* __restrict__ wc, int tid) {
  const int rg = tid & 3, nl = tid >> 2;
  for (int n0 = 0; n0 < N; n0 += 256) {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
    const int ncols = min(256, N - n0);
    // the weight chunks stream global -> shared through two buffers (cp.async): chunk i + 1 loads under chunk i's FMAs
    auto fetch = [&](int k0, float* buf) {
      const int kr = min(TS_KC, K - k0);
      for (int e = tid; e < kr * 256; e += TS_THREADS) {
        const int kk = e >> 8, c = e & 255;
        if (c < ncols) ts_cp4(buf + e, W + (size_t)(k0 + kk) * N + n0 + c);
        else buf[e] = 0.0f;
      }
      ts_cp_commit();
    };
    __syncthreads();   // the previous user of the chunk buffers is done
    fetch(0, wc);
    for (int k0 = 0, it = 0; k0 < K; k0 += TS_KC, ++it) {
      const int kr = min(TS_KC, K - k0);
      float* cur = wc + (it & 1) * TS_WC_HALF;
      if (k0 + TS_KC < K) {
        fetch(k0 + TS_KC, wc + ((it + 1) & 1) * TS_WC_HALF);
        ts_cp_wait<1>();
      } else {
        ts_cp_wait<0>();
      }
      __syncthreads();
#pragma unroll 4
      for (int kk = 0; kk < kr; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(in + ts_swz(k0 + kk, 4 * rg));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float w = cur[kk * 256 + nl + 64 * j];
          acc[j][0] = fmaf(a.x, w, acc[j][0]);
          acc[j][1] = fmaf(a.y, w, acc[j][1]);
          acc[j][2] = fmaf(a.z, w, acc[j][2]);
          acc[j][3] = fmaf(a.w, w, acc[j][3]);
        }
      }
      __syncthreads();   // everyone is done with `cur` before the fetch after next overwrites it
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + nl + 64 * j;
      if (n < N) {
        const float bb = b[n];
        float4 v = make_float4(acc[j][0] + bb, acc[j][1] + bb, acc[j][2] + bb, acc[j][3] + bb);
        if (LEAKY) {
          v.x = fmaxf(v.x, TS_SLOPE * v.x);
          v.y = fmaxf(v.y, TS_SLOPE * v.y);
          v.z = fmaxf(v.z, TS_SLOPE * v.z);
          v.w = fmaxf(v.w, TS_SLOPE * v.w);
        }
        *reinterpret_cast<float4*>(out + ts_swz(n, 4 * rg)) = v;
      }
    }
  }
  __syncthreads();
}

// gin[k][r] = scale(k, r) * sum_n W[k][n] gout[n][r], k < K (<= 256): the input gradient of a dense layer;
// act != nullptr: times LeakyReLU'(act[k][r]) (the activation the layer's input came out of)
__device__ void ts_gemm_dgrad(const float* __restrict__ gout, int N, const float* __restrict__ W, int K,
                              const float* __restrict__ act, float* __restrict__ gin, float* __restrict__ wc, int tid) {
  const int rg = tid & 3, kl = tid >> 2;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
  auto fetch = [&](int n0, float* buf) {
    const int nc = min(TS_NC, N - n0);
    for (int e = tid; e < K * TS_NC; e += TS_THREADS) {
      const int k = e / TS_NC, c = e % TS_NC;
      if (c < nc) ts_cp4(buf + k * (TS_NC + 1) + c, W + (size_t)k * N + n0 + c);
    }
    ts_cp_commit();
  };
  __syncthreads();
  fetch(0, wc);
  for (int n0 = 0, it = 0; n0 < N; n0 += TS_NC, ++it) {
    const int nc = min(TS_NC, N - n0);
    const float* cur = wc + (it & 1) * TS_WC_HALF;
    if (n0 + TS_NC < N) {
      fetch(n0 + TS_NC, wc + ((it + 1) & 1) * TS_WC_HALF);
      ts_cp_wait<1>();
    } else {
      ts_cp_wait<0>();
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
      const float4 g = *reinterpret_cast<const float4*>(gout + ts_swz(n0 + c, 4 * rg));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kl + 64 * j;
        if (k < K) {
          const float w = cur[k * (TS_NC + 1) + c];
          acc[j][0] = fmaf(g.x, w, acc[j][0]);
          acc[j][1] = fmaf(g.y, w, acc[j][1]);
          acc[j][2] = fmaf(g.z, w, acc[j][2]);
          acc[j][3] = fmaf(g.w, w, acc[j][3]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = kl + 64 * j;
    if (k < K) {
      float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      if (act != nullptr) {
        const float4 h = *reinterpret_cast<const float4*>(act + ts_swz(k, 4 * rg));
        v.x *= h.x >= 0.0f ? 1.0f : TS_SLOPE;
        v.y *= h.y >= 0.0f ? 1.0f : TS_SLOPE;
        v.z *= h.z >= 0.0f ? 1.0f : TS_SLOPE;
        v.w *= h.w >= 0.0f ? 1.0f : TS_SLOPE;
      }
      *reinterpret_cast<float4*>(gin + ts_swz(k, 4 * rg)) = v;
    }
  }
  __syncthreads();
}

// gW[k][n] = sum_r in[k][r] gout[n][r] and gb[n] = sum_r gout[n][r], written to this CTA's partial-gradient slice
__device__ void ts_gemm_wgrad(const float* __restrict__ in, int K, const float* __restrict__ gout, int N,
                              float* __restrict__ gW, float* __restrict__ gb, int tid) {
  const int nq_n = (N + 3) >> 2, kq_n = (K + 3) >> 2;
  for (int t = tid; t < nq_n * kq_n; t += TS_THREADS) {
    const int kq = t / nq_n, nq = t - kq * nq_n;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      float4 a[4], g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = min(4 * kq + i, K - 1), n = min(4 * nq + i, N - 1);
        a[i] = *reinterpret_cast<const float4*>(in + ts_swz(k, 4 * r4));
        g[i] = *reinterpret_cast<const float4*>(gout + ts_swz(n, 4 * r4));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(a[i].x, g[j].x, fmaf(a[i].y, g[j].y, fmaf(a[i].z, g[j].z, fmaf(a[i].w, g[j].w, acc[i][j]))));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * kq + i, n = 4 * nq + j;
        if (k < K && n < N) gW[(size_t)k * N + n] = acc[i][j];
      }
  }
  for (int n = tid; n < N; n += TS_THREADS) {
    float s = 0.0f;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      const float4 g = *reinterpret_cast<const float4*>(gout + ts_swz(n, 4 * r4));
      s += (g.x + g.y) + (g.z + g.w);
    }
    gb[n] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ float ts_softplus(float v) { return fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v))); }
__device__ __forceinline__ float ts_sigmoid(float v) { return 1.0f / (1.0f + expf(-v)); }

// The spline of one (row, dim) whose cpd logits sit in lg[(c0 + q)][row] (swizzled): bin, knots, sizes, derivatives.
// Same arithmetic as nsf_train.cu:spline_bin (pzflow/utils.py:113-161), forward direction.
struct TsBin {
  int idx, i0, i1;
  bool oob;
  float mx, xk, yk, wk, hk, d0, d1, sum_w, sum_h;
};
__device__ void ts_spline_bin(const float* __restrict__ lg, int c0, int row, int K, float B, float x, bool periodic, float* ew,
                              float* eh, TsBin& s) {
  float mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) {
    mw = fmaxf(mw, lg[ts_swz(c0 + j, row)]);
    mh = fmaxf(mh, lg[ts_swz(c0 + K + j, row)]);
  }
  s.sum_w = s.sum_h = 0.0f;
  for (int j = 0; j < K; ++j) {
    ew[j] = expf(lg[ts_swz(c0 + j, row)] - mw);
    eh[j] = expf(lg[ts_swz(c0 + K + j, row)] - mh);
    s.sum_w += ew[j];
    s.sum_h += eh[j];
  }
  const float twoB = 2.0f * B;
  s.oob = (x <= -B) || (x >= B);
  const float mx = s.oob ? (periodic ? fabsf(x) - B : 0.0f) : x;
  s.mx = mx;
  float run = 0.0f, left = 0.0f;
  int idx = 0;
  for (int j = 0; j < K - 1; ++j) {
    run += ew[j];
    if (-B + twoB * (run / s.sum_w) <= mx) {
      idx = j + 1;
      left = run;
    }
  }
  s.idx = idx;
  float lefth = 0.0f;
  for (int j = 0; j < idx; ++j) lefth += eh[j];
  s.xk = -B + twoB * (left / s.sum_w);
  s.yk = -B + twoB * (lefth / s.sum_h);
  s.wk = twoB * (ew[idx] / s.sum_w);
  s.hk = twoB * (eh[idx] / s.sum_h);
  if (periodic) {
    s.i0 = idx == 0 ? K - 1 : idx - 1;
    s.i1 = idx;
  } else {
    s.i0 = idx == 0 ? -1 : idx - 1;
    s.i1 = idx == K - 1 ? -1 : idx;
  }
  s.d0 = s.i0 < 0 ? 1.0f : ts_softplus(lg[ts_swz(c0 + 2 * K + s.i0, row)]);
  s.d1 = s.i1 < 0 ? 1.0f : ts_softplus(lg[ts_swz(c0 + 2 * K + s.i1, row)]);
}

__device__ __forceinline__ float ts_block_sum(float v, float* red, int tid) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  for (int w = 0; w < TS_THREADS / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(TS_THREADS, 1)
train_step_kernel(const PlanDev* __restrict__ plan, const TrainOffsets* __restrict__ offs, const float* __restrict__ params,
                  const float* __restrict__ X, const float* __restrict__ cond, const float* __restrict__ wrow, long long N,
                  float inv_n, long long row_scratch, float* __restrict__ scratch, long long n_params,
                  float* __restrict__ partials, float* __restrict__ loss_part, TrainBatch tb) {
  extern __shared__ __align__(16) float ts_smem[];
  const int D = plan->D, C = plan->C, ns = plan->n_steps;
  const TsSmem sl = ts_layout(D, C);
  float* xs = ts_smem + sl.xs;
  float* gx = ts_smem + sl.gx;
  float* cs = ts_smem + sl.cs;
  float* gld = ts_smem + sl.gld;
  float* ldsum = gld + TS_ROWS;
  float* ain = ts_smem + sl.ain;
  float* g0 = ts_smem + sl.g0;
  float* g1 = ts_smem + sl.g1;
  float* lg = ts_smem + sl.lg;
  float* wc = ts_smem + sl.wc;
  float* red = ts_smem + sl.red;
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * TS_ROWS;
  float* my_scratch = scratch + (size_t)blockIdx.x * TS_ROWS * row_scratch;   // this tile's [field][col][16 rows] blocks
  float* my_part = partials + (size_t)blockIdx.x * n_params;

  // (every parameter of the flat buffer belongs to some dense layer of the chain -- the host checks it -- so this
  // CTA's slice of the partial-gradient buffer is written completely below and needs no zeroing)

  // ---- load the tile (row-major global -> [col][row] shared): the batch's rows are gathered through the epoch's
  // permutation here, and the Gaussian error samples of convolve_errs (pzflow/flow.py:1068-1079, utils.py:52-61)
  // are drawn here, so a step launches nothing but this kernel, the partial-sum reduction and Adam ----
  for (int e = tid; e < D * TS_ROWS; e += TS_THREADS) {
    const int r = e / D, j = e - r * D;
    // the padding rows of a ragged last tile repeat the batch's last row (weight 0 below): zeros are not inside every
    // chain's domain (InvSoftplus(0) = -inf), and a non-finite activation would poison the tile's weight gradient
    const long long br = row0 + r < N ? row0 + r : N - 1;
    const long long src = tb.row_idx != nullptr ? tb.row_idx[br] : br;
    float v = X[src * D + j];
    if (tb.Xerr != nullptr) v = v + pz_err_normal(tb.seed, 0u, tb.noise_base + (unsigned long long)br, j) * tb.Xerr[src * D + j];
    xs[ts_swz(j, r)] = v;
  }
  for (int e = tid; e < C * TS_ROWS; e += TS_THREADS) {
    const int r = e / C, j = e - r * C;
    const long long br = row0 + r < N ? row0 + r : N - 1;
    const long long src = tb.row_idx != nullptr ? tb.row_idx[br] : br;
    float v = cond[src * C + j];
    if (tb.cond_err != nullptr)
      v = v + pz_err_normal(tb.seed, 1u, tb.noise_base + (unsigned long long)br, j) * tb.cond_err[src * C + j];
    cs[ts_swz(j, r)] = v;
  }
  if (tid < TS_ROWS) ldsum[tid] = 0.0f;
  __syncthreads();

  // ================================ forward ================================
  long long so = 0;  // float offset of the current stage's block inside this tile's scratch (per-row stride applied below)
  for (int si = 0; si < ns; ++si) {
    const StepDev step = plan->steps[si];
    if (step.kind == STEP_AFFINE) {
      const AffineDev& af = plan->affine[step.idx];
      const bool has_ld = af.kind == AFF_INV_SOFTPLUS;   // the one elementwise stage whose log-det depends on the row
      for (int e = tid; e < D * TS_ROWS; e += TS_THREADS) {
        const int j = e / TS_ROWS, r = e - j * TS_ROWS;
        float ld;
        xs[ts_swz(j, r)] = pz_affine_elem(af, false, j, xs[ts_swz(j, r)], ld);
        if (has_ld) gx[ts_swz(j, r)] = ld;               // (the gradient buffer is idle during the forward walk)
      }
      if (has_ld) {
        __syncthreads();
        if (tid < TS_ROWS) {                             // summed in column order: deterministic
          float sld = 0.0f;
          for (int j = 0; j < D; ++j) sld += gx[ts_swz(j, tid)];
          ldsum[tid] += sld;
        }
      } else if (tid < TS_ROWS) {
        ldsum[tid] += af.logdet_fwd;
      }
      __syncthreads();
    } else if (step.kind == STEP_COLOR) {
      // ColorTransform forward on one row per thread (bijectors.py:268-280: colours = -diff(mags), the reference
      // magnitude last; log-det 0), same column slots as pz_color_row
      if (tid < TS_ROWS) {
        const ColorDev& cd = plan->color[step.idx];
        const int r = tid, n = cd.n_mag;
        float prev = xs[ts_swz(cd.slot[0], r)];
        const float refv = xs[ts_swz(cd.slot[cd.ref], r)];
        for (int i = 0; i + 1 < n; ++i) {
          const float nxt = xs[ts_swz(cd.slot[i + 1], r)];
          xs[ts_swz(cd.slot[i], r)] = -(nxt - prev);
          prev = nxt;
        }
        xs[ts_swz(cd.slot[n - 1], r)] = refv;
      }
      __syncthreads();
    } else if (step.kind == STEP_NSC) {
      const NscDev& st = plan->nsc[step.idx];
      const TrainOffsets& of = offs[step.idx];
      const int U = st.U, L = st.L, in_dim = st.U + st.C, out_dim = st.cpd * st.L, H = st.H;
      float* sx = my_scratch + so * TS_ROWS;                      // [D][16] state before the stage
      float* sh = sx + D * TS_ROWS;                                // [n_dense - 1][H][16] hidden activations
      float* slg = sh + (size_t)(st.n_dense - 1) * H * TS_ROWS;    // [out_dim][16] logits
      for (int e = tid; e < D * TS_ROWS; e += TS_THREADS) sx[e] = xs[e];
      for (int e = tid; e < in_dim * TS_ROWS; e += TS_THREADS) {
        const int k = e / TS_ROWS, r = e - k * TS_ROWS;
        ain[ts_swz(k, r)] = k < U ? xs[ts_swz(st.perm[k], r)] : cs[ts_swz(k - U, r)];
      }
      __syncthreads();
      const float* cur = ain;
      int cur_dim = in_dim;
      for (int li = 0; li + 1 < st.n_dense; ++li) {
        float* nxt = (li & 1) ? g1 : g0;
        ts_gemm_fwd<true>(cur, cur_dim, params + of.w[li], params + of.b[li], H, nxt, wc, tid);
        for (int e = tid; e < H * TS_ROWS; e += TS_THREADS) sh[(size_t)li * H * TS_ROWS + e] = nxt[e];
        cur = nxt;
        cur_dim = H;
      }
      const int last = st.n_dense - 1;
      ts_gemm_fwd<false>(cur, cur_dim, params + of.w[last], params + of.b[last], out_dim, lg, wc, tid);
      for (int e = tid; e < out_dim * TS_ROWS; e += TS_THREADS) slg[e] = lg[e];
      // spline forward, one thread per (row, dim)
      for (int p = tid; p < L * TS_ROWS; p += TS_THREADS) {
        const int d = p / TS_ROWS, r = p - d * TS_ROWS;
        const int col = st.perm[U + d];
        const float xv = xs[ts_swz(col, r)];
        float ew[TS_KMAX], eh[TS_KMAX];
        TsBin b;
        ts_spline_bin(lg, d * st.cpd, r, st.K, st.B, xv, st.periodic != 0, ew, eh, b);
        const bool pass = b.oob && !st.periodic;
        const float sk = b.hk / b.wk;
        const float xi = (b.mx - b.xk) / b.wk, o = 1.0f - xi, t = xi * o;
        const float q = b.d1 + b.d0 - 2.0f * sk;
        const float num = b.hk * (sk * xi * xi + b.d0 * t), den = sk + q * t;
        const float dn = b.d1 * xi * xi + 2.0f * sk * t + b.d0 * o * o;
        xs[ts_swz(col, r)] = pass ? xv : b.yk + num / den;
        g0[d * TS_ROWS + r] = pass ? 0.0f : 2.0f * logf(sk) + logf(dn) - 2.0f * logf(den);   // (the hidden buffers are idle)
      }
      __syncthreads();
      if (tid < TS_ROWS) {   // log_det.sum(axis=1) in column order: deterministic
        float sld = 0.0f;
        for (int d = 0; d < L; ++d) sld += g0[d * TS_ROWS + tid];
        ldsum[tid] += sld;
      }
      __syncthreads();
      so += D + (long long)(st.n_dense - 1) * H + out_dim;
    }
    // chain brackets only fix the summation order of the log-determinants (one running sum here); Roll / Reverse are
    // folded into the stages' column maps
  }

  // ---- latent log-density, loss, d loss / d u ----
  float my_loss = 0.0f;
  if (tid < TS_ROWS) {
    const int r = tid;
    const LatentSeg sg = plan->lat[0];
    const float Bl = sg.B;
    float lat = 0.0f;
    bool inside = true;
    for (int j = 0; j < D; ++j) {
      const float u = xs[ts_swz(plan->perm_final[j], r)];
      if (sg.kind == 1) inside = inside && u >= -Bl && u <= Bl;
      else if (sg.kind == 3) lat += u * u;
      else {
        const float y = (u + Bl) / (2.0f * Bl);
        lat += (12.0f * logf(y) + 12.0f * log1pf(-y)) - sg.c0 - logf(2.0f * Bl);
        if (u > Bl || u < -Bl) inside = false;
      }
    }
    if (sg.kind == 1) lat = inside ? sg.c0 : -INFINITY;
    else if (sg.kind == 3) lat = -0.5f * ((float)D * 1.8378770664093453f + lat);
    else if (!inside) lat = -INFINITY;
    const float lp = lat + ldsum[r];
    const bool ok = row0 + r < N && isfinite(lp);                 // rows outside the support carry no gradient
    const float wv = ok ? wrow[tb.row_idx != nullptr ? tb.row_idx[row0 + r] : row0 + r] : 0.0f;
    my_loss = ok ? -(wv * lp) * inv_n : 0.0f;
    const float glp = -wv * inv_n;
    gld[r] = glp;
    for (int j = 0; j < D; ++j) {
      const int col = plan->perm_final[j];
      const float u = xs[ts_swz(col, r)];
      float du = 0.0f;
      if (sg.kind == 3) du = -u;
      else if (sg.kind == 2) du = 12.0f / (u + Bl) - 12.0f / (Bl - u);
      gx[ts_swz(col, r)] = ok ? glp * du : 0.0f;
    }
  }
  const float tile_loss = ts_block_sum(my_loss, red, tid);
  if (tid == 0) loss_part[blockIdx.x] = tile_loss;

  // ================================ backward ================================
  const int first_nsc = ts_first_nsc(*plan);   // nothing in front of it has parameters: its input gradient is not needed
  for (int si = ns - 1; si >= first_nsc; --si) {
    const StepDev step = plan->steps[si];
    if (step.kind == STEP_AFFINE) {
      const AffineDev& af = plan->affine[step.idx];
      for (int e = tid; e < D * TS_ROWS; e += TS_THREADS) {
        const int j = e / TS_ROWS, r = e - j * TS_ROWS;
        const float s = af.kind == AFF_SHIFT_BOUNDS ? af.B / af.b[j]
                        : (af.kind == AFF_STANDARD_SCALER ? 1.0f / af.b[j] : (af.kind == AFF_DEQUANTIZE ? 1.0f : af.B));
        gx[ts_swz(j, r)] *= s;
      }
      __syncthreads();
    } else if (step.kind == STEP_NSC) {
      const NscDev& st = plan->nsc[step.idx];
      const TrainOffsets& of = offs[step.idx];
      const int U = st.U, L = st.L, in_dim = st.U + st.C, out_dim = st.cpd * st.L, H = st.H, K = st.K;
      so -= D + (long long)(st.n_dense - 1) * H + out_dim;
      const float* sx = my_scratch + so * TS_ROWS;
      const float* sh = sx + D * TS_ROWS;
      const float* slg = sh + (size_t)(st.n_dense - 1) * H * TS_ROWS;
      for (int e = tid; e < D * TS_ROWS; e += TS_THREADS) xs[e] = sx[e];
      for (int e = tid; e < out_dim * TS_ROWS; e += TS_THREADS) lg[e] = slg[e];
      __syncthreads();
      // spline backward: logits -> their gradient in place, g_y -> g_x of the transformed column
      for (int p = tid; p < L * TS_ROWS; p += TS_THREADS) {
        const int d = p / TS_ROWS, r = p - d * TS_ROWS;
        const int col = st.perm[U + d], c0 = d * st.cpd;
        const float xv = xs[ts_swz(col, r)];
        float ew[TS_KMAX], eh[TS_KMAX];
        TsBin s;
        ts_spline_bin(lg, c0, r, K, st.B, xv, st.periodic != 0, ew, eh, s);
        const float g_y = gx[ts_swz(col, r)], g_l = gld[r];
        if (s.oob && !st.periodic) {
          for (int j = 0; j < st.cpd; ++j) lg[ts_swz(c0 + j, r)] = 0.0f;
          continue;                                             // identity: g_x = g_y
        }
        const float twoB = 2.0f * st.B;
        const float sk = s.hk / s.wk;
        const float xi = (s.mx - s.xk) / s.wk, o = 1.0f - xi, t = xi * o;
        const float q = s.d1 + s.d0 - 2.0f * sk;
        const float inner = sk * xi * xi + s.d0 * t;
        const float num = s.hk * inner, den = sk + q * t, f = num / den;
        const float dn = s.d1 * xi * xi + 2.0f * sk * t + s.d0 * o * o;
        const float om2 = 1.0f - 2.0f * xi;
        const float f_s = (s.hk * xi * xi - f * (1.0f - 2.0f * t)) / den;
        const float f_d0 = (s.hk * t - f * t) / den;
        const float f_d1 = (-f * t) / den;
        const float f_xi = (s.hk * (2.0f * sk * xi + s.d0 * om2) - f * q * om2) / den;
        const float f_h = inner / den;
        const float l_s = 2.0f / sk + (2.0f * t) / dn - 2.0f * (1.0f - 2.0f * t) / den;
        const float l_d0 = (o * o) / dn - 2.0f * t / den;
        const float l_d1 = (xi * xi) / dn - 2.0f * t / den;
        const float l_xi = (2.0f * s.d1 * xi + 2.0f * sk * om2 - 2.0f * s.d0 * o) / dn - 2.0f * q * om2 / den;
        const float G_s = g_y * f_s + g_l * l_s, G_d0 = g_y * f_d0 + g_l * l_d0, G_d1 = g_y * f_d1 + g_l * l_d1;
        const float G_xi = g_y * f_xi + g_l * l_xi;
        const float g_hk = g_y * f_h + G_s / s.wk;
        const float g_wk = -(G_s * sk + G_xi * xi) / s.wk;
        const float g_xk = -G_xi / s.wk, g_yk = g_y;
        gx[ts_swz(col, r)] = (s.oob && xv < 0.0f ? -1.0f : 1.0f) * G_xi / s.wk;
        const float r0 = s.i0 >= 0 ? lg[ts_swz(c0 + 2 * K + s.i0, r)] : 0.0f;   // derivative logits, before they are overwritten
        const float r1 = s.i1 >= 0 ? lg[ts_swz(c0 + 2 * K + s.i1, r)] : 0.0f;
        const int idx = s.idx;
        float dotw = 0.0f, doth = 0.0f;
        for (int j = 0; j <= idx; ++j) {
          const float Wj = twoB * (ew[j] / s.sum_w), Hj = twoB * (eh[j] / s.sum_h);
          dotw += (j < idx ? g_xk : g_wk) * Wj;
          doth += (j < idx ? g_yk : g_hk) * Hj;
        }
        dotw /= twoB;
        doth /= twoB;
        for (int j = 0; j < K; ++j) {
          const float Wj = twoB * (ew[j] / s.sum_w), Hj = twoB * (eh[j] / s.sum_h);
          const float gW = j < idx ? g_xk : (j == idx ? g_wk : 0.0f), gH = j < idx ? g_yk : (j == idx ? g_hk : 0.0f);
          lg[ts_swz(c0 + j, r)] = Wj * (gW - dotw);
          lg[ts_swz(c0 + K + j, r)] = Hj * (gH - doth);
        }
        for (int j = 2 * K; j < st.cpd; ++j) lg[ts_swz(c0 + j, r)] = 0.0f;
        if (s.i0 >= 0) lg[ts_swz(c0 + 2 * K + s.i0, r)] += G_d0 * ts_sigmoid(r0);
        if (s.i1 >= 0) lg[ts_swz(c0 + 2 * K + s.i1, r)] += G_d1 * ts_sigmoid(r1);
      }
      __syncthreads();
      // dense layers, last to first: weight / bias gradients, then the gradient of the layer's input
      const float* gout = lg;
      int gdim = out_dim;
      for (int li = st.n_dense - 1; li >= 0; --li) {
        const int kdim = li == 0 ? in_dim : H;
        if (li == 0) {
          for (int e = tid; e < in_dim * TS_ROWS; e += TS_THREADS) {
            const int k = e / TS_ROWS, r = e - k * TS_ROWS;
            ain[ts_swz(k, r)] = k < U ? xs[ts_swz(st.perm[k], r)] : cs[ts_swz(k - U, r)];
          }
        } else {
          for (int e = tid; e < H * TS_ROWS; e += TS_THREADS) ain[e] = sh[(size_t)(li - 1) * H * TS_ROWS + e];
        }
        __syncthreads();
        ts_gemm_wgrad(ain, kdim, gout, gdim, my_part + of.w[li], my_part + of.b[li], tid);
        float* gin = (gout == g0) ? g1 : g0;
        ts_gemm_dgrad(gout, gdim, params + of.w[li], kdim, li == 0 ? nullptr : ain, gin, wc, tid);
        gout = gin;
        gdim = kdim;
      }
      // gout = d loss / d key: the conditioner's inputs are the untouched columns (conditions take no gradient)
      for (int e = tid; e < U * TS_ROWS; e += TS_THREADS) {
        const int k = e / TS_ROWS, r = e - k * TS_ROWS;
        gx[ts_swz(st.perm[k], r)] += gout[ts_swz(k, r)];
      }
      __syncthreads();
    }
  }
}

// flat_grad[i] = sum over the CTAs' partial gradients (in CTA order: deterministic), non-finite entries -> 0;
// loss[0] = sum of the tiles' losses
__global__ void __launch_bounds__(256)
train_reduce_kernel(const float* __restrict__ partials, int n_cta, long long n_params, float* __restrict__ flat_grad,
                    const float* __restrict__ loss_part, float* __restrict__ loss) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_params; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * n_params + i];
    flat_grad[i] = isfinite(s) ? s : 0.0f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss != nullptr) {
    float s = 0.0f;
    for (int c = 0; c < n_cta; ++c) s += loss_part[c];
    loss[0] = s;
  }
}

size_t train_step_workspace_floats(const PlanDev& hp, long long N, long long n_params) {
  const long long n_cta = (N + TS_ROWS - 1) / TS_ROWS;
  return (size_t)(n_cta * TS_ROWS * train_row_scratch(hp) + n_cta * n_params + n_cta);
}

cudaError_t launch_train_step(const PlanDev* d_plan, const PlanDev& hp, const TrainOffsets* d_offs, const float* params,
                              const float* X, const float* cond, const float* w, long long N, float inv_n, long long n_params,
                              float* workspace, float* flat_grad, float* loss, int num_sms, const TrainBatch& tb,
                              cudaStream_t stream) {
  const long long n_cta = (N + TS_ROWS - 1) / TS_ROWS;
  if (n_cta == 0) {
    cudaError_t e = cudaMemsetAsync(flat_grad, 0, (size_t)n_params * sizeof(float), stream);
    if (e == cudaSuccess && loss != nullptr) e = cudaMemsetAsync(loss, 0, sizeof(float), stream);
    return e;
  }
  const long long rs = train_row_scratch(hp);
  float* scratch = workspace;
  float* partials = scratch + (size_t)n_cta * TS_ROWS * rs;
  float* loss_part = partials + (size_t)n_cta * n_params;
  const TsSmem sl = ts_layout(hp.D, hp.C);
  const size_t smem = (size_t)sl.total * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(train_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  train_step_kernel<<<(unsigned)n_cta, TS_THREADS, smem, stream>>>(d_plan, d_offs, params, X, cond, w, N, inv_n, rs, scratch,
                                                                   n_params, partials, loss_part, tb);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  long long blocks = (n_params + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  train_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partials, (int)n_cta, n_params, flat_grad, loss_part, loss);
  return cudaGetLastError();
}

}  // namespace pzb
