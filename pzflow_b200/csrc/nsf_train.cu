// Training support: the coupling stage's spline -- softmax / softplus knot construction, bin
// search, rational-quadratic spline and log-determinant (pzflow/bijectors.py:488-506, pzflow/utils.py:64-227) --
// as ONE forward kernel and ONE hand-derived backward kernel, one thread per (row, spline dimension).  The
// conditioner's dense layers around it are plain GEMMs and stay with cuBLAS (torch.matmul) in pzflow_b200/train.py.
//
// Backward (in-bounds; non-periodic out-of-bounds elements pass through, d out / d x = 1, no other gradient;
// periodic ones are evaluated at |x| - B and the derivative logits wrap around, utils.py:127-146):
//   s = h/w, xi = (x - xk)/w, o = 1 - xi, t = xi o, q = d1 + d0 - 2 s
//   num = h (s xi^2 + d0 t), den = s + q t, y = yk + num/den
//   dn  = d1 xi^2 + 2 s t + d0 o^2,  ld = 2 log s + log dn - 2 log den
// and the chain through xk = -B + sum_{j<idx} W_j, w = W_idx (same for y/H), W = 2B softmax(lw), d = softplus(ld).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pzb {

constexpr int TRAIN_KMAX = 64;

struct SplineBin {
  int idx, i0, i1;  // bin; derivative-logit index of its left / right knot (-1: fixed derivative 1)
  bool oob;
  float mx;         // the point the spline is evaluated at
  float xk, yk, wk, hk, d0, d1;   // left knot, bin sizes, derivatives at the bin's two knots
  float sum_w, sum_h, max_w, max_h;
};

// softmax terms of one axis into e[] (un-normalised), returns (max, sum)
__device__ __forceinline__ void axis_terms(const float* __restrict__ lg, int K, float* e, float& mx, float& sum) {
  mx = lg[0];
  for (int j = 1; j < K; ++j) mx = fmaxf(mx, lg[j]);
  sum = 0.0f;
  for (int j = 0; j < K; ++j) {
    e[j] = expf(lg[j] - mx);
    sum += e[j];
  }
}

__device__ __forceinline__ float softplus_f(float v) { return fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v))); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// logits of one (row, dim): [W logits K | H logits K | derivative logits K-1 (K if periodic)]; forward direction only
__device__ __forceinline__ void spline_bin(const float* __restrict__ lg, int K, float B, float x, bool periodic,
                                           float* ew, float* eh, SplineBin& s) {
  axis_terms(lg, K, ew, s.max_w, s.sum_w);
  axis_terms(lg + K, K, eh, s.max_h, s.sum_h);
  const float twoB = 2.0f * B;
  s.oob = (x <= -B) || (x >= B);
  // non-periodic: out-of-bounds elements evaluate a harmless point, their result is discarded (utils.py:208-209);
  // periodic: they are evaluated at |x| - B (utils.py:146) and kept
  const float mx = s.oob ? (periodic ? fabsf(x) - B : 0.0f) : x;
  s.mx = mx;
  // idx = #(knots <= x) - 1 over xk_0 = -B, xk_j = -B + cumsum(W)_j (pzflow/utils.py:150-152), clamped to [0, K-1]
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
  if (periodic) {  // K derivative logits, knot 0 and knot K share the last one (utils.py:127-128)
    s.i0 = idx == 0 ? K - 1 : idx - 1;
    s.i1 = idx;
  } else {         // K - 1 interior derivatives, 1 at both ends (utils.py:130-135)
    s.i0 = idx == 0 ? -1 : idx - 1;
    s.i1 = idx == K - 1 ? -1 : idx;
  }
  s.d0 = s.i0 < 0 ? 1.0f : softplus_f(lg[2 * K + s.i0]);
  s.d1 = s.i1 < 0 ? 1.0f : softplus_f(lg[2 * K + s.i1]);
}

__global__ void __launch_bounds__(128)
spline_train_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ x, long long n_elems, int K, float B,
                        int periodic, float* __restrict__ y, float* __restrict__ ld) {
  const int cpd = 3 * K - 1 + (periodic ? 1 : 0);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_elems; e += (long long)gridDim.x * blockDim.x) {
    float ew[TRAIN_KMAX], eh[TRAIN_KMAX];
    SplineBin s;
    const float xv = x[e];
    spline_bin(logits + e * cpd, K, B, xv, periodic != 0, ew, eh, s);
    const bool pass = s.oob && !periodic;
    const float sk = s.hk / s.wk;
    const float xi = (s.mx - s.xk) / s.wk, o = 1.0f - xi, t = xi * o;
    const float q = s.d1 + s.d0 - 2.0f * sk;
    const float num = s.hk * (sk * xi * xi + s.d0 * t), den = sk + q * t;
    const float dn = s.d1 * xi * xi + 2.0f * sk * t + s.d0 * o * o;
    y[e] = pass ? xv : s.yk + num / den;
    ld[e] = pass ? 0.0f : 2.0f * logf(sk) + logf(dn) - 2.0f * logf(den);
  }
}

__global__ void __launch_bounds__(128)
spline_train_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ x, const float* __restrict__ gy,
                        const float* __restrict__ gld, long long n_elems, int K, float B, int periodic,
                        float* __restrict__ glogits, float* __restrict__ gx) {
  const int cpd = 3 * K - 1 + (periodic ? 1 : 0);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_elems; e += (long long)gridDim.x * blockDim.x) {
    float ew[TRAIN_KMAX], eh[TRAIN_KMAX];
    SplineBin s;
    const float xv = x[e];
    const float* lg = logits + e * cpd;
    float* gl = glogits + e * cpd;
    spline_bin(lg, K, B, xv, periodic != 0, ew, eh, s);
    const float g_y = gy[e], g_l = gld[e];
    if (s.oob && !periodic) {
      for (int j = 0; j < cpd; ++j) gl[j] = 0.0f;
      gx[e] = g_y;
      continue;
    }
    const float twoB = 2.0f * B;
    const float sk = s.hk / s.wk;
    const float xi = (s.mx - s.xk) / s.wk, o = 1.0f - xi, t = xi * o;
    const float q = s.d1 + s.d0 - 2.0f * sk;
    const float inner = sk * xi * xi + s.d0 * t;
    const float num = s.hk * inner, den = sk + q * t, f = num / den;
    const float dn = s.d1 * xi * xi + 2.0f * sk * t + s.d0 * o * o;
    const float om2 = 1.0f - 2.0f * xi;
    // partials of f = num / den and of ld
    const float f_s = (s.hk * xi * xi - f * (1.0f - 2.0f * t)) / den;
    const float f_d0 = (s.hk * t - f * t) / den;
    const float f_d1 = (-f * t) / den;
    const float f_xi = (s.hk * (2.0f * sk * xi + s.d0 * om2) - f * q * om2) / den;
    const float f_h = inner / den;  // explicit h (s held fixed)
    const float l_s = 2.0f / sk + (2.0f * t) / dn - 2.0f * (1.0f - 2.0f * t) / den;
    const float l_d0 = (o * o) / dn - 2.0f * t / den;
    const float l_d1 = (xi * xi) / dn - 2.0f * t / den;
    const float l_xi = (2.0f * s.d1 * xi + 2.0f * sk * om2 - 2.0f * s.d0 * o) / dn - 2.0f * q * om2 / den;
    const float G_s = g_y * f_s + g_l * l_s, G_d0 = g_y * f_d0 + g_l * l_d0, G_d1 = g_y * f_d1 + g_l * l_d1;
    const float G_xi = g_y * f_xi + g_l * l_xi;
    const float g_hk = g_y * f_h + G_s / s.wk;
    const float g_wk = -(G_s * sk + G_xi * xi) / s.wk;
    const float g_xk = -G_xi / s.wk, g_yk = g_y;
    gx[e] = (s.oob && xv < 0.0f ? -1.0f : 1.0f) * G_xi / s.wk;  // d(|x| - B)/dx for wrapped periodic inputs
    // W_j = 2B softmax(lw)_j: g_lw_m = W_m (g_W_m - sum_j g_W_j W_j / 2B), g_W_j = g_xk (j < idx), g_wk (j == idx)
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
      gl[j] = Wj * (gW - dotw);
      gl[K + j] = Hj * (gH - doth);
    }
    for (int j = 2 * K; j < cpd; ++j) gl[j] = 0.0f;
    if (s.i0 >= 0) gl[2 * K + s.i0] += G_d0 * sigmoid_f(lg[2 * K + s.i0]);
    if (s.i1 >= 0) gl[2 * K + s.i1] += G_d1 * sigmoid_f(lg[2 * K + s.i1]);  // K = 1 periodic: i0 == i1
  }
}

// Adam (optax.adam as pzflow/flow.py:968-981 uses it: b1 0.9, b2 0.999, eps 1e-8, eps_root 0) over ONE flat parameter
// buffer.  The step count lives on the device so that the launch can sit inside a CUDA graph.
__global__ void adam_tick_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step) {
  const float t = (float)*step;
  const float c1 = 1.0f - powf(b1, t), c2 = 1.0f - powf(b2, t);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi / c1) / (sqrtf(vi / c2) + eps);
  }
}

// SM count of the current device (grid-stride kernels size their grids from it)
static int train_num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

cudaError_t launch_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, int* step, cudaStream_t stream) {
  adam_tick_kernel<<<1, 1, 0, stream>>>(step);
  if (n > 0) {
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)train_num_sms() * 8) blocks = (long long)train_num_sms() * 8;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, step);
  }
  return cudaGetLastError();
}

cudaError_t launch_spline_train_fwd(const float* logits, const float* x, long long n_elems, int K, float B, int periodic,
                                    float* y, float* ld, cudaStream_t stream) {
  if (n_elems == 0) return cudaSuccess;
  long long blocks = (n_elems + 127) / 128;
  if (blocks > (long long)train_num_sms() * 16) blocks = (long long)train_num_sms() * 16;
  spline_train_fwd_kernel<<<(unsigned)blocks, 128, 0, stream>>>(logits, x, n_elems, K, B, periodic, y, ld);
  return cudaGetLastError();
}

cudaError_t launch_spline_train_bwd(const float* logits, const float* x, const float* gy, const float* gld,
                                    long long n_elems, int K, float B, int periodic, float* glogits, float* gx,
                                    cudaStream_t stream) {
  if (n_elems == 0) return cudaSuccess;
  long long blocks = (n_elems + 127) / 128;
  if (blocks > (long long)train_num_sms() * 16) blocks = (long long)train_num_sms() * 16;
  spline_train_bwd_kernel<<<(unsigned)blocks, 128, 0, stream>>>(logits, x, gy, gld, n_elems, K, B, periodic, glogits, gx);
  return cudaGetLastError();
}

}  // namespace pzb
