// tcgen05 engine for WIDE / DEEP conditioners: hidden_dim <= 256, K <= 32 bins, any number (>= 1) of hidden layers,
// up to 63 conditioner inputs, any number of transformed columns -- BASELINE config 5 (32-D, K = 32, hidden 256) and
// every flow the default-family engine (nsf_tc.cu) turns away for K > 16, hidden_dim > 128 or hidden_layers != 2.
//
// Same arithmetic as nsf_tc.cu: every fp32 operand is split x = hi + lo (fp16 each) and a GEMM is three kind::f16
// MMAs, lo*hi + hi*lo + hi*hi, accumulated in fp32 in TMEM; activations never touch shared memory (the epilogue
// rewrites the fp32 accumulator columns IN PLACE as fp16 hi/lo A operands, tcgen05.mma reads A from TMEM).
//
// What differs is where the weights live.  One coupling stage of config 5 has 1.8 MB of split weights (W2 256 KB,
// W3 1.5 MB), so nothing is resident: EVERY weight matrix is cut into slabs of [<= 128 output rows x 64 K-elements]
// (hi + lo, K-major SWIZZLE_128B; the first layer: [128 x K1] SWIZZLE_NONE) that stream L2 -> shared memory through one
// TMA ring in exactly the order the MMA issuer consumes them, across layer, stage and tile boundaries.  A slab of the
// output layer (96 logit columns = one K <= 32 dimension or two K <= 16 dimensions, 24 KB) feeds 12 MMAs = 576 tensor
// clocks, i.e. the ring has to deliver ~43 B/clk/SM.
//
// One persistent CTA per SM, 320 threads, ONE 128-row tile in flight (a tile's A operand of 256 K-elements is 256 TMEM
// columns; with the accumulators that is the whole 512):
//   * 8 epilogue warps, two per TMEM lane quarter ("halves"): half h converts hidden columns [h HP/2, (h+1) HP/2) and
//     evaluates the output passes p = h, h + 2, ... (accumulator buffer h), so the spline of pass p runs under the
//     MMAs of pass p + 1;
//   * warp 8: MMA issuer (whole warp waits, one elected lane issues: operands stay on the uniform datapath);
//   * warp 9: TMA loader (parameter block of a stage, double buffered, then its slabs).
//   TMEM: region P = columns [0, 256), Q = [256, 512).  keys -> Q[0, K1); layer 1 -> P; layer 2 -> Q; ...; the output
//   passes land in the region the last hidden layer does NOT occupy, two buffers at +0 and +128.
//
// Follows the same reference lines as nsf_simt.cu / nsf_tc.cu (pzflow/bijectors.py:462-520, pzflow/utils.py:22-49, 64-227).
#include "pzb_tc_common.cuh"

namespace pzb {

constexpr int WD_TILE = 128;
constexpr int WD_EPI_THREADS = 256;
constexpr int WD_THREADS = 320;
constexpr int WD_SLOT_BYTES = 32768;
constexpr int WD_MAX_SLOTS = 7;
constexpr int WD_HP_MAX = 256;
constexpr int WD_IN_MAX = 63;              // conditioner inputs: K1 <= 64 minus the bias row
constexpr int WD_PASS_N = 96;              // logit columns of one output pass
constexpr int WD_HID_SLAB = 2 * 128 * 64 * 2;         // [128 x 64] hi + lo
constexpr int WD_OUT_SLAB = 2 * WD_PASS_N * 64 * 2;   // [96 x 64] hi + lo
constexpr int WD_SMEM_MAX = 232448;
// parameter block of a stage (floats / ints, see wd_pack_stage)
constexpr int WD_PI_HL = 0, WD_PI_HP = 1, WD_PI_K1 = 2, WD_PI_KP = 3, WD_PI_L = 4, WD_PI_U = 5, WD_PI_IN = 6,
              WD_PI_PERIODIC = 7, WD_PI_BINS = 8, WD_PI_K = 9, WD_PI_NPASS = 10, WD_PI_OFF_BH = 11, WD_PI_OFF_B3 = 12,
              WD_PI_OFF_KEY = 13, WD_PI_OFF_LOW = 14;
constexpr int WD_PF_S3INV = 16, WD_PF_B = 17, WD_PF_INV2B = 18, WD_PF_SINV = 20;  // sinv of hidden layer j >= 2 at 20 + j - 2
constexpr int WD_PAR_HEAD = 32;

struct WdGeom {
  int in_dim, K1, HP, hl, KP, NPd, L, n_pass, par_floats, par_bytes, w1_slab;
  int off_bh, off_b3, off_key, off_low;  // float offsets inside the parameter block
  size_t off_w1, off_hid, off_w3, total; // byte offsets inside the stage blob
  int n_slabs;
};

__host__ __device__ inline WdGeom wd_geom(int U, int C, int H, int n_dense, int K, int L) {
  WdGeom g;
  g.in_dim = U + C;
  g.K1 = (g.in_dim + 1 + 15) / 16 * 16;
  g.HP = H <= 128 ? 128 : 256;
  g.hl = n_dense - 1;
  g.KP = K <= 16 ? 16 : 32;
  g.NPd = 3 * g.KP;
  g.L = L;
  g.n_pass = g.KP == 32 ? L : (L + 1) / 2;
  g.off_bh = WD_PAR_HEAD;
  g.off_b3 = g.off_bh + (g.hl - 1) * g.HP;
  g.off_key = g.off_b3 + L * g.NPd;
  g.off_low = g.off_key + g.K1;
  g.par_floats = (g.off_low + L + 3) / 4 * 4;
  g.par_bytes = g.par_floats * 4;
  g.w1_slab = 2 * 128 * g.K1 * 2;
  const int hp2 = g.HP / 128, kb = g.HP / 64;
  g.off_w1 = (size_t)(g.par_bytes + 1023) / 1024 * 1024;
  g.off_hid = g.off_w1 + (size_t)hp2 * g.w1_slab;
  g.off_w3 = g.off_hid + (size_t)(g.hl - 1) * hp2 * kb * WD_HID_SLAB;
  g.total = g.off_w3 + (size_t)g.n_pass * kb * WD_OUT_SLAB;
  g.n_slabs = hp2 + (g.hl - 1) * hp2 * kb + g.n_pass * kb;
  return g;
}
static inline WdGeom wd_geom(const NscDev& st) { return wd_geom(st.U, st.C, st.H, st.n_dense, st.K, st.L); }

// ------------------------------------------------------------------------------------------
// host side: shape check + packing
// ------------------------------------------------------------------------------------------
bool wide_plan_supported(const PlanDev& hp) {
  if (hp.n_nsc < 1 || hp.n_nsc > 127) return false;
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    if (st.n_dense < 2 || st.H < 1 || st.H > WD_HP_MAX || st.K < 1 || st.K > 32 || st.L < 1) return false;
    if (st.periodic && st.K != 16 && st.K != 32) return false;  // padding bins would sit inside the wrap-around
    if (st.U + st.C > WD_IN_MAX) return false;
  }
  return true;
}

size_t wide_pack_bytes(const NscDev& st) { return wd_geom(st).total; }

// Where logit q of the kernel's layout ([KP widths | KP heights | KP derivatives]) comes from in the stage's own
// [K widths | K heights | K - 1 (periodic: K) derivatives] layout, or which padding it is (as tc_logit_source).
constexpr int WD_PAD_BIN = -1, WD_PAD_UNIT = -2, WD_PAD_NONE = -3;
static int wd_logit_source(const NscDev& st, int KP, int q) {
  const int K = st.K;
  if (q < KP) return q < K ? q : WD_PAD_BIN;
  if (q < 2 * KP) return q - KP < K ? K + (q - KP) : WD_PAD_BIN;
  const int j = q - 2 * KP, n_der = st.cpd - 2 * K;  // derivative j belongs to knot j + 1
  if (j < n_der) return 2 * K + j;
  return (j == n_der && K < KP) ? WD_PAD_UNIT : WD_PAD_NONE;
}

// element (n, k) of a K-major SWIZZLE_NONE [128 x K1] tile: 8 x 8 core matrices of 128 contiguous bytes, K-adjacent
// cores 128 B apart (LBO), 8-row groups K1 / 8 * 128 B apart (SBO)
static inline size_t wd_w1_index(int K1, int n, int k) {
  return (size_t)(n >> 3) * (K1 / 8) * 64 + (size_t)(k >> 3) * 64 + (size_t)(n & 7) * 8 + (k & 7);
}
// element (n, k), k < 64, of one K-major SWIZZLE_128B slab of `rows` rows
static inline size_t wd_sw128_index(int n, int k) { return (size_t)n * 64 + (size_t)(((k >> 3) ^ (n & 7)) * 8) + (k & 7); }

void wide_pack_stage(const PlanDev& hp, const NscDev& st, const float* const* weights, void* dst) {
  const WdGeom g = wd_geom(st);
  unsigned char* base = static_cast<unsigned char*>(dst);
  memset(base, 0, g.total);
  const int H = st.H, hp2 = g.HP / 128, kbn = g.HP / 64, out_dim = st.cpd * st.L;
  float* par = reinterpret_cast<float*>(base);
  int32_t* pi = reinterpret_cast<int32_t*>(base);
  pi[WD_PI_HL] = g.hl;
  pi[WD_PI_HP] = g.HP;
  pi[WD_PI_K1] = g.K1;
  pi[WD_PI_KP] = g.KP;
  pi[WD_PI_L] = st.L;
  pi[WD_PI_U] = st.U;
  pi[WD_PI_IN] = g.in_dim;
  pi[WD_PI_PERIODIC] = st.periodic;
  pi[WD_PI_BINS] = st.bins_off;
  pi[WD_PI_K] = st.K;
  pi[WD_PI_NPASS] = g.n_pass;
  pi[WD_PI_OFF_BH] = g.off_bh;
  pi[WD_PI_OFF_B3] = g.off_b3;
  pi[WD_PI_OFF_KEY] = g.off_key;
  pi[WD_PI_OFF_LOW] = g.off_low;
  par[WD_PF_B] = st.B;
  par[WD_PF_INV2B] = 1.0f / (2.0f * st.B);
  const int cond_col0 = hp.D + tc_log_det_levels(hp) + 2;
  for (int i = 0; i < g.K1; ++i)
    pi[g.off_key + i] = i < st.U ? (int)st.perm[i] : (i < g.in_dim ? cond_col0 + (i - st.U) : 0);
  for (int d = 0; d < st.L; ++d) pi[g.off_low + d] = (int)st.perm[st.U + d];

  {  // first layer: per N-half a [128 x K1] tile, K = (inputs..., 1 (bias row at k = in_dim), zero pad); unscaled
    const float *W1 = weights[0], *b1 = weights[1];
    for (int nh = 0; nh < hp2; ++nh) {
      __half* hi = reinterpret_cast<__half*>(base + g.off_w1 + (size_t)nh * g.w1_slab);
      __half* lo = hi + 128 * g.K1;
      for (int n = 0; n < 128; ++n)
        for (int k = 0; k < g.K1; ++k) {
          const int col = nh * 128 + n;
          float v = 0.0f;
          if (col < H) v = k < g.in_dim ? W1[(size_t)k * H + col] : (k == g.in_dim ? b1[col] : 0.0f);
          split_store(hi, lo, wd_w1_index(g.K1, n, k), v);
        }
    }
  }
  // hidden layers 2 .. hl: W[k][n] scaled by a power of two; slabs ordered (N-half, K-block)
  for (int j = 2; j <= g.hl; ++j) {
    const float *W = weights[2 * (j - 1)], *b = weights[2 * (j - 1) + 1];
    const int e = pow2_scale_exp(W, (size_t)H * H);
    const float s = ldexpf(1.0f, e);
    par[WD_PF_SINV + (j - 2)] = ldexpf(1.0f, -e);
    for (int k = 0; k < g.HP; ++k) par[g.off_bh + (j - 2) * g.HP + k] = k < H ? b[k] : 0.0f;
    for (int nh = 0; nh < hp2; ++nh)
      for (int kb = 0; kb < kbn; ++kb) {
        __half* hi = reinterpret_cast<__half*>(base + g.off_hid + ((size_t)((j - 2) * hp2 + nh) * kbn + kb) * WD_HID_SLAB);
        __half* lo = hi + 128 * 64;
        for (int n = 0; n < 128; ++n)
          for (int k = 0; k < 64; ++k) {
            const int col = nh * 128 + n, kk = kb * 64 + k;
            split_store(hi, lo, wd_sw128_index(n, k), (col < H && kk < H) ? W[(size_t)kk * H + col] * s : 0.0f);
          }
      }
  }
  {  // output layer: one slab per (pass, K-block); a pass is one K <= 32 dimension or two K <= 16 dimensions
    const float *W3 = weights[2 * g.hl], *b3 = weights[2 * g.hl + 1];
    const int e3 = pow2_scale_exp(W3, (size_t)H * out_dim);
    const float s3 = ldexpf(1.0f, e3);
    par[WD_PF_S3INV] = ldexpf(1.0f, -e3);
    const int dpp = g.KP == 32 ? 1 : 2;
    for (int p = 0; p < g.n_pass; ++p)
      for (int kb = 0; kb < kbn; ++kb) {
        __half* hi = reinterpret_cast<__half*>(base + g.off_w3 + ((size_t)p * kbn + kb) * WD_OUT_SLAB);
        __half* lo = hi + WD_PASS_N * 64;
        for (int jd = 0; jd < dpp; ++jd) {
          const int d = p * dpp + jd;
          if (d >= st.L) continue;
          for (int q = 0; q < g.NPd; ++q) {
            const int src = wd_logit_source(st, g.KP, q);
            if (src < 0) continue;
            const int col = d * st.cpd + src, n = jd * g.NPd + q;
            for (int k = 0; k < 64; ++k) {
              const int kk = kb * 64 + k;
              if (kk < H) split_store(hi, lo, wd_sw128_index(n, k), W3[(size_t)kk * out_dim + col] * s3);
            }
          }
        }
      }
    for (int d = 0; d < st.L; ++d)
      for (int q = 0; q < g.NPd; ++q) {
        const int src = wd_logit_source(st, g.KP, q);
        par[g.off_b3 + d * g.NPd + q] =
            src >= 0 ? b3[d * st.cpd + src] : (src == WD_PAD_BIN ? -1e30f : (src == WD_PAD_UNIT ? 0.5413248546129181f : 0.0f));
      }
  }
}

// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t wd_desc_w1(uint32_t saddr, int K1) {
  // K-major SWIZZLE_NONE: LBO = 128 B between K-adjacent core matrices, SBO = K1 / 8 * 128 B between 8-row groups
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((K1 * 16) >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void wd_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}

__device__ __forceinline__ uint32_t wd_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void wd_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA bulk copy global -> the same shared-memory offset of every CTA in `mask`, completing on each one's mbarrier
__device__ __forceinline__ void wd_tma_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset of every CTA in `mask`
__device__ __forceinline__ void wd_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// 32 un-normalised softmax terms of one axis (as softmax_terms, K = 32)
__device__ __forceinline__ void softmax_terms32(const float (&raw)[32], const float* __restrict__ bias, float sinv,
                                                float (&e)[32]) {
  constexpr float L2E = 1.4426950408889634f;
#pragma unroll
  for (int q = 0; q < 32; q += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + q);
    e[q + 0] = fmaf(raw[q + 0], sinv, b4.x);
    e[q + 1] = fmaf(raw[q + 1], sinv, b4.y);
    e[q + 2] = fmaf(raw[q + 2], sinv, b4.z);
    e[q + 3] = fmaf(raw[q + 3], sinv, b4.w);
  }
  float m0 = fmaxf(e[0], e[1]), m1 = fmaxf(e[2], e[3]);
#pragma unroll
  for (int j = 4; j < 32; j += 4) {
    m0 = fmaxf(m0, fmaxf(e[j], e[j + 1]));
    m1 = fmaxf(m1, fmaxf(e[j + 2], e[j + 3]));
  }
  const float nm = -(fmaxf(m0, m1) * L2E);
#pragma unroll
  for (int j = 0; j < 32; ++j) e[j] = ex2_approx(fmaf(e[j], L2E, nm));
}

// Two-level search over 32 terms: eight groups of four.  P[k] = sum of the groups 0..k (k < 7), tot = all eight.
struct BinPick32 {
  bool p[7], q1, q2, q3;  // "right edge <= target" flags of the groups / inside the selected group (prefixes of trues)
};
__device__ __forceinline__ void group_prefix32(const float (&e)[32], float (&P)[7], float& tot) {
  float g[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = (e[4 * k] + e[4 * k + 1]) + (e[4 * k + 2] + e[4 * k + 3]);
  P[0] = g[0];
#pragma unroll
  for (int k = 1; k < 7; ++k) P[k] = P[k - 1] + g[k];
  tot = P[6] + g[7];
}
template <bool SEARCH>
__device__ __forceinline__ void pick_bin32(const float (&e)[32], const float (&P)[7], float target, BinPick32& k, float& left,
                                           float& term) {
  if (SEARCH) {
#pragma unroll
    for (int i = 0; i < 7; ++i) k.p[i] = P[i] <= target;
  }
  float base = 0.0f, a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = e[i];
#pragma unroll
  for (int gq = 0; gq < 7; ++gq) {
    base = k.p[gq] ? P[gq] : base;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = k.p[gq] ? e[4 * (gq + 1) + i] : a[i];
  }
  const float c1 = base + a[0], c2 = c1 + a[1], c3 = c2 + a[2];
  if (SEARCH) {
    k.q1 = c1 <= target;
    k.q2 = c2 <= target;
    k.q3 = c3 <= target;
  }
  left = k.q3 ? c3 : (k.q2 ? c2 : (k.q1 ? c1 : base));
  term = k.q3 ? a[3] : (k.q2 ? a[2] : (k.q1 ? a[1] : a[0]));
}

// softmax knots, bin search, RQS and log-det of one (row, dim) with K = 32 (padded) bins, its 96 raw accumulator values
// read from TMEM in three parts: searched axis, other axis, derivatives (pzflow/bijectors.py:488-491 + utils.py:113-227).
// `release` is called once the last part has been read (the accumulator buffer may then be overwritten).
template <bool INV, typename Release>
__device__ __forceinline__ void wide_spline32(uint32_t t_acc, const float* __restrict__ bias, float sinv, float x, float B,
                                              float inv2B, bool periodic, Release release, float& out, float& ld, int& idx_out,
                                              bool& finite) {
  constexpr int SO = INV ? 32 : 0, OO = INV ? 0 : 32;
  const float twoB = 2.0f * B;
  const bool oob = (x <= -B) || (x >= B);
  const float mx = oob ? fabsf(x) - B : x;
  float ra[32], rb[32], e[32], P[7], tot, tot2;
  tmem_ld32(t_acc + SO, ra);
  tmem_ld32(t_acc + OO, rb);
  tc_wait_ld();
  BinPick32 k;
  softmax_terms32(ra, bias + SO, sinv, e);
  tmem_ld32(t_acc + 64, ra);  // derivative logits, in flight under the search
  group_prefix32(e, P, tot);
  const float target = (mx + B) * (tot * inv2B);  // (x + B) * sum(e) / 2B
  float cs_sel, s_sel;
  pick_bin32<true>(e, P, target, k, cs_sel, s_sel);
  const float rs = __fdividef(twoB, tot);
  int gi = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) gi += k.p[i] ? 1 : 0;
  const int wi = k.q3 ? 3 : (k.q2 ? 2 : (k.q1 ? 1 : 0));
  int idx = 4 * gi + wi;
  if (tot <= target) idx = 32;
  if (!(-B <= mx)) idx = -1;
  idx_out = idx;

  softmax_terms32(rb, bias + OO, sinv, e);
  group_prefix32(e, P, tot2);
  float co_sel, o_sel;
  pick_bin32<false>(e, P, 0.0f, k, co_sel, o_sel);
  const float ro = __fdividef(twoB, tot2);
  finite = (tot + tot2) < INFINITY;  // false for NaN too: an fp16 overflow upstream shows up here

  tc_wait_ld();
  release();
  // derivative logits D[idx - 1], D[idx]: ra[j] belongs to knot j + 1 (ra[31] is the zero pad unless periodic)
  float v[5];
  v[0] = ra[0];
#pragma unroll
  for (int i = 1; i < 5; ++i) v[i] = ra[i - 1];
#pragma unroll
  for (int gq = 0; gq < 7; ++gq)
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = k.p[gq] ? ra[4 * (gq + 1) - 1 + i] : v[i];
  float r0 = k.q3 ? v[3] : (k.q2 ? v[2] : (k.q1 ? v[1] : v[0]));
  float r1 = k.q3 ? v[4] : (k.q2 ? v[3] : (k.q1 ? v[2] : v[1]));
  const int ic = min(max(idx, 0), 31);
  r0 = fmaf(r0, sinv, bias[64 + max(ic - 1, 0)]);
  r1 = fmaf(r1, sinv, bias[64 + ic]);

  const float qnan = __int_as_float(0x7fc00000);
  const bool bad = (idx < 0) || (idx >= 32);  // take_along_axis "fill" -> NaN (pzflow/utils.py:155-161)
  SplineKnotSel s;
  s.idx = idx;
  const float ksel = fmaf(cs_sel, rs, -B), oknot = fmaf(co_sel, ro, -B);
  const float ssz = bad ? qnan : s_sel * rs, osz = bad ? qnan : o_sel * ro;
  s.wk = INV ? osz : ssz;
  s.hk = INV ? ssz : osz;
  s.xk = INV ? oknot : ksel;
  s.yk = INV ? ksel : oknot;
  if (periodic) {
    if (idx == 0) r0 = fmaf(ra[31], sinv, bias[64 + 31]);  // dk = pad(D, (1, 0), "wrap")
    s.dk = tc_softplus(r0);
    s.dkp1 = tc_softplus(r1);
  } else {
    s.dk = (idx == 0) ? 1.0f : tc_softplus(r0);
    s.dkp1 = (idx == 31) ? 1.0f : tc_softplus(r1);
  }
  tc_rqs_eval(s, x, mx, oob, INV, periodic, out, ld);
}

struct WdStage {
  const uint8_t* blob;
  uint32_t par_bytes;
  uint8_t k1_16, hp2, hl, n_pass, kp32, nsc_idx, pad0, pad1;
};

struct WdShared {
  uint64_t slot_full[WD_MAX_SLOTS + 1], slot_free[WD_MAX_SLOTS + 1];
  uint64_t par_full[2], par_free[2];
  uint64_t key_ready;     // 8 warp arrivals: the stage's key operand is written
  uint64_t h_full[2];     // [half]: the dense layer's accumulator columns of that half have landed (commit)
  uint64_t act_ready[2];  // [half] (4 warp arrivals): that half's columns are rewritten as the next A operand
  uint64_t d3_full[2];    // [buffer]: an output pass has landed
  uint64_t d3_free[2];    // [buffer] (4 warp arrivals): drained
  uint32_t tmem_base;
  uint32_t ovf[WD_TILE / 32];     // one bit per tile row: FINITE conditioner inputs gave non-finite logits in some stage
  uint32_t kok[2][WD_TILE / 32];  // [half][..] one bit per tile row: that half's block of the stage's conditioner inputs is finite
  int n_seq;
  WdStage stage[MAX_NSC];
};

struct WdLaunch {
  int n_slots, par_max, state_cols, n_lev;
  int cluster;  // CTAs per cluster (1, 2, 4 or 8): every CTA fetches 1 / cluster of each slab and multicasts it to all
  int debug;    // experiments only (PZB_WIDE_DEBUG; results are garbage): 1 = skip the spline arithmetic, 2 = fetch 1 KB per slab
};
struct WdSmem {
  int ring, par, state, shared, total;
};
__host__ __device__ inline WdSmem wd_smem_layout(const WdLaunch& lp) {
  WdSmem s;
  s.ring = 0;
  s.par = lp.n_slots * WD_SLOT_BYTES;
  s.state = s.par + 2 * lp.par_max;
  s.shared = (s.state + WD_TILE * lp.state_cols * 4 + 15) / 16 * 16;
  s.total = s.shared + (int)sizeof(WdShared);
  return s;
}

#define WD_BAR(field) (sh_addr + (uint32_t)offsetof(WdShared, field))

template <bool INV>
__global__ void __launch_bounds__(WD_THREADS, 1)
nsf_chain_wide_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, WdLaunch lp, unsigned long long* __restrict__ overflow_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const WdSmem sl = wd_smem_layout(lp);
  WdShared* sh = reinterpret_cast<WdShared*>(base + sl.shared);
  const uint32_t sh_addr = smem_u32(sh);
  float* state = reinterpret_cast<float*>(base + sl.state);
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  const int D = plan->D, C = plan->C, ns = plan->n_steps, n_lev = lp.n_lev;
  constexpr int CR = WD_TILE;
  const int COL_PEND = D + n_lev, COL_COND = D + n_lev + 2;
  const int n_slots = lp.n_slots;

  if (tid == 0) {
    for (int i = 0; i < n_slots; ++i) {
      mbar_init(WD_BAR(slot_full[0]) + 8u * i, 1);
      mbar_init(WD_BAR(slot_free[0]) + 8u * i, lp.cluster);  // every CTA of the cluster has consumed the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(WD_BAR(par_full[0]) + 8u * i, 1);
      mbar_init(WD_BAR(par_free[0]) + 8u * i, WD_EPI_THREADS / 32);
      mbar_init(WD_BAR(h_full[0]) + 8u * i, 1);
      mbar_init(WD_BAR(act_ready[0]) + 8u * i, WD_TILE / 32);
      mbar_init(WD_BAR(d3_full[0]) + 8u * i, 1);
      mbar_init(WD_BAR(d3_free[0]) + 8u * i, WD_TILE / 32);
    }
    mbar_init(WD_BAR(key_ready), WD_EPI_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int seen = 0;
    for (int si = 0; si < ns; ++si) {
      const StepDev step = plan->steps[INV ? ns - 1 - si : si];
      if (step.kind != STEP_NSC) continue;
      const NscDev& st = plan->nsc[step.idx];
      const WdGeom g = wd_geom(st.U, st.C, st.H, st.n_dense, st.K, st.L);
      WdStage& w = sh->stage[seen++];
      w.blob = static_cast<const uint8_t*>(st.wide_blob);
      w.par_bytes = (uint32_t)g.par_bytes;
      w.k1_16 = (uint8_t)(g.K1 / 16);
      w.hp2 = (uint8_t)(g.HP / 128);
      w.hl = (uint8_t)g.hl;
      w.n_pass = (uint8_t)g.n_pass;
      w.kp32 = (uint8_t)(g.KP == 32);
      w.nsc_idx = (uint8_t)step.idx;
    }
    sh->n_seq = seen;
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(WD_BAR(tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (lp.cluster > 1) wd_cluster_sync();  // every CTA's mbarriers are initialised before a peer's copy or commit may hit them
  tc_fence_after();
  const uint32_t tbase = sh->tmem_base;
  const int n_seq = sh->n_seq;

  // The CTAs of a cluster share every slab they fetch, so they walk the SAME number of tiles (the last round may hold
  // tiles past the end: those run on zero rows and write nothing).
  const long long n_tiles = (args.N + WD_TILE - 1) / WD_TILE;
  const int my_tiles = lp.cluster > 1 ? (int)((n_tiles + gridDim.x - 1) / gridDim.x)
                                      : (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t crank = lp.cluster > 1 ? wd_cluster_rank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << lp.cluster) - 1u);

  if (warp < 8) {
    // ======================================= epilogue warps =======================================
    const int half = warp >> 2;
    const int lrow = ((warp & 3) << 5) + lane;  // row within the tile == TMEM lane
    const uint32_t t_lane = tbase + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t bar_h = WD_BAR(h_full[0]) + 8u * half, bar_act = WD_BAR(act_ready[0]) + 8u * half;
    const uint32_t bar_d3 = WD_BAR(d3_full[0]) + 8u * half, bar_free = WD_BAR(d3_free[0]) + 8u * half;
    uint32_t ph_h = 0u, ph_d3 = 0u;
    int seq = 0;
    float* srow = state + lrow;

    int first_nsc = ns, last_nsc = -1;
    for (int si = 0; si < ns; ++si)
      if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
        if (first_nsc == ns) first_nsc = si;
        last_nsc = si;
      }

    for (int ti = 0; ti < my_tiles; ++ti) {
      const long long row0 = ((long long)blockIdx.x + (long long)ti * gridDim.x) * WD_TILE;
      const int nrows = (int)max(0ll, min((long long)WD_TILE, args.N - row0));
      const long long row = row0 + lrow;
      bool in_ok = false;   // (half 0) this row's inputs are all finite

      // ---- pass A: the tile's rows -> state, steps in front of the first coupling stage (one thread per row) ----
      if (half == 0) {
        if (lrow < nrows) {
          float chk = 0.0f;
          const RowIndex ri = pz_row_index(args, row);
          for (int j = 0; j < D; ++j) {
            const float v = pz_input_value(args, ri, j, D);
            srow[(INV ? plan->perm_final[j] : j) * CR] = v;
            chk += v * 0.0f;   // NaN iff some input is NaN or infinite
          }
          for (int j = 0; j < C; ++j) {
            const float v = pz_cond_value(args, ri, j, C);
            srow[(COL_COND + j) * CR] = v;
            chk += v * 0.0f;
          }
          in_ok = chk == 0.0f;
        } else {
          for (int j = 0; j < D; ++j) srow[j * CR] = 0.0f;
          for (int j = 0; j < C; ++j) srow[(COL_COND + j) * CR] = 0.0f;
        }
        for (int l = 0; l < n_lev; ++l) srow[(D + l) * CR] = 0.0f;
        if (lane == 0) sh->ovf[lrow >> 5] = 0u;
        int lv = 0;
        light_steps_cols(plan, INV, 0, first_nsc, lv, srow, CR);
      }
      int level = level_after(plan, INV, 0, first_nsc, 0);

      int prev_end = first_nsc;
      bool pending = false;
      int pend_level = 0;
      for (int si = first_nsc; si <= last_nsc; ++si) {
        if (plan->steps[INV ? ns - 1 - si : si].kind != STEP_NSC) continue;
        const bool has_light = si > prev_end;
        const int light0 = prev_end, level_in = level;
        if (has_light) level = level_after(plan, INV, prev_end, si, level);
        prev_end = si + 1;

        mbar_wait(WD_BAR(par_full[0]) + 8u * (seq & 1), (uint32_t)((seq >> 1) & 1));
        named_bar_sync(1, WD_EPI_THREADS);  // the previous stage's (or pass A's) state writes of the row's other half
        const float* par = reinterpret_cast<const float*>(base + sl.par + (seq & 1) * lp.par_max);
        const int* pint = reinterpret_cast<const int*>(par);
        const int hl = pint[WD_PI_HL], HP = pint[WD_PI_HP], K1 = pint[WD_PI_K1], L = pint[WD_PI_L];
        const int n_pass = pint[WD_PI_NPASS], in_dim = pint[WD_PI_IN];
        const bool kp32 = pint[WD_PI_KP] == 32;
        const bool periodic = pint[WD_PI_PERIODIC] != 0;

        if (pending || has_light) {
          if (half == 0) {
            if (pending) srow[(D + pend_level) * CR] += srow[COL_PEND * CR] + srow[(COL_PEND + 1) * CR];
            if (has_light) {
              int lv = level_in;
              light_steps_cols(plan, INV, light0, si, lv, srow, CR);
            }
          }
          if (has_light) named_bar_sync(1, WD_EPI_THREADS);
        }

        // ---- E0: conditioner inputs + the constant 1 of the bias row -> fp16 hi/lo A operand in Q[0, K1) ----
        {
          const int* key_col = pint + pint[WD_PI_OFF_KEY];
          float chk = 0.0f;
          for (int b = half; b * 32 < K1; b += 2) {
            uint32_t rh[16], rl[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k0 = 32 * b + 2 * j, k1 = k0 + 1;
              const float v0 = k0 < in_dim ? srow[key_col[k0] * CR] : (k0 == in_dim ? 1.0f : 0.0f);
              const float v1 = k1 < in_dim ? srow[key_col[k1] * CR] : (k1 == in_dim ? 1.0f : 0.0f);
              chk += v0 + v1;
              split2(v0, v1, rh[j], rl[j]);
            }
            tmem_st16(t_lane + 256u + 32u * b, rh);
            tmem_st16(t_lane + 256u + 32u * b + 16u, rl);
          }
          {  // finite inputs sum to something finite; NaN or infinite ones do not (inf - inf -> NaN)
            const uint32_t ok = __ballot_sync(0xffffffffu, fabsf(chk) < INFINITY);
            if (lane == 0) sh->kok[half][lrow >> 5] = ok;
          }
          tc_wait_st();
          tc_fence_before();
          warp_arrive(WD_BAR(key_ready), lane);
        }

        // ---- dense layers 1 .. hl: accumulator columns of this half -> LeakyReLU -> fp16 hi/lo, in place ----
        const int ncol = HP >> 1;  // columns per half
        for (int j = 1; j <= hl; ++j) {
          const uint32_t t_reg = t_lane + ((j & 1) ? 0u : 256u) + (uint32_t)(half * ncol);
          const float sinv = j >= 2 ? par[WD_PF_SINV + (j - 2)] : 1.0f;
          const float* bj = par + pint[WD_PI_OFF_BH] + (j >= 2 ? (j - 2) * HP : 0) + half * ncol;
          mbar_wait(bar_h, ph_h);
          ph_h ^= 1u;
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < ncol; c += 32) {
            float dv[32];
            tmem_ld32(t_reg + c, dv);
            tc_wait_ld();
            uint32_t rh[16], rl[16];
            if (j == 1) {  // bias rode in the GEMM
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                float h0, h1;
                leaky2(dv[2 * q], dv[2 * q + 1], h0, h1);
                split2(h0, h1, rh[q], rl[q]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 bb = *reinterpret_cast<const float4*>(bj + c + q * 4);
                const float2 sc = make_float2(sinv, sinv);
                const float2 p01 = __ffma2_rn(make_float2(dv[4 * q + 0], dv[4 * q + 1]), sc, make_float2(bb.x, bb.y));
                const float2 p23 = __ffma2_rn(make_float2(dv[4 * q + 2], dv[4 * q + 3]), sc, make_float2(bb.z, bb.w));
                float h0, h1, h2, h3;
                leaky2(p01.x, p01.y, h0, h1);
                leaky2(p23.x, p23.y, h2, h3);
                split2(h0, h1, rh[2 * q], rl[2 * q]);
                split2(h2, h3, rh[2 * q + 1], rl[2 * q + 1]);
              }
            }
            tmem_st16(t_reg + c, rh);
            tmem_st16(t_reg + c + 16, rl);
          }
          tc_wait_st();
          tc_fence_before();
          warp_arrive(bar_act, lane);
        }

        // ---- output passes p = half, half + 2, ...: logits -> knots -> spline ----
        const uint32_t t_acc = t_lane + ((hl & 1) ? 256u : 0u) + 128u * half;
        const float s3inv = par[WD_PF_S3INV], Bs = par[WD_PF_B], inv2B = par[WD_PF_INV2B];
        const float* b3 = par + pint[WD_PI_OFF_B3];
        const int* low_col = pint + pint[WD_PI_OFF_LOW];
        float ldsum = 0.0f;
#pragma unroll 1
        for (int p = half; p < n_pass; p += 2) {
          mbar_wait(bar_d3, ph_d3);
          ph_d3 ^= 1u;
          tc_fence_after();
          if (lp.debug & 1) {
            float raw[16];
            tmem_ld16(t_acc, raw);
            tc_wait_ld();
            tc_fence_before();
            warp_arrive(bar_free, lane);
            ldsum += raw[0];
          } else if (kp32) {
            const int col = low_col[p];
            const float xin = srow[col * CR];
            float y, ldd;
            int idx;
            bool fin;
            wide_spline32<INV>(t_acc, b3 + p * 96, s3inv, xin, Bs, inv2B, periodic,
                               [&]() {
                                 tc_fence_before();
                                 warp_arrive(bar_free, lane);
                               },
                               y, ldd, idx, fin);
            if (!fin && ((sh->kok[0][lrow >> 5] & sh->kok[1][lrow >> 5]) >> (lrow & 31)) & 1u)
              atomicOr(&sh->ovf[lrow >> 5], 1u << (lrow & 31));   // (rare) finite inputs, non-finite logits: pzb_plan_overflow_rows
            ldsum += ldd;
            srow[col * CR] = y;
            if (args.bins != nullptr && lrow < nrows)
              args.bins[row * plan->n_spline_dims + pint[WD_PI_BINS] + p] = min(idx, pint[WD_PI_K]);
          } else {
            const int nd = min(2, L - 2 * p);
#pragma unroll 1
            for (int jd = 0; jd < nd; ++jd) {
              const int d = 2 * p + jd;
              float raw[48];
              tmem_ld32(t_acc + 48u * jd, raw);
              tmem_ld16(t_acc + 48u * jd + 32u, raw + 32);
              tc_wait_ld();
              if (jd == nd - 1) {
                tc_fence_before();
                warp_arrive(bar_free, lane);
              }
              const int col = low_col[d];
              const float xin = srow[col * CR];
              float y, ldd;
              int idx;
              bool fin;
              spline_dim<INV>(raw, b3 + d * 48, s3inv, xin, Bs, inv2B, periodic, y, ldd, idx, fin);
              if (!fin && ((sh->kok[0][lrow >> 5] & sh->kok[1][lrow >> 5]) >> (lrow & 31)) & 1u)
                atomicOr(&sh->ovf[lrow >> 5], 1u << (lrow & 31));
              ldsum += ldd;
              srow[col * CR] = y;
              if (args.bins != nullptr && lrow < nrows)
                args.bins[row * plan->n_spline_dims + pint[WD_PI_BINS] + d] = min(idx, pint[WD_PI_K]);
            }
          }
        }
        srow[(COL_PEND + half) * CR] = INV ? -ldsum : ldsum;
        warp_arrive(WD_BAR(par_free[0]) + 8u * (seq & 1), lane);
        pending = true;
        pend_level = level;
        ++seq;
      }
      named_bar_sync(1, WD_EPI_THREADS);

      // ---- pass Z: trailing steps, latent log-density, outputs (one thread per row) ----
      bool bad_row = false;
      if (half == 0 && lrow < nrows) {
        if (pending) srow[(D + pend_level) * CR] += srow[COL_PEND * CR] + srow[(COL_PEND + 1) * CR];
        int lv = level;
        light_steps_cols(plan, INV, prev_end, ns, lv, srow, CR);
        bad_row = in_ok && ((sh->ovf[lrow >> 5] >> (lrow & 31)) & 1u);   // finite inputs, non-finite logits: pzb_plan_overflow_rows
        if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
          const float lat = pz_latent_logprob(plan, [&](int j) { return srow[plan->perm_final[j] * CR]; });
          const float lpv = pz_nan_to_num_lp(lat + srow[D * CR]);
          args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
        } else {
          for (int j = 0; j < D; ++j) args.out0[row * D + j] = srow[(INV ? j : plan->perm_final[j]) * CR];
          if (args.out1 != nullptr) args.out1[row] = srow[D * CR];
        }
      }
      if (half == 0 && overflow_rows != nullptr) {
        const unsigned m = __ballot_sync(0xffffffffu, bad_row);
        if (lane == 0 && m != 0u) {
          atomicAdd(overflow_rows, (unsigned long long)__popc(m));
          if (args.ovf_flag != nullptr) *args.ovf_flag = 1u;
        }
      }
      // (pass A of the next tile touches a row from the same thread as pass Z did: no barrier needed)
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t sb = smem_u32(base);
    const uint32_t tP = __shfl_sync(0xffffffffu, tbase, 0), tQ = tP + 256u;
    const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc96 = (1u << 4) | ((uint32_t)(WD_PASS_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t g = 0u;             // slabs consumed so far
    uint32_t ph_key = 0u, ph_act = 0u;
    uint32_t np[2] = {0u, 0u};   // passes issued so far into each accumulator buffer
    const int total_seq = my_tiles * n_seq;
    int sq = 0;
    for (int s = 0; s < total_seq; ++s, sq = (sq + 1 == n_seq ? 0 : sq + 1)) {
      const WdStage w = sh->stage[sq];
      const int k1_16 = w.k1_16, hp2 = w.hp2, hl = w.hl, n_pass = w.n_pass, kbn = 2 * w.hp2;
      const int K1 = 16 * k1_16;
      // ---- layer 1: keys (Q[0, K1)) x W1 -> P ----
      mbar_wait(WD_BAR(key_ready), ph_key);
      ph_key ^= 1u;
      tc_fence_after();
      for (int nh = 0; nh < hp2; ++nh) {
        const uint32_t slot = g % (uint32_t)n_slots;
        mbar_wait(WD_BAR(slot_full[0]) + 8u * slot, (g / (uint32_t)n_slots) & 1u);
        tc_fence_after();
        const uint32_t slab = sb + slot * WD_SLOT_BYTES;
        if (elect_one()) {
          const uint64_t dh = wd_desc_w1(slab, K1), dl = wd_desc_w1(slab + 128 * K1 * 2, K1);
#pragma unroll 1
          for (int prod = 0; prod < 3; ++prod)
#pragma unroll 1
            for (int k16 = 0; k16 < k1_16; ++k16)
              wd_mma(tP + 128u * nh, tQ + (prod == 0 ? 16u : 0u) + a_col_of(k16),
                     (prod == 1 ? dl : dh) + (uint64_t)(k16 * 16), idesc128, !(prod == 0 && k16 == 0));
          if (lp.cluster > 1) wd_commit_multicast(WD_BAR(slot_free[0]) + 8u * slot, cmask);
          else umma_commit(WD_BAR(slot_free[0]) + 8u * slot);
          if (hp2 == 2) {
            umma_commit(WD_BAR(h_full[0]) + 8u * nh);
          } else {
            umma_commit(WD_BAR(h_full[0]));
            umma_commit(WD_BAR(h_full[0]) + 8u);
          }
        }
        __syncwarp();
        ++g;
      }
      // ---- hidden layers 2 .. hl: region(j - 1) x Wj -> region(j) ----
      for (int j = 2; j <= hl; ++j) {
        const uint32_t src = (j & 1) ? tQ : tP, dst = (j & 1) ? tP : tQ;
        mbar_wait(WD_BAR(act_ready[0]), ph_act);
        mbar_wait(WD_BAR(act_ready[0]) + 8u, ph_act);
        ph_act ^= 1u;
        tc_fence_after();
        for (int nh = 0; nh < hp2; ++nh)
          for (int kb = 0; kb < kbn; ++kb) {
            const uint32_t slot = g % (uint32_t)n_slots;
            mbar_wait(WD_BAR(slot_full[0]) + 8u * slot, (g / (uint32_t)n_slots) & 1u);
            tc_fence_after();
            const uint32_t slab = sb + slot * WD_SLOT_BYTES;
            if (elect_one()) {
              const uint64_t dh = make_b_desc(slab), dl = make_b_desc(slab + WD_HID_SLAB / 2);
#pragma unroll
              for (int prod = 0; prod < 3; ++prod)
#pragma unroll
                for (int kl = 0; kl < 4; ++kl)
                  wd_mma(dst + 128u * nh, src + (prod == 0 ? 16u : 0u) + a_col_of(4 * kb + kl),
                         (prod == 1 ? dl : dh) + (uint64_t)(kl * 2), idesc128, !(kb == 0 && prod == 0 && kl == 0));
              if (lp.cluster > 1) wd_commit_multicast(WD_BAR(slot_free[0]) + 8u * slot, cmask);
              else umma_commit(WD_BAR(slot_free[0]) + 8u * slot);
              if (kb == kbn - 1) {
                if (hp2 == 2) {
                  umma_commit(WD_BAR(h_full[0]) + 8u * nh);
                } else {
                  umma_commit(WD_BAR(h_full[0]));
                  umma_commit(WD_BAR(h_full[0]) + 8u);
                }
              }
            }
            __syncwarp();
            ++g;
          }
      }
      // ---- output passes: region(hl) x W3 -> the other region, buffers at +0 / +128 ----
      {
        const uint32_t src = (hl & 1) ? tP : tQ, acc = (hl & 1) ? tQ : tP;
        mbar_wait(WD_BAR(act_ready[0]), ph_act);
        mbar_wait(WD_BAR(act_ready[0]) + 8u, ph_act);
        ph_act ^= 1u;
        tc_fence_after();
        for (int p = 0; p < n_pass; ++p) {
          const int b = p & 1;
          if (np[b] > 0u) {  // the buffer's previous pass must have been drained
            mbar_wait(WD_BAR(d3_free[0]) + 8u * b, (np[b] - 1u) & 1u);
            tc_fence_after();
          }
          for (int kb = 0; kb < kbn; ++kb) {
            const uint32_t slot = g % (uint32_t)n_slots;
            mbar_wait(WD_BAR(slot_full[0]) + 8u * slot, (g / (uint32_t)n_slots) & 1u);
            tc_fence_after();
            const uint32_t slab = sb + slot * WD_SLOT_BYTES;
            if (elect_one()) {
              const uint64_t dh = make_b_desc(slab), dl = make_b_desc(slab + WD_OUT_SLAB / 2);
#pragma unroll
              for (int prod = 0; prod < 3; ++prod)
#pragma unroll
                for (int kl = 0; kl < 4; ++kl)
                  wd_mma(acc + 128u * b, src + (prod == 0 ? 16u : 0u) + a_col_of(4 * kb + kl),
                         (prod == 1 ? dl : dh) + (uint64_t)(kl * 2), idesc96, !(kb == 0 && prod == 0 && kl == 0));
              if (lp.cluster > 1) wd_commit_multicast(WD_BAR(slot_free[0]) + 8u * slot, cmask);
              else umma_commit(WD_BAR(slot_free[0]) + 8u * slot);
              if (kb == kbn - 1) umma_commit(WD_BAR(d3_full[0]) + 8u * b);
            }
            __syncwarp();
            ++g;
          }
          ++np[b];
        }
      }
    }
  } else {
    // =============================== weight loader ===============================
    const uint32_t sb = smem_u32(base);
    uint32_t g = 0u;
    const int total_seq = my_tiles * n_seq;
    int sq = 0;
    for (int s = 0; s < total_seq; ++s, sq = (sq + 1 == n_seq ? 0 : sq + 1)) {
      const WdStage w = sh->stage[sq];
      const uint8_t* src = w.blob;
      {  // parameter block: buffer s & 1 was last used by stage s - 2
        if (s >= 2) mbar_wait(WD_BAR(par_free[0]) + 8u * (s & 1), (uint32_t)(((s >> 1) - 1) & 1));
        const uint32_t bar = WD_BAR(par_full[0]) + 8u * (s & 1);
        if (lane == 0) {
          mbar_expect_tx(bar, w.par_bytes);
          tma_bulk_g2s(sb + sl.par + (s & 1) * lp.par_max, src, w.par_bytes, bar);
        }
        __syncwarp();
      }
      const int hp2 = w.hp2, kbn = 2 * w.hp2;
      const uint32_t w1_slab = 2u * 128u * 16u * w.k1_16 * 2u;
      const int n_w1 = hp2, n_hid = (w.hl - 1) * hp2 * kbn, n_out = w.n_pass * kbn;
      const uint8_t* p = src + ((size_t)(w.par_bytes + 1023u) / 1024u * 1024u);
      for (int i = 0; i < n_w1 + n_hid + n_out; ++i) {
        uint32_t bytes = i < n_w1 ? w1_slab : (i < n_w1 + n_hid ? (uint32_t)WD_HID_SLAB : (uint32_t)WD_OUT_SLAB);
        const uint32_t stride = bytes;
        if (lp.debug & 2) bytes = 1024u;
        const uint32_t slot = g % (uint32_t)n_slots;
        if (g >= (uint32_t)n_slots) mbar_wait(WD_BAR(slot_free[0]) + 8u * slot, ((g / (uint32_t)n_slots) - 1u) & 1u);
        const uint32_t bar = WD_BAR(slot_full[0]) + 8u * slot;
        if (lane == 0) {
          mbar_expect_tx(bar, bytes);  // the whole slab lands here: this CTA's share and its peers'
          if (lp.cluster > 1) {
            const uint32_t part = bytes / (uint32_t)lp.cluster, o0 = crank * part;  // (slab sizes are multiples of 128 B x 8)
            for (uint32_t off = 0; off < part; off += 8192u)
              wd_tma_multicast(sb + slot * WD_SLOT_BYTES + o0 + off, p + o0 + off, min(8192u, part - off), bar, cmask);
          } else {
            for (uint32_t off = 0; off < bytes; off += 8192u)
              tma_bulk_g2s(sb + slot * WD_SLOT_BYTES + off, p + off, min(8192u, bytes - off), bar);
          }
        }
        __syncwarp();
        p += stride;
        ++g;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (lp.cluster > 1) wd_cluster_sync();  // no CTA leaves while a peer's copies or commits can still land in it
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

cudaError_t launch_wide(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms,
                        unsigned long long* overflow_rows, cudaStream_t stream) {
  WdLaunch lp;
  lp.par_max = 16;
  for (int s = 0; s < hp.n_nsc; ++s) lp.par_max = max(lp.par_max, wd_geom(hp.nsc[s]).par_bytes);
  lp.par_max = (lp.par_max + 127) / 128 * 128;
  lp.n_lev = tc_log_det_levels(hp);
  lp.state_cols = hp.D + lp.n_lev + 2 + hp.C;  // x, log-det levels, pending halves, conditions
  lp.n_slots = WD_MAX_SLOTS;
  if (const char* env = getenv("PZB_WIDE_SLOTS")) {  // experiments: cap the ring depth
    const int cap = atoi(env);
    if (cap >= 3 && cap < lp.n_slots) lp.n_slots = cap;
  }
  while (lp.n_slots > 2 && wd_smem_layout(lp).total + 1024 > WD_SMEM_MAX) --lp.n_slots;
  const WdSmem sl = wd_smem_layout(lp);
  const size_t smem = (size_t)sl.total + 1024;
  if (smem > (size_t)WD_SMEM_MAX || lp.n_slots < 3) return cudaErrorInvalidConfiguration;
  const bool inv = (args.mode == MODE_INVERSE);
  cudaError_t e = inv ? cudaFuncSetAttribute(nsf_chain_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(nsf_chain_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long n_tiles = (args.N + WD_TILE - 1) / WD_TILE;
  // Clusters share their weight fetches (each CTA fetches 1 / cluster of a slab and multicasts it), which is what the
  // L2 -> SM path needs: 148 CTAs streaming 1.8 MB per tile and stage on their own saturate it at half the tensor rate.
  lp.cluster = 1;
  lp.debug = 0;
  if (const char* env = getenv("PZB_WIDE_DEBUG")) lp.debug = atoi(env);
  if (const char* env = getenv("PZB_WIDE_CLUSTER")) {
    const int c = atoi(env);
    if (c == 1 || c == 2 || c == 4 || c == 8) lp.cluster = c;
  }
  long long grid = num_sms / lp.cluster * lp.cluster;
  const long long want = (n_tiles + lp.cluster - 1) / lp.cluster * lp.cluster;
  if (grid > want) grid = want;
  if (grid < lp.cluster) grid = lp.cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)lp.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (lp.cluster > 1) {
    // a persistent grid must be co-resident: as many clusters as the device can hold at once (GPC boundaries can strand SMs)
    int max_clusters = 0;
    cudaLaunchConfig_t probe = cfg;
    probe.gridDim = dim3((unsigned)(num_sms / lp.cluster * lp.cluster));
    cudaError_t oe = inv ? cudaOccupancyMaxActiveClusters(&max_clusters, nsf_chain_wide_kernel<true>, &probe)
                         : cudaOccupancyMaxActiveClusters(&max_clusters, nsf_chain_wide_kernel<false>, &probe);
    if (oe != cudaSuccess) return oe;
    if (max_clusters < 1) return cudaErrorInvalidConfiguration;
    if ((long long)max_clusters * lp.cluster < grid) cfg.gridDim = dim3((unsigned)(max_clusters * lp.cluster));
  }
  if (inv) return cudaLaunchKernelEx(&cfg, nsf_chain_wide_kernel<true>, d_plan, args, lp, overflow_rows);
  return cudaLaunchKernelEx(&cfg, nsf_chain_wide_kernel<false>, d_plan, args, lp, overflow_rows);
}

}  // namespace pzb
