// tcgen05 engine: fused chain kernel with the conditioner's two big GEMMs on the 5th-gen
// tensor cores (sm_100a), for the default conditioner family (hidden_layers = 2, hidden_dim = 128,
// K = 16, at most 8 conditioner inputs and 4 transformed columns: BASELINE configs 1-4).
//
// fp32 fidelity: every fp32 operand x is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits), and a GEMM is three kind::f16 MMAs, lo*hi + hi*lo + hi*hi, accumulated
// in fp32 in TMEM.  Weights are pre-scaled by a power of two so their lo parts stay normal.
// Measured against the float64 twin this is indistinguishable from an fp32 GEMM.
//
// Structure (one persistent CTA per SM, 544 threads):
//   * 16 epilogue warps in two groups of 8.  A group owns one 128-row tile at a time; TMEM lane =
//     row, and TWO threads share a row (warp w and w+4 of the group address the same TMEM lane
//     quarter): each does half of the hidden columns in E1/E2 and one of the two spline
//     dimensions of an output-layer pass in E3.  The groups ping-pong so the tensor pipe works on
//     one tile while CUDA cores run the other tile's epilogue.
//   * warp 16, one elected lane: an event loop that issues every tcgen05.mma / tcgen05.commit
//     (whichever group is ready first is served first) and streams the weights of the NEXT
//     coupling stage with TMA bulk copies as soon as the tensor pipe has released a buffer:
//     W2 (64 KB) after the stage's last hidden GEMM, W3 (48/96 KB) after its last output pass,
//     the small fp32 parameter block (W1, biases) double buffered.
//   * A operands (the activations) never touch shared memory: epilogue threads write the fp16
//     hi/lo pairs straight into TMEM (tcgen05.st) and the MMA reads A from TMEM (.ts form);
//     B (weights) comes from shared memory, K-major SWIZZLE_128B.
//   * TMEM (512 columns): group 0: A[0,128) D[128,256); group 1: A[256,384) D[384,512).
//   * stage-outer order: a CTA takes a chunk of up to 1024 rows whose state (x[D], log-det
//     levels, conditions) lives in shared memory as [column][row] for the whole chain, and walks
//     the chunk's tiles once per coupling stage.  HBM sees the input rows and the result only.
//
// Follows the same reference lines as nsf_simt.cu.
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "pzb_common.cuh"

namespace pzb {

constexpr int TC_TILE = 128;
constexpr int TC_GROUP_THREADS = 256;
constexpr int TC_EPI_THREADS = 512;
constexpr int TC_THREADS = 544;
constexpr int TC_IN_MAX = 8;
constexpr int TC_L_MAX = 4;
constexpr int TC_D_MAX = 8;
constexpr int TC_MAX_CHUNK_TILES = 16;
constexpr int TC_PASS_N = 96;                     // two spline dims x 48 logit columns per output-layer pass
constexpr int TC_W2_BYTES = 2 * 128 * 128 * 2;    // fp16 hi + lo [128 x 128] tiles
constexpr int TC_W3_PASS_BYTES = 2 * TC_PASS_N * 128 * 2;  // fp16 hi + lo [96 x 128] tiles of one pass
// fp32 parameter block: W1[TC_IN_MAX][128], b1[128], b2[128], b3[2 * 96], scal[4]
constexpr int TC_PAR_FLOATS = TC_IN_MAX * 128 + 128 + 128 + 2 * TC_PASS_N + 4;
constexpr int TC_PAR_BYTES = TC_PAR_FLOATS * 4;   // 5904, a multiple of 16
constexpr int TC_SMEM_MAX = 232448;               // 227 KB opt-in limit per CTA

// global blob of one coupling stage: [W2 hi | W2 lo | pass0 hi | pass0 lo | pass1 hi | pass1 lo | par]
struct TcLayout {
  int w3, par, total, n_pass;
  __host__ __device__ static TcLayout make(int L) {
    TcLayout t;
    t.n_pass = (L + 1) / 2;
    t.w3 = TC_W2_BYTES;
    t.par = t.w3 + t.n_pass * TC_W3_PASS_BYTES;
    t.total = t.par + TC_PAR_BYTES;
    return t;
  }
};

// ------------------------------------------------------------------------------------------
// host side: shape check + packing
// ------------------------------------------------------------------------------------------
bool tc_plan_supported(const PlanDev& hp) {
  if (hp.n_nsc < 1 || hp.D > TC_D_MAX) return false;
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    if (st.n_dense != 3 || st.H != 128 || st.K != 16) return false;
    if (st.U + st.C > TC_IN_MAX || st.L > TC_L_MAX || st.L < 1) return false;
  }
  return true;
}

size_t tc_pack_bytes(const NscDev& st) { return (size_t)TcLayout::make(st.L).total; }

static int pow2_scale_exp(const float* w, size_t n) {
  float m = 0.0f;
  for (size_t i = 0; i < n; ++i) m = fmaxf(m, fabsf(w[i]));
  if (!(m > 0.0f) || !std::isfinite(m)) return 0;
  int e = (int)floorf(log2f(16384.0f / m));
  if (e > 24) e = 24;
  if (e < -24) e = -24;
  return e;
}

// element (n, k) of a K-major SWIZZLE_128B tile with `rows` rows and K = 128 halfs
static inline size_t sw128_index(int rows, int n, int k) {
  const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
  return (size_t)kb * rows * 64 + (size_t)n * 64 + (size_t)((c ^ (n & 7)) * 8) + e;
}

static inline void split_store(__half* hi, __half* lo, size_t idx, float v) {
  const __half h = __float2half_rn(v);
  hi[idx] = h;
  lo[idx] = __float2half_rn(v - __half2float(h));
}

void tc_pack_stage(const NscDev& st, const float* const* weights, void* dst) {
  const TcLayout lay = TcLayout::make(st.L);
  unsigned char* base = static_cast<unsigned char*>(dst);
  memset(base, 0, lay.total);
  const float *W1 = weights[0], *b1 = weights[1], *W2 = weights[2], *b2 = weights[3], *W3 = weights[4],
              *b3 = weights[5];
  const int in_dim = st.U + st.C, out_dim = st.cpd * st.L;
  const int e2 = pow2_scale_exp(W2, 128 * 128), e3 = pow2_scale_exp(W3, (size_t)128 * out_dim);
  const float s2 = ldexpf(1.0f, e2), s3 = ldexpf(1.0f, e3);
  __half* w2hi = reinterpret_cast<__half*>(base);
  __half* w2lo = reinterpret_cast<__half*>(base + TC_W2_BYTES / 2);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) split_store(w2hi, w2lo, sw128_index(128, n, k), W2[k * 128 + n] * s2);
  for (int p = 0; p < lay.n_pass; ++p) {
    __half* hi = reinterpret_cast<__half*>(base + lay.w3 + p * TC_W3_PASS_BYTES);
    __half* lo = reinterpret_cast<__half*>(base + lay.w3 + p * TC_W3_PASS_BYTES + TC_W3_PASS_BYTES / 2);
    for (int j = 0; j < 2; ++j) {
      const int d = 2 * p + j;
      if (d >= st.L) continue;
      for (int q = 0; q < st.cpd; ++q) {
        const int col = d * st.cpd + q, n = 48 * j + q;
        for (int k = 0; k < 128; ++k) split_store(hi, lo, sw128_index(TC_PASS_N, n, k), W3[(size_t)k * out_dim + col] * s3);
      }
    }
  }
  float* par = reinterpret_cast<float*>(base + lay.par);
  for (int i = 0; i < in_dim; ++i)
    for (int k = 0; k < 128; ++k) par[i * 128 + k] = W1[i * 128 + k];
  float* pb1 = par + TC_IN_MAX * 128;
  float* pb2 = pb1 + 128;
  float* pb3 = pb2 + 128;
  for (int k = 0; k < 128; ++k) {
    pb1[k] = b1[k];
    pb2[k] = b2[k];
  }
  for (int d = 0; d < st.L; ++d)
    for (int q = 0; q < st.cpd; ++q) pb3[(d / 2) * TC_PASS_N + 48 * (d % 2) + q] = b3[d * st.cpd + q];
  float* scal = pb3 + 2 * TC_PASS_N;
  scal[0] = ldexpf(1.0f, -e2);
  scal[1] = ldexpf(1.0f, -e3);
}

// ------------------------------------------------------------------------------------------
// device side: PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {  // never suspends the thread
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);  // leave the issue slots to the other group's warps
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  // K-major SWIZZLE_128B operand: start address, SBO = 1024 B (8 rows x 128 B), version 1 (sm100)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fp32 pair -> fp16 hi pair + fp16 lo pair (k even in the low half of each 32-bit column)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// D[128 x N] = A(hi,lo)[128 x 128] * B(hi,lo)[N x 128]^T as lo*hi + hi*lo + hi*hi (24 MMAs)
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_tmem, uint32_t sb_hi, uint32_t sb_lo, int n_rows,
                                           uint32_t idesc) {
  uint32_t acc = 0;
#pragma unroll
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_col = a_tmem + (prod == 0 ? 64 : 0);
    const uint32_t sb = (prod == 1) ? sb_lo : sb_hi;
#pragma unroll
    for (int k16 = 0; k16 < 8; ++k16) {
      const uint64_t bdesc = make_b_desc(sb + (k16 >> 2) * (n_rows * 128) + (k16 & 3) * 32);
      umma_f16_ts(d_tmem, a_col + k16 * 8, bdesc, idesc, acc);
      acc = 1;
    }
  }
}

// softmax knots, bin search, RQS and log-det of one (row, dim) from its 48 raw accumulator values
// (pzflow/bijectors.py:488-491 + pzflow/utils.py:113-227; K = 16).  The bin search runs on the
// un-normalised cumulative sums (knot_j <= x  <=>  cumsum(e)_j <= (x + B) * sum(e) / 2B), so only
// the selected bin's knots and sizes are ever scaled.  The searched axis (widths forward, heights
// inverse) is processed first, then the other axis and the two derivatives the bin needs.
template <bool INV>
__device__ __forceinline__ void spline_regs(const float (&raw)[48], const float* __restrict__ bias, float sinv, float x,
                                            float B, bool periodic, float& out, float& ld, int& idx_out) {
  constexpr float L2E = 1.4426950408889634f;
  constexpr int SO = INV ? 16 : 0, OO = INV ? 0 : 16;
  const float twoB = 2.0f * B;
  const bool oob = (x <= -B) || (x >= B);
  const float mx = oob ? fabsf(x) - B : x;

  // ---- searched axis ----
  float e[16];
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + SO + q);
    e[q + 0] = fmaf(raw[SO + q + 0], sinv, b4.x);
    e[q + 1] = fmaf(raw[SO + q + 1], sinv, b4.y);
    e[q + 2] = fmaf(raw[SO + q + 2], sinv, b4.z);
    e[q + 3] = fmaf(raw[SO + q + 3], sinv, b4.w);
  }
  float m = e[0];
#pragma unroll
  for (int j = 1; j < 16; ++j) m = fmaxf(m, e[j]);
  float nm = -(m * L2E);
  float run = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    e[j] = ex2_approx(fmaf(e[j], L2E, nm));
    run += e[j];
  }
  const float rs = __fdividef(twoB, run);              // 2B / sum(e)
  const float target = __fdividef(mx + B, rs);         // (x + B) * sum(e) / 2B
  int cnt = (-B <= mx) ? 1 : 0;
  float cs_sel = 0.0f, s_sel = e[0];
  run = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    run += e[j];
    const bool p = run <= target;  // knot j+1 <= x: a prefix of trues (cumsum of positives is monotone)
    cnt += p ? 1 : 0;
    cs_sel = p ? run : cs_sel;
    if (j < 15) s_sel = p ? e[j + 1] : s_sel;
  }
  const int idx = cnt - 1;
  idx_out = idx;
  // prefix count over the 16 interior/right knots; the other-axis selections replay it as (j < np)
  const int np = cnt - ((-B <= mx) ? 1 : 0);

  // ---- other axis ----
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + OO + q);
    e[q + 0] = fmaf(raw[OO + q + 0], sinv, b4.x);
    e[q + 1] = fmaf(raw[OO + q + 1], sinv, b4.y);
    e[q + 2] = fmaf(raw[OO + q + 2], sinv, b4.z);
    e[q + 3] = fmaf(raw[OO + q + 3], sinv, b4.w);
  }
  m = e[0];
#pragma unroll
  for (int j = 1; j < 16; ++j) m = fmaxf(m, e[j]);
  nm = -(m * L2E);
  run = 0.0f;
  float co_sel = 0.0f, o_sel = 0.0f;
  float r0 = 0.0f, r1 = fmaf(raw[32], sinv, bias[32]);
  bool pprev = true;  // (j - 1 < np)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float ej = ex2_approx(fmaf(e[j], L2E, nm));
    run += ej;
    o_sel = pprev ? ej : o_sel;  // ends as e[min(np, 15)]
    const bool p = j < np;
    co_sel = p ? run : co_sel;
    if (j < 15) r1 = p ? raw[32 + j + 1] : r1;  // raw[47] is the zero pad unless periodic
    r0 = p ? raw[32 + j] : r0;
    pprev = p;
  }
  const float ro = __fdividef(twoB, run);
  // the two derivative logits: bias applied after the selection (bias index = selected index)
  {
    const int i0 = min(max(np - 1, 0), 15), i1 = min(np, 15);
    if (np > 0) r0 = fmaf(r0, sinv, bias[32 + i0]);
    if (np > 0) r1 = fmaf(r1, sinv, bias[32 + i1]);
  }

  const float qnan = __int_as_float(0x7fc00000);
  const bool bad = (idx < 0) || (idx >= 16);  // take_along_axis "fill" -> NaN (pzflow/utils.py:155-161)
  SplineKnotSel s;
  s.idx = idx;
  const float ksel = fmaf(cs_sel, rs, -B), oknot = fmaf(co_sel, ro, -B);
  const float ssz = bad ? qnan : s_sel * rs, osz = bad ? qnan : o_sel * ro;
  s.wk = INV ? osz : ssz;
  s.hk = INV ? ssz : osz;
  s.xk = INV ? oknot : ksel;
  s.yk = INV ? ksel : oknot;
  if (periodic) {
    if (idx == 0) r0 = fmaf(raw[32 + 15], sinv, bias[32 + 15]);  // dk = pad(D, (1, 0), "wrap")
    s.dk = pz_softplus(r0);
    s.dkp1 = pz_softplus(r1);
  } else {
    s.dk = (idx == 0) ? 1.0f : pz_softplus(r0);
    s.dkp1 = (idx == 15) ? 1.0f : pz_softplus(r1);
  }
  pz_rqs_eval<true>(s, x, mx, oob, INV, periodic, out, ld);
}

struct TcShared {
  uint64_t go[2];        // epilogue group -> MMA thread (256 arrivals): operands ready / accumulator drained
  uint64_t done[2];      // MMA thread -> epilogue group (tcgen05.commit)
  uint64_t lay_done[2];  // epilogue group finished a coupling stage (256 arrivals): its parameter block is free
  uint64_t w2full, w3full, parfull[2];  // TMA complete_tx
  uint64_t w2free, w3free;              // tcgen05.commit after the stage's last hidden GEMM / output pass
  uint32_t tmem_base;
  uint8_t order[MAX_NSC];  // coupling stages in walk order (index into PlanDev::nsc)
};

struct TcSmem {  // byte offsets from the 1024-aligned base
  int w2, w3, par, state, shared, total;
};

__host__ __device__ inline TcSmem tc_smem_layout(int max_pass, int chunk_tiles, int state_cols) {
  TcSmem s;
  s.w2 = 0;
  s.w3 = TC_W2_BYTES;
  s.par = s.w3 + max_pass * TC_W3_PASS_BYTES;
  s.state = s.par + 2 * TC_PAR_BYTES;
  s.shared = s.state + chunk_tiles * TC_TILE * state_cols * 4;
  s.shared = (s.shared + 15) / 16 * 16;
  s.total = s.shared + (int)sizeof(TcShared);
  return s;
}

struct TcLaunch {
  int chunk_tiles, max_pass, state_cols, n_lev;
};

// state column indices: x[0, D), log-det levels [D, D + n_lev), pending log-det halves (2), conditions (C)
__device__ __forceinline__ void apply_affine_cols(const AffineDev& af, bool inverse, int D, float* st, int CR) {
  for (int j = 0; j < D; ++j) {
    const float x = st[j * CR];
    float y;
    if (af.kind == AFF_SHIFT_BOUNDS)
      y = inverse ? (x * af.b[j]) / af.B + af.a[j] : (af.B * (x - af.a[j])) / af.b[j];
    else
      y = inverse ? x * af.b[j] + af.a[j] : (x - af.a[j]) / af.b[j];
    st[j * CR] = y;
  }
}

// Applies the non-NSC steps in [s0, s1) (already mapped to walk order) to one state row.
__device__ __forceinline__ void light_steps_cols(const PlanDev* plan, bool inverse, int s0, int s1, int& level, float* st,
                                                 int CR) {
  const int D = plan->D, ns = plan->n_steps;
  float* ldl = st + D * CR;
  for (int si = s0; si < s1; ++si) {
    const StepDev step = plan->steps[inverse ? ns - 1 - si : si];
    int kind = step.kind;
    if (inverse && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
    else if (inverse && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
    if (kind == STEP_CHAIN_BEGIN) {
      ++level;
      ldl[level * CR] = 0.0f;
    } else if (kind == STEP_CHAIN_END) {
      ldl[(level - 1) * CR] += ldl[level * CR];
      --level;
    } else if (kind == STEP_AFFINE) {
      const AffineDev& af = plan->affine[step.idx];
      apply_affine_cols(af, inverse, D, st, CR);
      ldl[level * CR] += inverse ? af.logdet_inv : af.logdet_fwd;
    }
  }
}

__device__ __forceinline__ int level_after(const PlanDev* plan, bool inverse, int s0, int s1, int level) {
  const int ns = plan->n_steps;
  for (int si = s0; si < s1; ++si) {
    int kind = plan->steps[inverse ? ns - 1 - si : si].kind;
    if (inverse && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
    else if (inverse && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
    level += (kind == STEP_CHAIN_BEGIN) - (kind == STEP_CHAIN_END);
  }
  return level;
}

// E1 for one half row: 64 hidden columns of LeakyReLU(key @ W1 + b1) -> fp16 hi/lo in TMEM
template <int IN>
__device__ __forceinline__ void e1_half(const float (&key)[TC_IN_MAX], const float* __restrict__ sW1,
                                        const float* __restrict__ sb1, int half, uint32_t t_a) {
#pragma unroll
  for (int bt = 0; bt < 2; ++bt) {
    uint32_t rh[16], rl[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k0 = half * 64 + bt * 32 + q * 4;
      float4 acc = *reinterpret_cast<const float4*>(sb1 + k0);
#pragma unroll
      for (int i = 0; i < IN; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(sW1 + i * 128 + k0);
        acc.x = fmaf(key[i], w.x, acc.x);
        acc.y = fmaf(key[i], w.y, acc.y);
        acc.z = fmaf(key[i], w.z, acc.z);
        acc.w = fmaf(key[i], w.w, acc.w);
      }
      split2(pz_leaky(acc.x), pz_leaky(acc.y), rh[2 * q], rl[2 * q]);
      split2(pz_leaky(acc.z), pz_leaky(acc.w), rh[2 * q + 1], rl[2 * q + 1]);
    }
    tmem_st16(t_a + half * 32 + bt * 16, rh);
    tmem_st16(t_a + 64 + half * 32 + bt * 16, rl);
  }
}

template <bool INV>
__global__ void __launch_bounds__(TC_THREADS, 1)
nsf_chain_tc_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, TcLaunch lp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, derived by pointer arithmetic so loads stay LDS (not generic LD)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const TcSmem sl = tc_smem_layout(lp.max_pass, lp.chunk_tiles, lp.state_cols);
  TcShared* sh = reinterpret_cast<TcShared*>(base + sl.shared);
  float* state = reinterpret_cast<float*>(base + sl.state);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = plan->D, C = plan->C, ns = plan->n_steps, n_nsc = plan->n_nsc, n_lev = lp.n_lev;
  const int CR = lp.chunk_tiles * TC_TILE;           // rows of state per column
  const int COL_PEND = D + n_lev, COL_COND = D + n_lev + 2;
  const bool is_epi = warp < 16;
  const int grp = (warp >> 3) & 1;                   // epilogue group
  const int half = (warp >> 2) & 1;                  // which half of the row's work
  const int lrow = ((warp & 3) << 5) + lane;         // row within the tile == TMEM lane

  if (tid == 0) {
    for (int g = 0; g < 2; ++g) {
      mbar_init(smem_u32(&sh->go[g]), TC_GROUP_THREADS);
      mbar_init(smem_u32(&sh->done[g]), 1);
      mbar_init(smem_u32(&sh->lay_done[g]), TC_GROUP_THREADS);
      mbar_init(smem_u32(&sh->parfull[g]), 1);
    }
    mbar_init(smem_u32(&sh->w2full), 1);
    mbar_init(smem_u32(&sh->w3full), 1);
    mbar_init(smem_u32(&sh->w2free), 1);
    mbar_init(smem_u32(&sh->w3free), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int seen = 0;
    for (int si = 0; si < plan->n_steps; ++si) {
      const StepDev step = plan->steps[INV ? plan->n_steps - 1 - si : si];
      if (step.kind == STEP_NSC) sh->order[seen++] = (uint8_t)step.idx;
    }
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sh->tmem_base;

  const long long chunk_rows = (long long)CR;
  const long long n_chunks = (args.N + chunk_rows - 1) / chunk_rows;
  const int my_chunks = (int)((n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x);  // blockIdx.x < n_chunks

  // NSC steps in walk order
  int first_nsc = ns, last_nsc = -1;
  for (int si = 0; si < ns; ++si)
    if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
      if (first_nsc == ns) first_nsc = si;
      last_nsc = si;
    }

  if (is_epi) {
    // ======================================= epilogue warps =======================================
    const uint32_t t_lane = tbase + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t t_a = t_lane + (uint32_t)grp * 256u, t_d = t_lane + (uint32_t)grp * 256u + 128u;
    const uint32_t bar_go = smem_u32(&sh->go[grp]), bar_done = smem_u32(&sh->done[grp]);
    uint32_t ph_done = 0u;
    int seq = 0;  // coupling stages processed so far by this CTA (over all its chunks)

    for (int ci = 0; ci < my_chunks; ++ci) {
      const long long row0 = ((long long)blockIdx.x + (long long)ci * gridDim.x) * chunk_rows;
      const int nrows = (int)min(chunk_rows, args.N - row0);
      const int ntiles = max(2, (nrows + TC_TILE - 1) / TC_TILE);
      const int my_tiles = (ntiles - grp + 1) / 2;

      // ---- pass A: rows -> state, steps in front of the first coupling stage ----
      for (int r = tid; r < nrows; r += TC_EPI_THREADS) {
        const long long row = row0 + r;
        float* st = state + r;
        for (int j = 0; j < D; ++j) {
          float v;
          if (args.mode == MODE_POSTERIOR) {
            const long long g = row / args.G;
            const int gi = (int)(row - g * args.G);
            v = (j == args.col_idx) ? args.grid[gi] : args.X[g * (D - 1) + (j < args.col_idx ? j : j - 1)];
          } else {
            v = args.X[row * D + j];
          }
          st[(INV ? plan->perm_final[j] : j) * CR] = v;
        }
        for (int l = 0; l < n_lev; ++l) st[(D + l) * CR] = 0.0f;
        const long long crow = (args.mode == MODE_POSTERIOR) ? row / args.G : row;
        for (int j = 0; j < C; ++j) st[(COL_COND + j) * CR] = args.cond ? args.cond[crow * C + j] : 0.0f;
        int level = 0;
        light_steps_cols(plan, INV, 0, first_nsc, level, st, CR);
      }
      int level = level_after(plan, INV, 0, first_nsc, 0);  // chain depth (row independent)
      named_bar_sync(1, TC_EPI_THREADS);

      // ---- the coupling stages ----
      int prev_end = first_nsc;
      bool pending = false;  // the previous stage's two log-det halves still sit in the pending columns
      int pend_level = 0;
      for (int si = first_nsc; si <= last_nsc; ++si) {
        const StepDev step = plan->steps[INV ? ns - 1 - si : si];
        if (step.kind != STEP_NSC) continue;
        const bool has_light = si > prev_end;
        const int light0 = prev_end;
        const int level_in = level;
        level = level_after(plan, INV, prev_end, si, level);
        prev_end = si + 1;
        const NscDev& st = plan->nsc[step.idx];
        const int U = st.U, L = st.L, in_dim = st.U + st.C, n_pass = (L + 1) / 2;
        const float B = st.B;
        const bool periodic = st.periodic != 0;

        // parameter block of this stage (double buffered); previous-stage state writes of the row's other half
        mbar_wait(smem_u32(&sh->parfull[seq & 1]), (uint32_t)((seq >> 1) & 1));
        named_bar_sync(2 + grp, TC_GROUP_THREADS);
        const float* par = reinterpret_cast<const float*>(base + sl.par + (seq & 1) * TC_PAR_BYTES);
        const float* sW1 = par;
        const float* sb1 = par + TC_IN_MAX * 128;
        const float* sb2 = sb1 + 128;
        const float* sb3 = sb2 + 128;
        const float s2inv = sb3[2 * TC_PASS_N], s3inv = sb3[2 * TC_PASS_N + 1];

        for (int k = 0; k < my_tiles; ++k) {
          const int r = (2 * k + grp) * TC_TILE + lrow;  // row within the chunk (state is allocated for every tile)
          const bool valid = r < nrows;
          float* srow = state + r;

          if (pending || has_light) {
            if (half == 0) {
              if (pending) srow[(D + pend_level) * CR] += srow[COL_PEND * CR] + srow[(COL_PEND + 1) * CR];
              if (has_light) {
                int lv = level_in;
                light_steps_cols(plan, INV, light0, si, lv, srow, CR);
              }
            }
            if (has_light) named_bar_sync(2 + grp, TC_GROUP_THREADS);
          }

          float key[TC_IN_MAX];
#pragma unroll
          for (int i = 0; i < TC_IN_MAX; ++i)
            key[i] = (i < in_dim) ? (i < U ? srow[st.perm[i] * CR] : srow[(COL_COND + (i - U)) * CR]) : 0.0f;

          // ---- E1: h1 = LeakyReLU(key @ W1 + b1) -> A (fp16 hi/lo) in TMEM ----
          switch (in_dim) {
            case 1: e1_half<1>(key, sW1, sb1, half, t_a); break;
            case 2: e1_half<2>(key, sW1, sb1, half, t_a); break;
            case 3: e1_half<3>(key, sW1, sb1, half, t_a); break;
            case 4: e1_half<4>(key, sW1, sb1, half, t_a); break;
            case 5: e1_half<5>(key, sW1, sb1, half, t_a); break;
            case 6: e1_half<6>(key, sW1, sb1, half, t_a); break;
            case 7: e1_half<7>(key, sW1, sb1, half, t_a); break;
            default: e1_half<8>(key, sW1, sb1, half, t_a); break;
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_go);

          // ---- E2: h2 = LeakyReLU(D * 2^-e2 + b2) -> A ----
          mbar_wait(bar_done, ph_done);
          ph_done ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int bt = 0; bt < 2; ++bt) {
            float dv[32];
            tmem_ld16(t_d + half * 64 + bt * 32, dv);
            tmem_ld16(t_d + half * 64 + bt * 32 + 16, dv + 16);
            tc_wait_ld();
            uint32_t rh[16], rl[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = *reinterpret_cast<const float4*>(sb2 + half * 64 + bt * 32 + q * 4);
              const float h0 = pz_leaky(fmaf(dv[4 * q + 0], s2inv, bb.x));
              const float h1 = pz_leaky(fmaf(dv[4 * q + 1], s2inv, bb.y));
              const float h2 = pz_leaky(fmaf(dv[4 * q + 2], s2inv, bb.z));
              const float h3 = pz_leaky(fmaf(dv[4 * q + 3], s2inv, bb.w));
              split2(h0, h1, rh[2 * q], rl[2 * q]);
              split2(h2, h3, rh[2 * q + 1], rl[2 * q + 1]);
            }
            tmem_st16(t_a + half * 32 + bt * 16, rh);
            tmem_st16(t_a + 64 + half * 32 + bt * 16, rl);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_go);

          // ---- E3: logits -> knots -> spline; this thread takes dim 2p + half of pass p ----
          float ldsum = 0.0f;
          for (int p = 0; p < n_pass; ++p) {
            mbar_wait(bar_done, ph_done);
            ph_done ^= 1u;
            tc_fence_after();
            const int d = 2 * p + half;
            const bool mine = d < L;  // warp-uniform
            float raw[48];
            if (mine) {
              tmem_ld16(t_d + 48 * half, raw);
              tmem_ld16(t_d + 48 * half + 16, raw + 16);
              tmem_ld16(t_d + 48 * half + 32, raw + 32);
              tc_wait_ld();
            }
            if (p + 1 < n_pass) {  // accumulator drained: the next pass may overwrite it
              tc_fence_before();
              mbar_arrive(bar_go);
            }
            if (mine) {
              const int col = st.perm[U + d];
              const float xin = srow[col * CR];
              float y, ldd;
              int idx;
              spline_regs<INV>(raw, sb3 + p * TC_PASS_N + 48 * half, s3inv, xin, B, periodic, y, ldd, idx);
              ldsum += ldd;
              srow[col * CR] = y;
              if (valid && args.bins != nullptr)
                args.bins[(row0 + r) * plan->n_spline_dims + st.bins_off + d] = idx;
            }
          }
          srow[(COL_PEND + half) * CR] = INV ? -ldsum : ldsum;
        }
        mbar_arrive(smem_u32(&sh->lay_done[grp]));
        pending = true;
        pend_level = level;
        ++seq;
      }
      named_bar_sync(1, TC_EPI_THREADS);

      // ---- pass Z: trailing steps, latent log-density, outputs ----
      for (int r = tid; r < nrows; r += TC_EPI_THREADS) {
        const long long row = row0 + r;
        float* st = state + r;
        if (pending) st[(D + pend_level) * CR] += st[COL_PEND * CR] + st[(COL_PEND + 1) * CR];
        int lv = level;
        light_steps_cols(plan, INV, prev_end, ns, lv, st, CR);
        if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
          const float lat = pz_latent_logprob(plan->latent_kind, D, plan->latent_B, plan->latent_const,
                                              [&](int j) { return st[plan->perm_final[j] * CR]; });
          const float lpv = pz_nan_to_num_lp(lat + st[D * CR]);
          args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
        } else {
          for (int j = 0; j < D; ++j) args.out0[row * D + j] = st[(INV ? j : plan->perm_final[j]) * CR];
          if (args.out1 != nullptr) args.out1[row] = st[D * CR];
        }
      }
      // (pass A of the next chunk touches a row from the same thread as pass Z did: no barrier needed)
    }
  } else if (lane == 0) {
    // =============================== MMA issuer + weight streamer ===============================
    const uint32_t sb = smem_u32(base);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc3 = (1u << 4) | ((uint32_t)(TC_PASS_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int total_seq = my_chunks * n_nsc;

    // coupling stage (walk order) of sequence number s, and the tile count of its chunk
    auto stage_of = [&](int s) -> const NscDev& { return plan->nsc[sh->order[s % n_nsc]]; };
    auto tiles_of = [&](int s) -> int {
      const long long row0 = ((long long)blockIdx.x + (long long)(s / n_nsc) * gridDim.x) * chunk_rows;
      const int nrows = (int)min(chunk_rows, args.N - row0);
      return max(2, (nrows + TC_TILE - 1) / TC_TILE);
    };

    // per group: sequence number, tile, phase (0 = hidden GEMM, p + 1 = output pass p) and the cached
    // shape of its current stage
    int gs[2] = {0, 0}, gk[2] = {0, 0}, gp[2] = {0, 0}, gnp[2], gnt[2];
    uint32_t gpar[2] = {0u, 0u};
    bool gready[2] = {false, false};
    gnp[0] = gnp[1] = total_seq > 0 ? (stage_of(0).L + 1) / 2 : 1;
    gnt[0] = gnt[1] = total_seq > 0 ? tiles_of(0) : 2;
    int w2_next = 0, w3_next = 0, par_next = 0;      // next sequence number to load
    int w2_ready = 0, w3_ready = 0;                  // loads observed complete
    int w2_freed = 0, w3_freed = 0;                  // buffer releases observed
    int lay_seen[2] = {0, 0};                        // stages each group has finished
    int c2_cnt = 0, c3_cnt = 0;                      // GEMMs issued in the stage currently holding W2 / W3

    while (gs[0] < total_seq || gs[1] < total_seq) {
      bool progress = false;
      // ---- weight streaming ----
      if (w2_next < total_seq) {
        if (w2_next > w2_freed && mbar_test_wait(smem_u32(&sh->w2free), (uint32_t)(w2_freed & 1))) ++w2_freed;
        if (w2_next <= w2_freed) {
          const NscDev& st = stage_of(w2_next);
          const uint32_t bar = smem_u32(&sh->w2full);
          mbar_expect_tx(bar, TC_W2_BYTES);
          const uint8_t* src = static_cast<const uint8_t*>(st.tc.blob);
          for (int off = 0; off < TC_W2_BYTES; off += 16384) tma_bulk_g2s(sb + sl.w2 + off, src + off, 16384u, bar);
          ++w2_next;
          progress = true;
        }
      }
      if (w3_next < total_seq) {
        if (w3_next > w3_freed && mbar_test_wait(smem_u32(&sh->w3free), (uint32_t)(w3_freed & 1))) ++w3_freed;
        if (w3_next <= w3_freed) {
          const NscDev& st = stage_of(w3_next);
          const TcLayout lay = TcLayout::make(st.L);
          const uint32_t bar = smem_u32(&sh->w3full);
          const int nb = lay.n_pass * TC_W3_PASS_BYTES;
          mbar_expect_tx(bar, (uint32_t)nb);
          const uint8_t* src = static_cast<const uint8_t*>(st.tc.blob) + lay.w3;
          for (int off = 0; off < nb; off += 16384) tma_bulk_g2s(sb + sl.w3 + off, src + off, 16384u, bar);
          ++w3_next;
          progress = true;
        }
      }
      if (par_next < total_seq) {
        // block par_next reuses the buffer of stage par_next - 2: both groups must have finished it
        for (int g = 0; g < 2; ++g)
          if (lay_seen[g] < par_next - 1 && mbar_test_wait(smem_u32(&sh->lay_done[g]), (uint32_t)(lay_seen[g] & 1)))
            ++lay_seen[g];
        if (par_next < 2 || (lay_seen[0] >= par_next - 1 && lay_seen[1] >= par_next - 1)) {
          const NscDev& st = stage_of(par_next);
          const TcLayout lay = TcLayout::make(st.L);
          const uint32_t bar = smem_u32(&sh->parfull[par_next & 1]);
          mbar_expect_tx(bar, TC_PAR_BYTES);
          tma_bulk_g2s(sb + sl.par + (par_next & 1) * TC_PAR_BYTES, static_cast<const uint8_t*>(st.tc.blob) + lay.par,
                       TC_PAR_BYTES, bar);
          ++par_next;
          progress = true;
        }
      }
      // ---- GEMMs: serve whichever group is ready ----
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (gs[g] >= total_seq) continue;
        if (!gready[g]) {
          if (!mbar_test_wait(smem_u32(&sh->go[g]), gpar[g])) continue;
          gpar[g] ^= 1u;
          gready[g] = true;
        }
        const int s = gs[g];
        if (gp[g] == 0) {
          if (w2_ready <= s) {
            if (w2_next > s && mbar_test_wait(smem_u32(&sh->w2full), (uint32_t)(s & 1))) w2_ready = s + 1;
            else continue;
          }
        } else {
          if (w3_ready <= s) {
            if (w3_next > s && mbar_test_wait(smem_u32(&sh->w3full), (uint32_t)(s & 1))) w3_ready = s + 1;
            else continue;
          }
        }
        tc_fence_after();
        const uint32_t a_t = tbase + g * 256, d_t = tbase + g * 256 + 128;
        if (gp[g] == 0) {
          issue_gemm(d_t, a_t, sb + sl.w2, sb + sl.w2 + TC_W2_BYTES / 2, 128, idesc2);
          umma_commit(smem_u32(&sh->done[g]));
          if (++c2_cnt == gnt[g]) {  // every hidden GEMM of the stage is issued: release W2 when they complete
            umma_commit(smem_u32(&sh->w2free));
            c2_cnt = 0;
          }
        } else {
          const uint32_t w3 = sb + sl.w3 + (gp[g] - 1) * TC_W3_PASS_BYTES;
          issue_gemm(d_t, a_t, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N, idesc3);
          umma_commit(smem_u32(&sh->done[g]));
          if (++c3_cnt == gnt[g] * gnp[g]) {
            umma_commit(smem_u32(&sh->w3free));
            c3_cnt = 0;
          }
        }
        gready[g] = false;
        progress = true;
        if (++gp[g] > gnp[g]) {
          gp[g] = 0;
          if (++gk[g] == (gnt[g] - g + 1) / 2) {
            gk[g] = 0;
            if (++gs[g] < total_seq) {
              gnp[g] = (stage_of(gs[g]).L + 1) / 2;
              gnt[g] = tiles_of(gs[g]);
            }
          }
        }
      }
      if (!progress) __nanosleep(20);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

cudaError_t launch_tc(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms, void* workspace,
                      cudaStream_t stream) {
  (void)workspace;
  TcLaunch lp;
  lp.max_pass = 1;
  for (int s = 0; s < hp.n_nsc; ++s) lp.max_pass = max(lp.max_pass, (hp.nsc[s].L + 1) / 2);
  // chain nesting depth actually used (log-det levels)
  int depth = 0, level = 0;
  for (int i = 0; i < hp.n_steps; ++i) {
    if (hp.steps[i].kind == STEP_CHAIN_BEGIN) depth = max(depth, ++level);
    else if (hp.steps[i].kind == STEP_CHAIN_END) --level;
  }
  lp.n_lev = depth + 1;
  lp.state_cols = hp.D + lp.n_lev + 2 + hp.C;
  // rows per chunk: as many 128-row tiles as shared memory holds, fewer when that leaves SMs idle
  int fit = TC_MAX_CHUNK_TILES;
  while (fit > 2 && tc_smem_layout(lp.max_pass, fit, lp.state_cols).total + 1024 > TC_SMEM_MAX) --fit;
  const long long tiles_total = (args.N + TC_TILE - 1) / TC_TILE;
  long long want = (tiles_total + num_sms - 1) / num_sms;
  if (want < 2) want = 2;
  lp.chunk_tiles = (int)(want < fit ? want : fit);
  if (lp.chunk_tiles > 8 && want > fit) {
    // balance: the largest chunk size <= fit that leaves the last wave as full as possible is not worth
    // a search here; 8 tiles keep the per-stage weight reload below 2 % of the stage's math
    lp.chunk_tiles = fit;
  }
  const TcSmem sl = tc_smem_layout(lp.max_pass, lp.chunk_tiles, lp.state_cols);
  const size_t smem = (size_t)sl.total + 1024;
  if (smem > (size_t)TC_SMEM_MAX) return cudaErrorInvalidConfiguration;
  const bool inv = (args.mode == MODE_INVERSE);
  cudaError_t e = inv ? cudaFuncSetAttribute(nsf_chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(nsf_chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long chunk_rows = (long long)lp.chunk_tiles * TC_TILE;
  const long long n_chunks = (args.N + chunk_rows - 1) / chunk_rows;
  long long grid = num_sms;
  if (grid > n_chunks) grid = n_chunks;
  if (grid < 1) grid = 1;
  if (inv)
    nsf_chain_tc_kernel<true><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp);
  else
    nsf_chain_tc_kernel<false><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp);
  return cudaGetLastError();
}

}  // namespace pzb
