// Pieces shared by the two tcgen05 engines (nsf_tc.cu: the default conditioner family; nsf_wide.cu: wide / deep
// conditioners): PTX wrappers (mbarrier, TMA bulk copies, TMEM loads / stores, tcgen05.mma with the A operand in TMEM),
// the fp16 hi/lo split, the fast spline arithmetic and the non-coupling ("light") steps on a column-major state tile.
#pragma once
#include <cuda_fp16.h>

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "pzb_common.cuh"

namespace pzb {

// chain nesting depth actually used (number of log-det levels kept per row)
static inline int tc_log_det_levels(const PlanDev& hp) {
  int depth = 0, level = 0;
  for (int i = 0; i < hp.n_steps; ++i) {
    if (hp.steps[i].kind == STEP_CHAIN_BEGIN) depth = max(depth, ++level);
    else if (hp.steps[i].kind == STEP_CHAIN_END) --level;
  }
  return depth + 1;
}

static inline int pow2_scale_exp(const float* w, size_t n) {
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

// element (n, k) of the K-major SWIZZLE_NONE [128 x 16] tile: 8 x 8 core matrices of 128 contiguous bytes,
// K-adjacent cores 128 B apart (LBO), 8-row groups 256 B apart (SBO)
static inline size_t w1_index(int n, int k) { return (size_t)(n >> 3) * 128 + (size_t)(k >> 3) * 64 + (size_t)(n & 7) * 8 + (k & 7); }

static inline void split_store(__half* hi, __half* lo, size_t idx, float v) {
  const __half h = __float2half_rn(v);
  hi[idx] = h;
  lo[idx] = __float2half_rn(v - __half2float(h));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Blocks until the phase with the given parity has completed.  try_wait suspends the thread in hardware
// (up to the time hint) instead of polling, so waiting warps leave the issue slots to working ones.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "PZB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra PZB_DONE;\n\t"
      "bra PZB_WAIT;\n\t"
      "PZB_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival per warp (lane 0, after the warp has converged): 32x fewer mbarrier events than per-thread arrivals.
// Every arrival wakes all NANOSLEEP.SYNCS sleepers of the SM, so fewer arrivals also mean fewer futile wake-ups.
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t u[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  // K-major SWIZZLE_128B operand: start address, SBO = 1024 B (8 rows x 128 B), version 1 (sm100)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t make_b_desc_w1(uint32_t saddr) {
  // K-major SWIZZLE_NONE operand: LBO = 128 B between K-adjacent core matrices, SBO = 256 B between 8-row groups
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// One lane of a converged warp.  The region it guards stays on the uniform datapath (operands of
// tcgen05.mma / commit in uniform registers), unlike a `lane == 0` branch, which ptxas treats as divergent
// and pays ~10 instructions per MMA for (R2UR of every operand + its own ELECT).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fp32 pair -> fp16 hi pair + fp16 lo pair (k even in the low half of each 32-bit column)
// The residual a - float(hi) is one mixed-precision FMA per element (fma.rn.f32.f16 -> FHFMA: hi * -1 + a,
// exact), so a pair costs F2FP + 2 FHFMA + F2FP.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16), m1 = 0xBC00;  // -1.0
  float r0, r1;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r0) : "h"(h0), "h"(m1), "f"(a));
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r1) : "h"(h1), "h"(m1), "f"(b));
  const __half2 l = __floats2half2_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Two LeakyReLUs, max(x, 0.01 x) each (pz_leaky): the two products are ONE packed fp32x2 multiply (sm_100 FMUL2),
// bit-identical to two FMULs.
__device__ __forceinline__ void leaky2(float a, float b, float& oa, float& ob) {
  const float2 t = __fmul2_rn(make_float2(a, b), make_float2(0.01f, 0.01f));
  oa = fmaxf(a, t.x);
  ob = fmaxf(b, t.y);
}

// A operand layout in a 128-column TMEM region (K = 128 fp16 hi/lo): the 32 elements [32b, 32b+32)
// occupy columns [32b, 32b+32): 16 hi columns then 16 lo columns (two elements per 32-bit column), so an
// epilogue thread can rewrite 32 fp32 accumulator columns in place.  K-step k16 (16 elements):
__device__ __forceinline__ constexpr uint32_t a_col_of(int k16) { return (uint32_t)((k16 >> 1) * 32 + (k16 & 1) * 8); }

// D[128 x N] (+)= A(hi,lo)[128 x 16 K] * B(hi,lo)[N x 16 K]^T over the K-steps [K0, K1) of 16 elements, as
// lo*hi + hi*lo + hi*hi (3 MMAs per K-step).  FRESH: the first MMA overwrites the accumulator.
// sb_hi / sb_lo: shared addresses of the first B row of the (sub-)tile; kblk: bytes between its two K blocks.
template <int K0, int K1, bool FRESH>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_tmem, uint32_t sb_hi, uint32_t sb_lo, uint32_t kblk,
                                           uint32_t idesc) {
  const uint64_t dh = make_b_desc(sb_hi), dl = make_b_desc(sb_lo);
  const uint32_t kb16 = kblk >> 4;
#pragma unroll
  for (int prod = 0; prod < 3; ++prod) {
    const uint32_t a_off = (prod == 0) ? 16u : 0u;  // lo * hi first
    const uint64_t db = (prod == 1) ? dl : dh;
#pragma unroll
    for (int k16 = K0; k16 < K1; ++k16) {
      const uint64_t bdesc = db + (uint64_t)((k16 >> 2) * kb16 + (k16 & 3) * 2);
      if (FRESH && prod == 0 && k16 == K0)
        umma_f16_ts<false>(d_tmem, a_tmem + a_off + a_col_of(k16), bdesc, idesc);
      else
        umma_f16_ts<true>(d_tmem, a_tmem + a_off + a_col_of(k16), bdesc, idesc);
    }
  }
}

// jax.nn.softplus = logaddexp(x, 0) = max(x, 0) + log1p(exp(-|x|)), with log1p(t) = 2 atanh(t / (2 + t)):
// t in (0, 1] keeps s = t / (2 + t) <= 1/3, where the odd series through s^13 is exact to 1.4e-8 relative.
__device__ __forceinline__ float tc_softplus(float x) {
  const float t = ex2_approx(fabsf(x) * -1.4426950408889634f);
  const float s = __fdividef(t, 2.0f + t);
  const float z = s * s;
  float p = fmaf(z, 0.07692307692307693f, 0.09090909090909091f);
  p = fmaf(z, p, 0.1111111111111111f);
  p = fmaf(z, p, 0.14285714285714285f);
  p = fmaf(z, p, 0.2f);
  p = fmaf(z, p, 0.3333333333333333f);
  p = fmaf(z, p, 1.0f);
  return fmaf(s + s, p, fmaxf(x, 0.0f));
}

// pzflow/utils.py:163-227 once the bin's seven values are known (same op order as pz_rqs_eval), with the
// three quotients as MUFU.RCP * x (2 ulp) and the logarithms as lg2.approx.
__device__ __forceinline__ void tc_rqs_eval(const SplineKnotSel& s, float x, float mx, bool oob, bool inverse,
                                            bool periodic, float& out, float& ld) {
  const float sk = __fdividef(s.hk, s.wk);
  const float dsum = (s.dkp1 + s.dk) - 2.0f * sk;
  float relx;
  if (inverse) {
    const float t = mx - s.yk;
    const float a = s.hk * (sk - s.dk) + t * dsum;
    const float b = s.hk * s.dk - t * dsum;
    const float c = (-sk) * t;
    float root;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(root) : "f"(b * b - (4.0f * a) * c));
    relx = __fdividef(2.0f * c, (-b) - root);
    out = relx * s.wk + s.xk;
  } else {
    relx = __fdividef(mx - s.xk, s.wk);
    const float num = s.hk * (sk * (relx * relx) + (s.dk * relx) * (1.0f - relx));
    const float den = sk + (dsum * relx) * (1.0f - relx);
    out = s.yk + __fdividef(num, den);
  }
  const float omr = 1.0f - relx;
  const float dnum = (s.dkp1 * (relx * relx) + ((2.0f * sk) * relx) * omr) + s.dk * (omr * omr);
  const float dden = sk + (dsum * relx) * omr;
  ld = 0.6931471805599453f * ((2.0f * pz_lg2_approx(sk) + pz_lg2_approx(dnum)) - 2.0f * pz_lg2_approx(dden));
  if (!periodic && oob) {
    out = x;
    ld = 0.0f;
  }
}

// state column indices: x[0, D), log-det levels [D, D + n_lev), pending log-det halves (2), conditions (C)
__device__ __forceinline__ float apply_affine_cols(const AffineDev& af, bool inverse, int D, float* st, int CR) {
  float ldsum = 0.0f;
  for (int j = 0; j < D; ++j) {
    float ld;
    st[j * CR] = pz_affine_elem(af, inverse, j, st[j * CR], ld);
    ldsum += ld;
  }
  return ldsum;
}

// Applies the non-NSC steps in [s0, s1) (already mapped to walk order) to one state row.
static __device__ __noinline__ void light_steps_cols(const PlanDev* plan, bool inverse, int s0, int s1, int& level, float* st,
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
      const float ld = apply_affine_cols(af, inverse, D, st, CR);
      ldl[level * CR] += (af.kind == AFF_INV_SOFTPLUS) ? ld : (inverse ? af.logdet_inv : af.logdet_fwd);
    } else if (kind == STEP_COLOR) {
      pz_color_row(plan->color[step.idx], inverse, st, CR);
    }
  }
}

static __device__ __noinline__ int level_after(const PlanDev* plan, bool inverse, int s0, int s1, int level) {
  const int ns = plan->n_steps;
  for (int si = s0; si < s1; ++si) {
    int kind = plan->steps[inverse ? ns - 1 - si : si].kind;
    if (inverse && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
    else if (inverse && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
    level += (kind == STEP_CHAIN_BEGIN) - (kind == STEP_CHAIN_END);
  }
  return level;
}

// 16 un-normalised softmax terms of one axis: e_j = 2^((l_j - max l) * log2 e), l = raw * sinv + bias
__device__ __forceinline__ void softmax_terms(const float* __restrict__ raw, const float* __restrict__ bias, float sinv,
                                              float (&e)[16]) {
  constexpr float L2E = 1.4426950408889634f;
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + q);
    e[q + 0] = fmaf(raw[q + 0], sinv, b4.x);
    e[q + 1] = fmaf(raw[q + 1], sinv, b4.y);
    e[q + 2] = fmaf(raw[q + 2], sinv, b4.z);
    e[q + 3] = fmaf(raw[q + 3], sinv, b4.w);
  }
  float m = fmaxf(e[0], e[1]);
#pragma unroll
  for (int j = 2; j < 16; j += 2) m = fmaxf(m, fmaxf(e[j], e[j + 1]));
  const float nm = -(m * L2E);
#pragma unroll
  for (int j = 0; j < 16; ++j) e[j] = ex2_approx(fmaf(e[j], L2E, nm));
}

// One bin of a 16-term cumulative sum, found in two levels of four: sums of the four groups, then the
// prefix inside the selected group.  Returns the prefix at the bin's left edge and the bin's own term.
struct BinPick {
  bool p1, p2, p3, q1, q2, q3;  // group / in-group "right edge <= target" flags (each a prefix of trues)
};
__device__ __forceinline__ void group_sums(const float (&e)[16], float& G1, float& G2, float& G3, float& tot) {
  const float g0 = (e[0] + e[1]) + (e[2] + e[3]), g1 = (e[4] + e[5]) + (e[6] + e[7]);
  const float g2 = (e[8] + e[9]) + (e[10] + e[11]), g3 = (e[12] + e[13]) + (e[14] + e[15]);
  G1 = g0;
  G2 = G1 + g1;
  G3 = G2 + g2;
  tot = G3 + g3;
}
template <bool SEARCH>
__device__ __forceinline__ void pick_bin(const float (&e)[16], float G1, float G2, float G3, float target, BinPick& k,
                                         float& left, float& term) {
  if (SEARCH) {
    k.p1 = G1 <= target;
    k.p2 = G2 <= target;
    k.p3 = G3 <= target;
  }
  const float base = k.p3 ? G3 : (k.p2 ? G2 : (k.p1 ? G1 : 0.0f));
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = k.p3 ? e[12 + i] : (k.p2 ? e[8 + i] : (k.p1 ? e[4 + i] : e[i]));
  const float c1 = base + a[0], c2 = c1 + a[1], c3 = c2 + a[2];
  if (SEARCH) {
    k.q1 = c1 <= target;
    k.q2 = c2 <= target;
    k.q3 = c3 <= target;
  }
  left = k.q3 ? c3 : (k.q2 ? c2 : (k.q1 ? c1 : base));
  term = k.q3 ? a[3] : (k.q2 ? a[2] : (k.q1 ? a[1] : a[0]));
}

// softmax knots, bin search, RQS and log-det of one (row, dim) from its 48 raw accumulator values
// (pzflow/bijectors.py:488-491 + pzflow/utils.py:113-227; K = 16).  The bin search runs on the
// un-normalised cumulative sums (knot_j <= x  <=>  cumsum(e)_j <= (x + B) * sum(e) / 2B), so only the
// selected bin's knots and sizes are ever scaled.  The searched axis (widths forward, heights inverse)
// is processed first, then the other axis and the two derivatives the bin needs.
template <bool INV>
__device__ __forceinline__ void spline_dim(const float (&raw)[48], const float* __restrict__ bias, float sinv, float x,
                                           float B, float inv2B, bool periodic, float& out, float& ld, int& idx_out,
                                           bool& finite) {
  constexpr int SO = INV ? 16 : 0, OO = INV ? 0 : 16;
  const float twoB = 2.0f * B;
  const bool oob = (x <= -B) || (x >= B);
  const float mx = oob ? fabsf(x) - B : x;
  BinPick k;
  // both axes' softmax terms and group sums up front: two independent dependency chains for the scheduler
  float e[16], f[16], G1, G2, G3, tot, H1, H2, H3, tot2;
  softmax_terms(raw + SO, bias + SO, sinv, e);
  softmax_terms(raw + OO, bias + OO, sinv, f);
  group_sums(e, G1, G2, G3, tot);
  group_sums(f, H1, H2, H3, tot2);
  const float target = (mx + B) * (tot * inv2B);  // (x + B) * sum(e) / 2B
  finite = (tot + tot2) < INFINITY;  // false for NaN too: non-finite logits (an fp16 overflow upstream, or non-finite inputs)
  float cs_sel, s_sel;
  pick_bin<true>(e, G1, G2, G3, target, k, cs_sel, s_sel);
  const float rs = __fdividef(twoB, tot);
  // knots <= x: the left edge -B, then a prefix of the 16 cumulative sums (pzflow/utils.py:150-152)
  const int gi = k.p3 ? 3 : (k.p2 ? 2 : (k.p1 ? 1 : 0)), wi = k.q3 ? 3 : (k.q2 ? 2 : (k.q1 ? 1 : 0));
  int idx = 4 * gi + wi;
  if (tot <= target) idx = 16;
  if (!(-B <= mx)) idx = -1;
  idx_out = idx;

  float co_sel, o_sel;
  pick_bin<false>(f, H1, H2, H3, 0.0f, k, co_sel, o_sel);
  const float ro = __fdividef(twoB, tot2);

  // derivative logits D[idx - 1], D[idx] (raw[32 + j]; raw[47] is the zero pad unless periodic)
  float v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = k.p3 ? raw[43 + i] : (k.p2 ? raw[39 + i] : (k.p1 ? raw[35 + i] : raw[31 + i]));
  float r0 = k.q3 ? v[3] : (k.q2 ? v[2] : (k.q1 ? v[1] : v[0]));
  float r1 = k.q3 ? v[4] : (k.q2 ? v[3] : (k.q1 ? v[2] : v[1]));
  const int ic = min(max(idx, 0), 15);
  r0 = fmaf(r0, sinv, bias[32 + max(ic - 1, 0)]);
  r1 = fmaf(r1, sinv, bias[32 + ic]);

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
    s.dk = tc_softplus(r0);
    s.dkp1 = tc_softplus(r1);
  } else {
    s.dk = (idx == 0) ? 1.0f : tc_softplus(r0);
    s.dkp1 = (idx == 15) ? 1.0f : tc_softplus(r1);
  }
  tc_rqs_eval(s, x, mx, oob, INV, periodic, out, ld);
}

}  // namespace pzb
