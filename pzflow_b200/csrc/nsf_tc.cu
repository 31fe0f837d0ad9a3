// tcgen05 engine: fused chain kernel with the conditioner's two big GEMMs on the 5th-gen
// tensor cores (sm_100a), for the default conditioner family (hidden_layers = 2, hidden_dim = 128,
// K = 16, at most 8 conditioner inputs and 4 transformed columns: BASELINE configs 1-4).
//
// fp32 fidelity: every fp32 operand x is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits), and a GEMM is three kind::f16 MMAs, lo*hi + hi*lo + hi*hi, accumulated
// in fp32 in TMEM.  Weights are pre-scaled by a power of two so their lo parts stay normal.
// Measured against the float64 twin this is indistinguishable from an fp32 GEMM.
//
// Data flow per CTA (1 per SM, persistent, 288 threads):
//   * warps 0-3 / 4-7: two epilogue groups, each owns one 128-row tile at a time (thread = row,
//     TMEM lane = row).  They ping-pong so the tensor pipe works on one tile while CUDA cores
//     run the other tile's epilogue.
//   * warp 8 (one elected lane): TMA bulk loads of the layer's packed weights into shared
//     memory, and every tcgen05.mma / tcgen05.commit.
//   * A operands (the activations) never touch shared memory: epilogue threads write the
//     fp16 hi/lo pairs straight into TMEM (tcgen05.st) and the MMA reads A from TMEM (.ts form);
//     B (weights) comes from shared memory, K-major SWIZZLE_128B.
//   * TMEM (512 columns): tile X: A[0,128) D[128,256); tile Y: A[256,384) D[384,512).
//   * layer-outer order: a CTA takes a chunk of 2048 rows, keeps the layer's weights resident
//     (~166 KB) and walks the chunk's 16 tiles; the per-row state (x[D], log-det stack,
//     conditions) lives in an L2-resident workspace between layers.
//
// Follows the same reference lines as nsf_simt.cu.
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "pzb_common.cuh"

namespace pzb {

constexpr int TC_TILE = 128;
constexpr int TC_CHUNK_TILES = 16;
constexpr int TC_CHUNK = TC_TILE * TC_CHUNK_TILES;
constexpr int TC_THREADS = 288;
constexpr int TC_IN_MAX = 8;
constexpr int TC_L_MAX = 4;
constexpr int TC_D_MAX = 8;
constexpr int TC_PASS_N = 96;          // two spline dims x 48 logit columns per output-layer pass
constexpr int TC_W2_BYTES = 128 * 128 * 2;          // one fp16 [128 x 128] tile
constexpr int TC_W3_BYTES = TC_PASS_N * 128 * 2;    // one fp16 [96 x 128] tile

struct TcLayout {
  int w2hi, w2lo, w3, par, total, n_pass;
  // par (floats): W1[TC_IN_MAX][128], b1[128], b2[128], b3[n_pass*96], scal[4]
  __host__ __device__ static TcLayout make(int L) {
    TcLayout t;
    t.n_pass = (L + 1) / 2;
    t.w2hi = 0;
    t.w2lo = TC_W2_BYTES;
    t.w3 = 2 * TC_W2_BYTES;
    t.par = t.w3 + t.n_pass * 2 * TC_W3_BYTES;
    t.total = t.par + (TC_IN_MAX * 128 + 128 + 128 + t.n_pass * TC_PASS_N + 4) * 4;
    t.total = (t.total + 15) / 16 * 16;
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
  __half* w2hi = reinterpret_cast<__half*>(base + lay.w2hi);
  __half* w2lo = reinterpret_cast<__half*>(base + lay.w2lo);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) split_store(w2hi, w2lo, sw128_index(128, n, k), W2[k * 128 + n] * s2);
  for (int p = 0; p < lay.n_pass; ++p) {
    __half* hi = reinterpret_cast<__half*>(base + lay.w3 + p * 2 * TC_W3_BYTES);
    __half* lo = reinterpret_cast<__half*>(base + lay.w3 + p * 2 * TC_W3_BYTES + TC_W3_BYTES);
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
  float* scal = pb3 + lay.n_pass * TC_PASS_N;
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);  // leave the issue slots to the other tile's warps
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

// softmax knots, bin search, RQS and log-det of one (row, dim) with the 48 logits in registers
// (pzflow/bijectors.py:488-491 + pzflow/utils.py:113-227; K = 16).  The bin search runs on the
// un-normalised cumulative sums (knot_j <= x  <=>  cumsum(e)_j <= (x + B) * sum(e) / 2B), so only
// the selected bin's knots and sizes are ever scaled.
template <bool INV>
__device__ __forceinline__ void spline_regs(const float (&lg)[48], float x, float B, bool periodic, float& out,
                                            float& ld, int& idx_out) {
  constexpr float L2E = 1.4426950408889634f;
  const float twoB = 2.0f * B;
  float mw = lg[0], mh = lg[16];
#pragma unroll
  for (int j = 1; j < 16; ++j) {
    mw = fmaxf(mw, lg[j]);
    mh = fmaxf(mh, lg[16 + j]);
  }
  const float nmw = -(mw * L2E), nmh = -(mh * L2E);
  float W[16], H[16];
  float sw = 0.0f, sh = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    W[j] = ex2_approx(fmaf(lg[j], L2E, nmw));
    H[j] = ex2_approx(fmaf(lg[16 + j], L2E, nmh));
    sw += W[j];
    sh += H[j];
  }
  const float rw = __fdividef(twoB, sw), rh = __fdividef(twoB, sh);  // 2B / sum(e)
  const bool oob = (x <= -B) || (x >= B);
  const float mx = oob ? fabsf(x) - B : x;
  // searched axis S (widths forward, heights inverse), other axis O
  const float rs = INV ? rh : rw, ro = INV ? rw : rh;
  const float target = __fdividef(mx + B, rs);  // (x + B) * sum(e_S) / 2B
  int cnt = (-B <= mx) ? 1 : 0;
  float cs = 0.0f, co = 0.0f;
  float cs_sel = 0.0f, co_sel = 0.0f;           // cumulative sums at the selected bin's left knot
  float s_sel = INV ? H[0] : W[0], o_sel = INV ? W[0] : H[0];
  float r0 = 0.0f, r1 = lg[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    cs += INV ? H[j] : W[j];
    co += INV ? W[j] : H[j];
    const bool p = cs <= target;  // knot j+1 <= x: a prefix of trues (cumsum of positives is monotone)
    cnt += p ? 1 : 0;
    cs_sel = p ? cs : cs_sel;
    co_sel = p ? co : co_sel;
    if (j < 15) {
      s_sel = p ? (INV ? H[j + 1] : W[j + 1]) : s_sel;
      o_sel = p ? (INV ? W[j + 1] : H[j + 1]) : o_sel;
      r1 = p ? lg[32 + j + 1] : r1;  // lg[47] is the zero pad unless periodic
    }
    r0 = p ? lg[32 + j] : r0;
  }
  const int idx = cnt - 1;
  idx_out = idx;
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
    if (idx == 0) r0 = lg[32 + 15];  // dk = pad(D, (1, 0), "wrap")
    s.dk = pz_softplus(r0);
    s.dkp1 = pz_softplus(r1);
  } else {
    s.dk = (idx == 0) ? 1.0f : pz_softplus(r0);
    s.dkp1 = (idx == 15) ? 1.0f : pz_softplus(r1);
  }
  pz_rqs_eval<true>(s, x, mx, oob, INV, periodic, out, ld);
}

struct TcShared {
  uint64_t go[2];    // epilogue group -> MMA thread (128 arrivals): operands ready / accumulator drained
  uint64_t done[2];  // MMA thread -> epilogue group (tcgen05.commit)
  uint64_t wbar;     // weight blob landed (TMA complete_tx)
  uint32_t tmem_base;
  int16_t perm[MAX_D];
};

// One row's light steps between NSC stages, on the workspace row in global/L2 memory.
__device__ __forceinline__ void apply_affine_row(const AffineDev& af, bool inverse, int D, float* row) {
  for (int j = 0; j < D; ++j) {
    const float x = row[j];
    float y;
    if (af.kind == AFF_SHIFT_BOUNDS)
      y = inverse ? (x * af.b[j]) / af.B + af.a[j] : (af.B * (x - af.a[j])) / af.b[j];
    else
      y = inverse ? x * af.b[j] + af.a[j] : (x - af.a[j]) / af.b[j];
    row[j] = y;
  }
}

// Applies the non-NSC steps in [s0, s1) (already mapped to walk order) to one workspace row.
__device__ __forceinline__ void light_steps_row(const PlanDev* plan, bool inverse, int s0, int s1, int& level, float* row) {
  const int D = plan->D, ns = plan->n_steps;
  float* ldl = row + D;
  for (int si = s0; si < s1; ++si) {
    const StepDev step = plan->steps[inverse ? ns - 1 - si : si];
    int kind = step.kind;
    if (inverse && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
    else if (inverse && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
    if (kind == STEP_CHAIN_BEGIN) {
      ++level;
      ldl[level] = 0.0f;
    } else if (kind == STEP_CHAIN_END) {
      ldl[level - 1] += ldl[level];
      --level;
    } else if (kind == STEP_AFFINE) {
      const AffineDev& af = plan->affine[step.idx];
      apply_affine_row(af, inverse, D, row);
      ldl[level] += inverse ? af.logdet_inv : af.logdet_fwd;
    }
  }
}

template <bool INV>
__global__ void __launch_bounds__(TC_THREADS, 1)
nsf_chain_tc_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, float* __restrict__ wsp, int smem_blob_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, derived by pointer arithmetic so loads stay LDS (not generic LD)
  uint8_t* blob = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  TcShared* sh = reinterpret_cast<TcShared*>(blob + smem_blob_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = plan->D, C = plan->C, ns = plan->n_steps;
  const int RS = D + MAX_CHAIN_DEPTH + C;  // workspace row: x[D], ld[4], cond[C]
  float* ws = wsp + (size_t)blockIdx.x * TC_CHUNK * RS;
  const bool is_epi = warp < 8;
  const int grp = warp >> 2;                       // epilogue group (tile X / Y)
  const int lrow = ((warp & 3) << 5) + lane;       // row within the tile == TMEM lane

  if (tid == 0) {
    mbar_init(smem_u32(&sh->go[0]), 128);
    mbar_init(smem_u32(&sh->go[1]), 128);
    mbar_init(smem_u32(&sh->done[0]), 1);
    mbar_init(smem_u32(&sh->done[1]), 1);
    mbar_init(smem_u32(&sh->wbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sh->tmem_base;
  const uint32_t t_lane = tbase + ((uint32_t)((warp & 3) << 5) << 16);
  const uint32_t a_col = (uint32_t)grp * 256u, d_col = (uint32_t)grp * 256u + 128u;
  uint32_t ph_go[2] = {0u, 0u}, ph_done = 0u, ph_w = 0u;

  const long long n_chunks = (args.N + TC_CHUNK - 1) / TC_CHUNK;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const long long row0 = chunk * TC_CHUNK;
    const int nrows = (int)min((long long)TC_CHUNK, args.N - row0);
    const int npairs = ((nrows + TC_TILE - 1) / TC_TILE + 1) / 2;

    // first / last NSC step in walk order
    int first_nsc = ns, last_nsc = -1;
    for (int si = 0; si < ns; ++si)
      if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
        if (first_nsc == ns) first_nsc = si;
        last_nsc = si;
      }

    // ---- pass A: load rows into the workspace, run the steps in front of the first NSC ----
    if (is_epi) {
      for (int r = tid; r < nrows; r += 256) {
        const long long row = row0 + r;
        float* wr = ws + (size_t)r * RS;
        for (int j = 0; j < D; ++j) {
          float v;
          if (args.mode == MODE_POSTERIOR) {
            const long long g = row / args.G;
            const int gi = (int)(row - g * args.G);
            v = (j == args.col_idx) ? args.grid[gi] : args.X[g * (D - 1) + (j < args.col_idx ? j : j - 1)];
          } else {
            v = args.X[row * D + j];
          }
          wr[INV ? plan->perm_final[j] : j] = v;
        }
        for (int l = 0; l < MAX_CHAIN_DEPTH; ++l) wr[D + l] = 0.0f;
        const long long crow = (args.mode == MODE_POSTERIOR) ? row / args.G : row;
        for (int j = 0; j < C; ++j) wr[D + MAX_CHAIN_DEPTH + j] = args.cond ? args.cond[crow * C + j] : 0.0f;
        int level = 0;
        light_steps_row(plan, INV, 0, first_nsc, level, wr);
      }
    }
    int level = 0;  // chain depth after the steps walked so far (row independent)
    for (int si = 0; si < first_nsc; ++si) {
      int kind = plan->steps[INV ? ns - 1 - si : si].kind;
      if (INV && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
      else if (INV && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
      level += (kind == STEP_CHAIN_BEGIN) - (kind == STEP_CHAIN_END);
    }
    __syncthreads();

    // ---- the coupling stages ----
    int prev_end = first_nsc;
    for (int si = first_nsc; si <= last_nsc; ++si) {
      const StepDev step = plan->steps[INV ? ns - 1 - si : si];
      if (step.kind != STEP_NSC) continue;
      if (si > prev_end) {  // light steps between two coupling stages
        if (is_epi)
          for (int r = tid; r < nrows; r += 256) {
            int lv = level;
            light_steps_row(plan, INV, prev_end, si, lv, ws + (size_t)r * RS);
          }
        for (int sj = prev_end; sj < si; ++sj) {
          int kind = plan->steps[INV ? ns - 1 - sj : sj].kind;
          if (INV && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
          else if (INV && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
          level += (kind == STEP_CHAIN_BEGIN) - (kind == STEP_CHAIN_END);
        }
        __syncthreads();
      }
      prev_end = si + 1;
      const NscDev& st = plan->nsc[step.idx];
      const TcLayout lay = TcLayout::make(st.L);
      const int U = st.U, L = st.L, in_dim = st.U + st.C;

      if (tid < D) sh->perm[tid] = st.perm[tid];
      if (warp == 8 && lane == 0) {
        // weights of this stage: global (L2) -> shared memory, one TMA bulk copy per 16 KB
        const uint32_t bar = smem_u32(&sh->wbar);
        mbar_expect_tx(bar, (uint32_t)lay.total);
        const uint8_t* src = static_cast<const uint8_t*>(st.tc.blob);
        for (int off = 0; off < lay.total; off += 16384) {
          const int nb = min(16384, lay.total - off);
          tma_bulk_g2s(smem_u32(blob + off), src + off, (uint32_t)nb, bar);
        }
      }
      __syncthreads();
      mbar_wait(smem_u32(&sh->wbar), ph_w);
      ph_w ^= 1u;

      const float* par = reinterpret_cast<const float*>(blob + lay.par);
      const float* sW1 = par;
      const float* sb1 = par + TC_IN_MAX * 128;
      const float* sb2 = sb1 + 128;
      const float* sb3 = sb2 + 128;
      const float s2inv = sb3[lay.n_pass * TC_PASS_N], s3inv = sb3[lay.n_pass * TC_PASS_N + 1];

      if (is_epi) {
        // =============================== epilogue groups ===============================
        for (int pair = 0; pair < npairs; ++pair) {
          const int r = (2 * pair + grp) * TC_TILE + lrow;
          const bool valid = r < nrows;
          float* wr = ws + (size_t)(valid ? r : 0) * RS;
          float key[TC_IN_MAX], low[TC_L_MAX];
#pragma unroll
          for (int k = 0; k < TC_IN_MAX; ++k)
            key[k] = (valid && k < in_dim) ? (k < U ? wr[sh->perm[k]] : wr[D + MAX_CHAIN_DEPTH + (k - U)]) : 0.0f;
#pragma unroll
          for (int d = 0; d < TC_L_MAX; ++d) low[d] = (valid && d < L) ? wr[sh->perm[U + d]] : 0.0f;

          // ---- E1: h1 = LeakyReLU(key @ W1 + b1) -> A (fp16 hi/lo) in TMEM ----
#pragma unroll
          for (int bt = 0; bt < 4; ++bt) {
            uint32_t rh[16], rl[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int k0 = bt * 32 + q * 4;
              float4 acc = *reinterpret_cast<const float4*>(sb1 + k0);
#pragma unroll
              for (int i = 0; i < TC_IN_MAX; ++i) {
                if (i < in_dim) {
                  const float4 w = *reinterpret_cast<const float4*>(sW1 + i * 128 + k0);
                  acc.x = fmaf(key[i], w.x, acc.x);
                  acc.y = fmaf(key[i], w.y, acc.y);
                  acc.z = fmaf(key[i], w.z, acc.z);
                  acc.w = fmaf(key[i], w.w, acc.w);
                }
              }
              split2(pz_leaky(acc.x), pz_leaky(acc.y), rh[2 * q], rl[2 * q]);
              split2(pz_leaky(acc.z), pz_leaky(acc.w), rh[2 * q + 1], rl[2 * q + 1]);
            }
            tmem_st16(t_lane + a_col + bt * 16, rh);
            tmem_st16(t_lane + a_col + 64 + bt * 16, rl);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(smem_u32(&sh->go[grp]));

          // ---- E2: h2 = LeakyReLU(D * 2^-e2 + b2) -> A ----
          mbar_wait(smem_u32(&sh->done[grp]), ph_done);
          ph_done ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int bt = 0; bt < 4; ++bt) {
            float dv[32];
            tmem_ld16(t_lane + d_col + bt * 32, dv);
            tmem_ld16(t_lane + d_col + bt * 32 + 16, dv + 16);
            tc_wait_ld();
            uint32_t rh[16], rl[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = *reinterpret_cast<const float4*>(sb2 + bt * 32 + q * 4);
              const float h0 = pz_leaky(fmaf(dv[4 * q + 0], s2inv, bb.x));
              const float h1 = pz_leaky(fmaf(dv[4 * q + 1], s2inv, bb.y));
              const float h2 = pz_leaky(fmaf(dv[4 * q + 2], s2inv, bb.z));
              const float h3 = pz_leaky(fmaf(dv[4 * q + 3], s2inv, bb.w));
              split2(h0, h1, rh[2 * q], rl[2 * q]);
              split2(h2, h3, rh[2 * q + 1], rl[2 * q + 1]);
            }
            tmem_st16(t_lane + a_col + bt * 16, rh);
            tmem_st16(t_lane + a_col + 64 + bt * 16, rl);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(smem_u32(&sh->go[grp]));

          // ---- E3: logits -> knots -> spline, two dims per output-layer pass ----
          float ldsum = 0.0f;
          for (int p = 0; p < lay.n_pass; ++p) {
            mbar_wait(smem_u32(&sh->done[grp]), ph_done);
            ph_done ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int d = 2 * p + j;
              if (d < L) {  // warp-uniform
                float lg[48];
                tmem_ld16(t_lane + d_col + 48 * j, lg);
                tmem_ld16(t_lane + d_col + 48 * j + 16, lg + 16);
                tmem_ld16(t_lane + d_col + 48 * j + 32, lg + 32);
                tc_wait_ld();
                const float* bb = sb3 + p * TC_PASS_N + 48 * j;
#pragma unroll
                for (int q = 0; q < 48; q += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bb + q);
                  lg[q + 0] = fmaf(lg[q + 0], s3inv, b4.x);
                  lg[q + 1] = fmaf(lg[q + 1], s3inv, b4.y);
                  lg[q + 2] = fmaf(lg[q + 2], s3inv, b4.z);
                  lg[q + 3] = fmaf(lg[q + 3], s3inv, b4.w);
                }
                float xin = 0.0f;
#pragma unroll
                for (int dd = 0; dd < TC_L_MAX; ++dd) xin = (dd == d) ? low[dd] : xin;
                float y, ldd;
                int idx;
                spline_regs<INV>(lg, xin, st.B, st.periodic != 0, y, ldd, idx);
                ldsum += ldd;  // log_det.sum(axis=1), dims in order
                if (valid) {
                  wr[sh->perm[U + d]] = y;
                  if (args.bins != nullptr) args.bins[(row0 + r) * plan->n_spline_dims + st.bins_off + d] = idx;
                }
              }
            }
            if (p + 1 < lay.n_pass) {  // accumulator drained: the next pass may overwrite it
              tc_fence_before();
              mbar_arrive(smem_u32(&sh->go[grp]));
            }
          }
          if (valid) wr[D + level] += INV ? -ldsum : ldsum;
        }
      } else if (lane == 0) {
        // =============================== MMA issuer ===============================
        const uint32_t sb = smem_u32(blob);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc3 = (1u << 4) | ((uint32_t)(TC_PASS_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int pair = 0; pair < npairs; ++pair) {
          for (int g = 0; g < 2; ++g) {
            mbar_wait(smem_u32(&sh->go[g]), ph_go[g]);
            ph_go[g] ^= 1u;
            tc_fence_after();
            issue_gemm(tbase + g * 256 + 128, tbase + g * 256, sb + lay.w2hi, sb + lay.w2lo, 128, idesc2);
            umma_commit(smem_u32(&sh->done[g]));
          }
          for (int p = 0; p < lay.n_pass; ++p) {
            for (int g = 0; g < 2; ++g) {
              mbar_wait(smem_u32(&sh->go[g]), ph_go[g]);
              ph_go[g] ^= 1u;
              tc_fence_after();
              const uint32_t w3 = sb + lay.w3 + p * 2 * TC_W3_BYTES;
              issue_gemm(tbase + g * 256 + 128, tbase + g * 256, w3, w3 + TC_W3_BYTES, TC_PASS_N, idesc3);
              umma_commit(smem_u32(&sh->done[g]));
            }
          }
        }
      }
      tc_fence_before();
      __syncthreads();  // weights, TMEM and the workspace rows are settled before the next stage
      tc_fence_after();
    }

    // ---- pass Z: trailing steps, latent log-density, outputs ----
    if (is_epi) {
      for (int r = tid; r < nrows; r += 256) {
        const long long row = row0 + r;
        float* wr = ws + (size_t)r * RS;
        int lv = level;
        light_steps_row(plan, INV, prev_end, ns, lv, wr);
        if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
          const float lat = pz_latent_logprob(plan->latent_kind, D, plan->latent_B, plan->latent_const,
                                              [&](int j) { return wr[plan->perm_final[j]]; });
          const float lp = pz_nan_to_num_lp(lat + wr[D]);
          args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lp) : lp;
        } else {
          for (int j = 0; j < D; ++j) args.out0[row * D + j] = wr[INV ? j : plan->perm_final[j]];
          if (args.out1 != nullptr) args.out1[row] = wr[D];
        }
      }
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

cudaError_t launch_tc(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms, void* workspace,
                      cudaStream_t stream) {
  (void)workspace;
  int blob = 0;
  for (int s = 0; s < hp.n_nsc; ++s) blob = max(blob, TcLayout::make(hp.nsc[s].L).total);
  const size_t smem = (size_t)blob + sizeof(TcShared) + 1024 + 64;
  const bool inv = (args.mode == MODE_INVERSE);
  cudaError_t e = inv ? cudaFuncSetAttribute(nsf_chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(nsf_chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long n_chunks = (args.N + TC_CHUNK - 1) / TC_CHUNK;
  long long grid = num_sms;
  if (grid > n_chunks) grid = n_chunks;
  if (grid < 1) grid = 1;
  const int RS = hp.D + MAX_CHAIN_DEPTH + hp.C;
  float* ws = nullptr;
  e = cudaMallocAsync(&ws, (size_t)grid * TC_CHUNK * RS * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  if (inv)
    nsf_chain_tc_kernel<true><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, ws, blob);
  else
    nsf_chain_tc_kernel<false><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, ws, blob);
  e = cudaGetLastError();
  cudaError_t e2 = cudaFreeAsync(ws, stream);
  return e != cudaSuccess ? e : e2;
}

}  // namespace pzb
