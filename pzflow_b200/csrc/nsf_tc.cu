// tcgen05 engine: fused chain kernel with ALL THREE conditioner GEMMs on the 5th-gen tensor cores
// (sm_100a), for the default conditioner family (hidden_layers = 2, hidden_dim <= 128, K <= 16 -- narrower
// shapes are padded into 128 / 16 by the packer, exactly -- at most
// 15 conditioner inputs and 31 data columns: BASELINE configs 1-4 and every default flow up to 30-D).
// A coupling stage with more than 4 transformed columns runs as sub-stages of 4 (tc_sub_stages below).
//
// fp32 fidelity: every fp32 operand x is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits), and a GEMM is three kind::f16 MMAs, lo*hi + hi*lo + hi*hi, accumulated
// in fp32 in TMEM.  W2 / W3 are pre-scaled by a power of two so their lo parts stay normal.
// Measured against the float64 twin this is indistinguishable from an fp32 GEMM.
//
// Structure (one persistent CTA per SM, 608 threads):
//   * 16 epilogue warps in two groups of 8.  A group owns one 128-row tile at a time; TMEM lane =
//     row, and TWO threads share a row (warp w and w+4 of the group address the same TMEM lane
//     quarter): "half" h does hidden columns [64h, 64h+64) in E1/E2 and the spline dimensions
//     h, h+2 in E3.  Per tile and coupling stage:
//       E0  the <= 15 conditioner inputs + a constant 1 -> K = 16 A operand in Y[0,24)  (1 thread/row)
//       G1  keys x W1 tile (bias row included) -> X                                      3 MMAs
//       E1  X: LeakyReLU, fp16 hi/lo split, rewritten in place
//       G2  X x W2 -> Y in four N = 32 quarters, one commit each, alternating between the halves,
//           so E2 of a quarter runs while the tensor pipe computes the next ones
//       E2  Y: * 2^-e2 + b2, LeakyReLU, split, in place
//       G3  Y x W3 -> X[0,96): one N = 96 pass per PAIR of spline dimensions (dims 2p, 2p + 1 land in X[0,48) /
//           X[48,96); a tcgen05.mma costs the issuing warp ~5 clk on top of its pipe time, so fewer, wider
//           instructions).  The first pass starts on half 0's 64 K-elements and takes half 1's when it arrives,
//           the second goes out once both halves have the first in registers
//       E3  48 logits of a (row, dim) in registers -> knots, bin search, RQS, log-det
//     The half that finishes E3 first writes the NEXT tile's E0, so G1 runs under the last spline.
//   * warps 16/17: MMA issuer of group 0/1.  The whole warp waits on the mbarriers, one elect.sync
//     lane issues, which keeps the tcgen05.mma operands on the uniform datapath.
//   * warp 18: streams the weights of the NEXT coupling stage with TMA bulk copies as soon as
//     tcgen05.commit has released a buffer (W2 64 KB, W3 48/96 KB, W1 tile 8 KB; the fp32 parameter
//     block with b2, b3, scales and the stage's column map is double buffered).
//   * A operands (activations) never touch shared memory: epilogue threads write fp16 hi/lo pairs
//     straight into TMEM (tcgen05.st) and the MMA reads A from TMEM (.ts form); B (weights) sits in
//     shared memory, K-major SWIZZLE_128B (W1: SWIZZLE_NONE).
//   * TMEM (512 columns), per group g two 128-column regions X = 256g, Y = 256g + 128.  32 fp32
//     accumulator columns become 16 hi + 16 lo columns IN PLACE, so a tile needs exactly 256 columns
//     and two tiles are in flight.  A thread only ever overwrites TMEM columns it alone has read;
//     every cross-thread hand-over is an mbarrier with one arrival per warp.
//   * stage-outer order: a CTA takes a chunk of up to 1024 rows (an even number of tiles, so both
//     groups get the same count) whose state (x[D], log-det levels, conditions) lives in shared
//     memory as [column][row] for the whole chain, and walks the chunk's tiles once per coupling
//     stage.  HBM sees the input rows and the result only.
//   * -DPZB_TC_TRACE builds SM-clock phase stamps of CTA 0 into the kernel (pzb_debug_tc_trace).
//
// Follows the same reference lines as nsf_simt.cu.
#include "pzb_tc_common.cuh"

#ifndef PZB_TC_PIPE3_DEFAULT
#define PZB_TC_PIPE3_DEFAULT 0   // which kernel runs when the environment does not say (PZB_TC_PIPE3)
#endif
#ifndef PZB_TC_PAIR_PASS
#define PZB_TC_PAIR_PASS 1   // output layer in N = 96 passes of two spline dimensions (0: one N = 48 pass per dimension, the form up to round 2; +2 %, same bits: DESIGN.md section 4.5b)
#endif
#ifndef PZB_TC_G2_HALVES
#define PZB_TC_G2_HALVES 0   // experiment: hidden GEMM in two N = 64 halves instead of four N = 32 quarters
#endif
#ifndef PZB_TC_STUB
#define PZB_TC_STUB 0   // bit 0 / bit 1: measurement builds without the spline / the activation arithmetic
#endif

namespace pzb {

constexpr int TC_TILE = 128;
constexpr int TC_GROUP_THREADS = 256;
constexpr int TC_EPI_THREADS = 512;
constexpr int TC_THREADS = 608;                   // 16 epilogue warps + 2 MMA issuers + 1 loader
constexpr int TC_IN_MAX = 15;                     // conditioner inputs: K = 16 minus the bias row
constexpr int TC_L_MAX = 4;
constexpr int TC_D_MAX = 31;
constexpr int TC_SEQ_MAX = 255;                   // sub-stages of one chain walk (uint8 tables in shared memory)
constexpr int TC_MAX_CHUNK_TILES = 16;
constexpr int TC_DIM_N = 48;                      // logit columns of one spline dimension (47 or 48 used)
constexpr int TC_PASS_N = 96;                     // rows of one packed W3 tile: two spline dimensions
constexpr int TC_W2_BYTES = 2 * 128 * 128 * 2;    // fp16 hi + lo [128 x 128] tiles
constexpr int TC_W3_PASS_BYTES = 2 * TC_PASS_N * 128 * 2;  // fp16 hi + lo [96 x 128] tiles
constexpr int TC_W1_BYTES = 2 * 128 * 16 * 2;     // fp16 hi + lo [128 x 16] tiles: K = key (<= 8), 1 (bias row), zero pad
// fp32 parameter block of a stage: b2[128], b3[2 * 96], scal[4], then 24 ints
constexpr int TC_PF_B2 = 0;
constexpr int TC_PF_B3 = TC_PF_B2 + 128;
constexpr int TC_PF_SCAL = TC_PF_B3 + 2 * TC_PASS_N;  // s2inv, s3inv, B, 1/(2B)
constexpr int TC_PI = TC_PF_SCAL + 4;                 // ints: key_col[16], low_col[4], U, L, in_dim, periodic, bins_off
constexpr int TC_PI_KEY = 0, TC_PI_LOW = 16, TC_PI_U = 20, TC_PI_L = 21, TC_PI_IN = 22, TC_PI_PERIODIC = 23,
              TC_PI_BINS = 24, TC_PI_K = 25;
constexpr int TC_PAR_FLOATS = TC_PI + 32;
constexpr int TC_PAR_BYTES = TC_PAR_FLOATS * 4;   // 1424, a multiple of 16
constexpr int TC_SMEM_MAX = 232448;               // 227 KB opt-in limit per CTA
static_assert(TC_PAR_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

// global blob of one coupling stage: [W2 hi | W2 lo | pass0 hi | pass0 lo | pass1 hi | pass1 lo | W1 hi | W1 lo | par]
struct TcLayout {
  int w3, w1, par, total, n_pass;
  __host__ __device__ static TcLayout make(int L) {
    TcLayout t;
    t.n_pass = (L + 1) / 2;
    t.w3 = TC_W2_BYTES;
    t.w1 = t.w3 + t.n_pass * TC_W3_PASS_BYTES;
    t.par = t.w1 + TC_W1_BYTES;
    t.total = t.par + TC_PAR_BYTES;
    return t;
  }
};


// ------------------------------------------------------------------------------------------
// host side: shape check + packing
// ------------------------------------------------------------------------------------------
// A coupling stage with more than four spline dimensions runs as ceil(L / 4) SUB-STAGES: each recomputes the
// conditioner's two hidden layers from the same (untouched) conditioner inputs and evaluates its own four columns of
// the output layer, so to the kernel a sub-stage is an ordinary stage with its own weight blob.
__host__ __device__ inline int tc_sub_stages(int L) { return (L + TC_L_MAX - 1) / TC_L_MAX; }
__host__ __device__ inline int tc_sub_dims(int L, int sub) { return min(TC_L_MAX, L - TC_L_MAX * sub); }

bool tc_plan_supported(const PlanDev& hp) {
  if (hp.n_nsc < 1 || hp.D > TC_D_MAX) return false;
  int total = 0;
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    if (st.n_dense != 3 || st.H < 1 || st.H > 128 || st.K < 1 || st.K > 16) return false;
    if (st.K < 16 && st.periodic) return false;  // padding bins would sit inside the wrap-around
    if (st.U + st.C > TC_IN_MAX || st.L > TC_L_MAX * TC_SUB_MAX || st.L < 1) return false;
    total += tc_sub_stages(st.L);
  }
  return total <= TC_SEQ_MAX;
}

int tc_stage_subs(const NscDev& st) { return tc_sub_stages(st.L); }
size_t tc_pack_bytes(const NscDev& st, int sub) { return (size_t)TcLayout::make(tc_sub_dims(st.L, sub)).total; }


// state columns of a chunk row: x[0, D), log-det levels [D, D + n_lev), pending log-det halves (2),
// conditions (C)
// Where logit q of the kernel's K = 16 layout ([16 widths | 16 heights | 15 or 16 derivatives]) comes from in the
// stage's own [K widths | K heights | K - 1 (periodic: K) derivatives] layout, or which padding it is.
constexpr int TC_PAD_BIN = -1, TC_PAD_UNIT = -2, TC_PAD_NONE = -3;
static int tc_logit_source(const NscDev& st, int q) {
  const int K = st.K;
  if (q < 16) return q < K ? q : TC_PAD_BIN;
  if (q < 32) return q - 16 < K ? K + (q - 16) : TC_PAD_BIN;
  const int j = q - 32, n_der = st.cpd - 2 * K;   // derivative j belongs to knot j + 1
  if (j < n_der) return 2 * K + j;
  return (j == n_der && K < 16) ? TC_PAD_UNIT : TC_PAD_NONE;
}

void tc_pack_stage(const PlanDev& hp, const NscDev& st, int sub, const float* const* weights, void* dst) {
  const int d0 = TC_L_MAX * sub, Ls = tc_sub_dims(st.L, sub);  // this sub-stage's spline dimensions [d0, d0 + Ls)
  const TcLayout lay = TcLayout::make(Ls);
  unsigned char* base = static_cast<unsigned char*>(dst);
  memset(base, 0, lay.total);
  const float *W1 = weights[0], *b1 = weights[1], *W2 = weights[2], *b2 = weights[3], *W3 = weights[4],
              *b3 = weights[5];
  const int in_dim = st.U + st.C, out_dim = st.cpd * st.L, H = st.H;
  // The kernel is written for hidden_dim 128 and K = 16.  Narrower conditioners are PADDED into that shape, exactly:
  // hidden units H..127 get zero weights and biases (LeakyReLU(0) = 0 contributes nothing downstream), and a spline
  // with K < 16 bins gets 16 - K trailing bins of zero width and height (bin logits with zero weights and a bias of
  // -1e30: their softmax terms are exactly 0, and a bin whose left knot is B is never selected), with the knot that
  // closes the last real bin given derivative softplus(log(e - 1)) = 1, the boundary value (pzflow/utils.py:130-135).
  const int e2 = pow2_scale_exp(W2, (size_t)H * H), e3 = pow2_scale_exp(W3, (size_t)H * out_dim);
  const float s2 = ldexpf(1.0f, e2), s3 = ldexpf(1.0f, e3);
  __half* w2hi = reinterpret_cast<__half*>(base);
  __half* w2lo = reinterpret_cast<__half*>(base + TC_W2_BYTES / 2);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k)
      split_store(w2hi, w2lo, sw128_index(128, n, k), (n < H && k < H) ? W2[k * H + n] * s2 : 0.0f);
  for (int p = 0; p < lay.n_pass; ++p) {
    __half* hi = reinterpret_cast<__half*>(base + lay.w3 + p * TC_W3_PASS_BYTES);
    __half* lo = reinterpret_cast<__half*>(base + lay.w3 + p * TC_W3_PASS_BYTES + TC_W3_PASS_BYTES / 2);
    for (int j = 0; j < 2; ++j) {
      const int d = 2 * p + j;
      if (d >= Ls) continue;
      for (int q = 0; q < TC_DIM_N; ++q) {
        const int src = tc_logit_source(st, q);
        if (src < 0) continue;  // padding: zero weights
        const int col = (d0 + d) * st.cpd + src, n = TC_DIM_N * j + q;
        for (int k = 0; k < H; ++k) split_store(hi, lo, sw128_index(TC_PASS_N, n, k), W3[(size_t)k * out_dim + col] * s3);
      }
    }
  }
  {  // first layer as a [128 x 16] B operand, K-major without swizzle: K = (key_0 .. key_7, 1 (bias row, k = 8), key_8 .. key_14)
     // so the bias rides in the GEMM.  Unscaled: |W1| and b1 are O(1), and the lo parts of small weights are
     // fp16 subnormals, which the tensor core takes at full precision of their 2^-24 grid.
    __half* hi = reinterpret_cast<__half*>(base + lay.w1);
    __half* lo = reinterpret_cast<__half*>(base + lay.w1 + TC_W1_BYTES / 2);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < 16; ++k) {  // k = 8 is the bias row; inputs 8.. sit one row further up
        const int i = k < 8 ? k : k - 1;
        const float v = n >= H ? 0.0f : (k == 8 ? b1[n] : (i < in_dim ? W1[i * H + n] : 0.0f));
        split_store(hi, lo, w1_index(n, k), v);
      }
  }
  float* par = reinterpret_cast<float*>(base + lay.par);
  for (int k = 0; k < 128; ++k) par[TC_PF_B2 + k] = k < H ? b2[k] : 0.0f;
  for (int d = 0; d < Ls; ++d)
    for (int q = 0; q < TC_DIM_N; ++q) {
      const int src = tc_logit_source(st, q);
      par[TC_PF_B3 + d * TC_DIM_N + q] = src >= 0 ? b3[(d0 + d) * st.cpd + src]
                                         : (src == TC_PAD_BIN ? -1e30f : (src == TC_PAD_UNIT ? 0.5413248546129181f : 0.0f));
    }
  par[TC_PF_SCAL + 0] = ldexpf(1.0f, -e2);
  par[TC_PF_SCAL + 1] = ldexpf(1.0f, -e3);
  par[TC_PF_SCAL + 2] = st.B;
  par[TC_PF_SCAL + 3] = 1.0f / (2.0f * st.B);
  int32_t* pi = reinterpret_cast<int32_t*>(par + TC_PI);
  const int cond_col0 = hp.D + tc_log_det_levels(hp) + 2;
  for (int i = 0; i < TC_IN_MAX; ++i)
    pi[TC_PI_KEY + i] = i < st.U ? (int)st.perm[i] : (i < in_dim ? cond_col0 + (i - st.U) : 0);
  for (int d = 0; d < TC_L_MAX; ++d) pi[TC_PI_LOW + d] = d < Ls ? (int)st.perm[st.U + d0 + d] : 0;
  pi[TC_PI_U] = st.U;
  pi[TC_PI_L] = Ls;
  pi[TC_PI_IN] = in_dim;
  pi[TC_PI_PERIODIC] = st.periodic;
  pi[TC_PI_BINS] = st.bins_off + d0;
  pi[TC_PI_K] = st.K;
}

// ------------------------------------------------------------------------------------------
// device side: PTX wrappers
// ------------------------------------------------------------------------------------------

// barriers of nsf_chain_tc3_kernel (nsf_tc_pipe3.cuh)
struct TcPipe3Bars {
  uint64_t d1_full[3];      // [region]: first-layer GEMM landed in X[region] (commit)
  uint64_t e1_ready[3];     // [region] (16 warp arrivals): h1 is written
  uint64_t d3_full[3][2];   // [region][pass]: that output pass (two spline dimensions, N = 96) landed (commit)
  uint64_t d3_free[3][2];   // [region][pass] (8 warp arrivals): the pass's two quarters have their 48 logits in registers
  uint64_t d2_full[2];      // [half]: that N = 64 half of the hidden GEMM landed in Y (commit)
  uint64_t e2_ready[2];     // [half] (8 warp arrivals): that half of h2 (two quarters) is written
  uint64_t key_ready;       // (4 warp arrivals): the key operand of the next tile is in shared memory
  uint64_t key_free;        // the first-layer GEMM that read it has completed (commit)
  uint16_t lt0[TC_SEQ_MAX + 1], lt1[TC_SEQ_MAX + 1];   // per (sub-)stage: the steps between the previous coupling stage and it
  int8_t lvl_in[TC_SEQ_MAX + 1], lvl[TC_SEQ_MAX + 1];  // chain depth in front of those steps / at the stage
};
struct TcShared {
  uint64_t e_ready[2][2];  // [group][half] -> the group's MMA issuer (4 warp arrivals): that half's 64 K-elements of the
                           // A operand are written (after E1 and after E2)
  uint64_t d2_full[2][4];  // [group][2 * half + batch]: that N = 32 quarter of the hidden GEMM has completed (commit)
  uint64_t d3_full[2][2];  // [group][buffer]: an output pass has landed in that accumulator buffer
  uint64_t d3_free[2][2];  // [group][buffer]: the half's 4 warps have drained the buffer
  uint64_t k_ready[2];     // [group] (8 warp arrivals): the next tile's key operand is written and X is drained
  uint64_t d1_full[2];     // [group]: the first-layer GEMM has landed in X
  uint64_t g3_done[2];     // [group]: every output pass of the tile has completed (Y is free again)
  uint64_t w1_full, w2_full, w3_full, par_full[2];  // TMA complete_tx
  uint64_t w1_free, w2_free, w3_free;      // both issuers' tcgen05.commit after the stage's last GEMM on that buffer
  uint64_t par_free[2];                    // 16 warp arrivals: every epilogue warp is done with that parameter buffer
  uint32_t tmem_base;
  uint32_t ovf[TC_MAX_CHUNK_TILES * TC_TILE / 32];  // one bit per chunk row: FINITE conditioner inputs gave non-finite logits in some stage
  uint32_t kok[TC_MAX_CHUNK_TILES * TC_TILE / 32];  // one bit per chunk row: the current stage's conditioner inputs are finite
  int n_seq;                    // sub-stages of one chain walk
  uint8_t order[TC_SEQ_MAX + 1];  // sub-stages in walk order: coupling stage (index into PlanDev::nsc) ...
  uint8_t sub[TC_SEQ_MAX + 1];    // ... and which four of its spline dimensions
  TcPipe3Bars p3;
};

struct TcSmem {  // byte offsets from the 1024-aligned base
  int w2, w3, w1, par, keys, state, acc, shared, total;
};

// key_bytes: the three-tile kernel keeps the first layer's key operand in shared memory (TC_W1_BYTES), the other one in TMEM (0)
__host__ __device__ inline TcSmem tc_smem_layout(int max_pass, int chunk_tiles, int state_cols, int acc_floats = 0,
                                                 int key_bytes = 0) {
  TcSmem s;
  s.w2 = 0;
  s.w3 = TC_W2_BYTES;
  s.w1 = s.w3 + max_pass * TC_W3_PASS_BYTES;
  s.par = s.w1 + TC_W1_BYTES;
  s.keys = s.par + 2 * TC_PAR_BYTES;
  s.state = s.keys + key_bytes;
  s.acc = s.state + chunk_tiles * TC_TILE * state_cols * 4;
  s.shared = s.acc + acc_floats * 4;
  s.shared = (s.shared + 15) / 16 * 16;
  s.total = s.shared + (int)sizeof(TcShared);
  return s;
}

struct TcLaunch {
  int chunk_tiles, max_pass, state_cols, n_lev;
  int key_bytes;  // TC_W1_BYTES: nsf_chain_tc3_kernel (three tiles in flight), 0: nsf_chain_tc_kernel
  // fused posterior reduction (LaunchArgs::fuse): a unit = one galaxy = unit_rows = max(V,1) * max(S,1) * G consecutive rows.
  //   upc > 0: a chunk holds upc whole units (unit_rows <= chunk rows);
  //   cpu > 0: a unit spans cpu chunks of spc samples (of G rows) each, all walked by the same CTA, sums kept in shared memory.
  int unit_rows, upc, cpu, spc;
  long long n_units;
};

// rows [row0, row0 + nrows) of the ci-th chunk of this CTA; unit0 = its first unit, sub = which part of a long unit
__device__ __forceinline__ void tc_chunk_span(const TcLaunch& lp, const LaunchArgs& a, int ci, long long& row0, int& nrows,
                                              long long& unit0, int& sub) {
  const long long CR = (long long)lp.chunk_tiles * TC_TILE;
  sub = 0;
  if (lp.unit_rows == 0) {
    row0 = ((long long)blockIdx.x + (long long)ci * gridDim.x) * CR;
    nrows = (int)min(CR, a.N - row0);
    unit0 = 0;
  } else if (lp.upc > 0) {
    unit0 = ((long long)blockIdx.x + (long long)ci * gridDim.x) * lp.upc;
    row0 = unit0 * lp.unit_rows;
    nrows = (int)(min((long long)lp.upc, lp.n_units - unit0) * lp.unit_rows);
  } else {
    unit0 = (long long)blockIdx.x + (long long)(ci / lp.cpu) * gridDim.x;
    sub = ci % lp.cpu;
    const int vs = lp.unit_rows / a.G;
    row0 = unit0 * lp.unit_rows + (long long)sub * lp.spc * a.G;
    nrows = min(lp.spc, vs - sub * lp.spc) * a.G;
  }
}
__device__ __forceinline__ int tc_my_chunks(const TcLaunch& lp, const LaunchArgs& a) {
  const long long CR = (long long)lp.chunk_tiles * TC_TILE;
  long long n;  // chunks (or units) handed out round robin over the CTAs
  if (lp.unit_rows == 0) n = (a.N + CR - 1) / CR;
  else if (lp.upc > 0) n = (lp.n_units + lp.upc - 1) / lp.upc;
  else n = lp.n_units;
  const int mine = n > blockIdx.x ? (int)((n - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  return (lp.unit_rows != 0 && lp.upc == 0) ? mine * lp.cpu : mine;
}

__device__ __forceinline__ float tc_nan_to_num0(float v) {
  if (isnan(v)) return 0.0f;
  if (v == INFINITY) return 3.4028234663852886e38f;
  if (v == -INFINITY) return -3.4028234663852886e38f;
  return v;
}


// Reductions of the fused launches, by all 16 epilogue warps, once pass Z has left exp(lp) (posterior) or lp (log_prob with
// error samples) of the chunk's rows in the state column COL_PEND.
__device__ __forceinline__ void tc_fused_reduction(const TcLaunch& lp, const LaunchArgs& args, const TcSmem& sl, uint8_t* base,
                                                   float* state, int CR, int COL_PEND, int nrows, long long unit0, int sub,
                                                   int tid, int warp, int lane) {
  // ---- posterior reduction: mean over the error samples, sum over the variants, per-galaxy trapezoid, NaN -> 0
  // (pzflow/flow.py:720-737, 651-654), by all 16 epilogue warps, before the ONLY write of the pdfs to HBM ----
  named_bar_sync(1, TC_EPI_THREADS);
  const float* P = state + COL_PEND * CR;
  const int G = args.G, S = max(args.S, 1), V = max(args.V, 1);
  const bool inner_nan0 = args.V > 0 && args.nan_to_zero != 0;  // an inner call's nan_to_num, per variant
  const bool lme = args.mode == MODE_LOGPROB;  // log(mean_s exp(lp)) of a row's error samples (flow.py:463-464), G == 1
  const int etid = (int)tid;
  auto finish = [&](const float* Yg, long long unit) {  // one warp: trapezoid of Yg[0, G), then the galaxy's pdf
    float integral = 1.0f;
    if (args.normalize) {
      float part = 0.0f;
      for (int i = lane; i + 1 < G; i += 32) part += (args.grid[i + 1] - args.grid[i]) * (Yg[i + 1] + Yg[i]);
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      integral = 0.5f * part;
    }
    return integral;
  };
  if (lp.upc > 0) {
    float* Y = state + (COL_PEND + 1) * CR;
    const int RU = lp.unit_rows, nunits = nrows / RU;
    for (int e = etid; e < nunits * G; e += TC_EPI_THREADS) {
      const int u = e / G, g = e - u * G;
      const float* pu = P + u * RU + g;
      float tot = 0.0f;
      for (int v = 0; v < V; ++v) {
        float sm = lme ? expf(pu[(v * S) * G]) : pu[(v * S) * G];
        for (int k = 1; k < S; ++k) sm += lme ? expf(pu[(v * S + k) * G]) : pu[(v * S + k) * G];
        float y = (S > 1 || lme) ? sm / (float)S : sm;
        if (inner_nan0) y = tc_nan_to_num0(y);
        tot = v == 0 ? y : tot + y;
      }
      if (lme) args.out0[unit0 + u] = logf(tot);
      else Y[e] = tot;
    }
    named_bar_sync(1, TC_EPI_THREADS);
    for (int u = warp; u < nunits && !lme; u += TC_EPI_THREADS / 32) {
      const float* Yg = Y + u * G;
      const float integral = finish(Yg, unit0 + u);
      float* dst = args.out0 + (unit0 + u) * (long long)G;
      for (int i = lane; i < G; i += 32) {
        float y = Yg[i];
        if (args.normalize) y = y / integral;
        if (args.nan_to_zero) y = tc_nan_to_num0(y);
        dst[i] = y;
      }
    }
  } else {
    float* accS = reinterpret_cast<float*>(base + sl.acc);
    float* accV = accS + G;
    const int nsamp = nrows / G, j0 = sub * lp.spc;
    for (int g = etid; g < G; g += TC_EPI_THREADS) {
      float aS = accS[g], aV = accV[g];
      for (int j = 0; j < nsamp; ++j) {
        const int si = j0 + j, k = si % S, v = si / S;
        const float pv = lme ? expf(P[j * G + g]) : P[j * G + g];
        aS = k == 0 ? pv : aS + pv;
        if (k == S - 1) {
          float y = (S > 1 || lme) ? aS / (float)S : aS;
          if (inner_nan0) y = tc_nan_to_num0(y);
          aV = v == 0 ? y : aV + y;
        }
      }
      accS[g] = aS;
      accV[g] = aV;
    }
    if (sub == lp.cpu - 1 && lme) {
      if (etid == 0) args.out0[unit0] = logf(accV[0]);
    } else if (sub == lp.cpu - 1) {
      named_bar_sync(1, TC_EPI_THREADS);
      const float integral = finish(accV, unit0);  // every warp computes the same sum in the same order
      float* dst = args.out0 + unit0 * (long long)G;
      for (int g = etid; g < G; g += TC_EPI_THREADS) {
        float y = accV[g];
        if (args.normalize) y = y / integral;
        if (args.nan_to_zero) y = tc_nan_to_num0(y);
        dst[g] = y;
      }
    }
  }
  named_bar_sync(1, TC_EPI_THREADS);  // P / Y / the sums are free again before the next chunk's first stage ends
}

#define TC_BAR(field) (sh_addr + (uint32_t)offsetof(TcShared, field))

// Wait of the four warps of one (group, half) on an mbarrier.  With PZB_TC_LEADER_WAIT only the half's first warp
// watches the mbarrier (try_wait compiles to a SYNCS / NANOSLEEP / BRA polling loop: a quarter of all executed
// instructions of the first version of this kernel); the other three block on a hardware named barrier, which issues
// nothing while it waits.
#ifdef PZB_TC_LEADER_WAIT
#define TC_HALF_WAIT(bar, parity)                              \
  do {                                                         \
    if ((warp & 3) == 0) mbar_wait((bar), (parity));           \
    named_bar_sync(4 + 2 * grp + half, TC_TILE);               \
  } while (0)
#else
#define TC_HALF_WAIT(bar, parity) mbar_wait((bar), (parity))
#endif

#ifdef PZB_TC_TRACE
// debug build only: SM-clock timestamps of the epilogue phases of CTA 0, second coupling stage of its first chunk
__device__ unsigned int g_tc_trace[19 * 8 * 8];
__device__ unsigned int g_tc_trace2[16 * 8];
#define TC_TRACE_S(ev)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && lane == 0 && seq == 2) g_tc_trace2[warp * 8 + (ev)] = (unsigned int)clock64(); \
  } while (0)
#define TC_TRACE(ev)                                                                         \
  do {                                                                                       \
    if (blockIdx.x == 0 && lane == 0 && (seq == 1 || seq == 2) && k < 4)                      \
      g_tc_trace[(warp * 8 + k + 4 * (seq - 1)) * 8 + (ev)] = (unsigned int)clock64();        \
  } while (0)
#define TC_TRACE_I(ev)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && lane == 0 && (s == 1 || s == 2) && k < 4)                          \
      g_tc_trace[(warp * 8 + k + 4 * (s - 1)) * 8 + (ev)] = (unsigned int)clock64();          \
  } while (0)
#else
#define TC_TRACE_S(ev) \
  do {                 \
  } while (0)
#define TC_TRACE_I(ev) \
  do {                 \
  } while (0)
#define TC_TRACE(ev) \
  do {               \
  } while (0)
#endif

template <bool INV>
__global__ void __launch_bounds__(TC_THREADS, 1)
nsf_chain_tc_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, TcLaunch lp,
                    unsigned long long* __restrict__ overflow_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, derived by pointer arithmetic so loads stay LDS (not generic LD)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool fuse = lp.unit_rows > 0;
  const TcSmem sl = tc_smem_layout(lp.max_pass, lp.chunk_tiles, lp.state_cols, (fuse && lp.upc == 0) ? 2 * args.G : 0);
  TcShared* sh = reinterpret_cast<TcShared*>(base + sl.shared);
  const uint32_t sh_addr = smem_u32(sh);
  float* state = reinterpret_cast<float*>(base + sl.state);
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));  // volatile: read once, never rematerialised as S2R
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  const int D = plan->D, C = plan->C, ns = plan->n_steps, n_lev = lp.n_lev;
  const int CR = lp.chunk_tiles * TC_TILE;           // rows of state per column
  const int COL_PEND = D + n_lev, COL_COND = D + n_lev + 2;

  if (tid == 0) {
    for (int g = 0; g < 2; ++g) {
      for (int q = 0; q < 4; ++q) mbar_init(TC_BAR(d2_full[g][q]), 1);
      for (int h = 0; h < 2; ++h) {
        mbar_init(TC_BAR(e_ready[g][h]), TC_TILE / 32);
        mbar_init(TC_BAR(d3_full[g][h]), 1);
        mbar_init(TC_BAR(d3_free[g][h]), TC_TILE / 32);
      }
      mbar_init(TC_BAR(par_full[g]), 1);
      mbar_init(TC_BAR(par_free[g]), TC_EPI_THREADS / 32);
      mbar_init(TC_BAR(k_ready[g]), TC_GROUP_THREADS / 32);
      mbar_init(TC_BAR(d1_full[g]), 1);
      mbar_init(TC_BAR(g3_done[g]), 1);
    }
    mbar_init(TC_BAR(w1_full), 1);
    mbar_init(TC_BAR(w2_full), 1);
    mbar_init(TC_BAR(w3_full), 1);
    mbar_init(TC_BAR(w1_free), 2);
    mbar_init(TC_BAR(w2_free), 2);
    mbar_init(TC_BAR(w3_free), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int seen = 0;
    for (int si = 0; si < ns; ++si) {
      const StepDev step = plan->steps[INV ? ns - 1 - si : si];
      if (step.kind != STEP_NSC) continue;
      for (int sub = 0; sub < plan->nsc[step.idx].n_sub; ++sub) {
        sh->order[seen] = (uint8_t)step.idx;
        sh->sub[seen++] = (uint8_t)sub;
      }
    }
    sh->n_seq = seen;
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(TC_BAR(tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sh->tmem_base;

  const int my_chunks = tc_my_chunks(lp, args);

  if (warp < 16) {
    // ======================================= epilogue warps =======================================
    const int grp = (warp >> 3) & 1;                   // epilogue group
    const int half = (warp >> 2) & 1;                  // which half of the row's work
    const int lrow = ((warp & 3) << 5) + lane;         // row within the tile == TMEM lane
    const int gtid = (int)(tid & (TC_GROUP_THREADS - 1));  // thread index within the group
    const uint32_t t_x = tbase + ((uint32_t)((warp & 3) << 5) << 16) + (uint32_t)grp * 256u + (uint32_t)half * 64u;
    const uint32_t t_y = t_x + 128u;                   // this half's 64 columns of X / Y
    const uint32_t bar_ready = TC_BAR(e_ready[0][0]) + 16u * grp + 8u * half;
    const uint32_t bar_d2 = TC_BAR(d2_full[0][0]) + 32u * grp + 16u * half;  // + 8 * batch
    const uint32_t bar_d3 = TC_BAR(d3_full[0][0]) + 16u * grp + 8u * half;
    const uint32_t bar_free = TC_BAR(d3_free[0][0]) + 16u * grp + 8u * half;
    const uint32_t t_key = tbase + ((uint32_t)((warp & 3) << 5) << 16) + (uint32_t)grp * 256u + 128u;  // Y[0, 24)
    const uint32_t bar_key = TC_BAR(k_ready[0]) + 8u * grp, bar_d1 = TC_BAR(d1_full[0]) + 8u * grp;
    const uint32_t bar_g3 = TC_BAR(g3_done[0]) + 8u * grp;
    uint32_t ph_d1 = 0u, ph_d2 = 0u, ph_d3 = 0u, ph_g3 = 0u;
    int seq = 0;  // coupling stages processed so far by this CTA (over all its chunks)

    // NSC steps in walk order
    int first_nsc = ns, last_nsc = -1;
    for (int si = 0; si < ns; ++si)
      if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
        if (first_nsc == ns) first_nsc = si;
        last_nsc = si;
      }

    for (int ci = 0; ci < my_chunks; ++ci) {
      long long row0, unit0;
      int nrows, sub;
      tc_chunk_span(lp, args, ci, row0, nrows, unit0, sub);
      const int ntiles = max(2, (nrows + TC_TILE - 1) / TC_TILE);
      const int my_tiles = (ntiles - grp + 1) / 2;

      // ---- pass A: rows of this group's tiles -> state, steps in front of the first coupling stage ----
      // (the two groups never wait for each other: they only meet through the weight buffers)
      uint32_t in_ok = 0u;  // bit i: the i-th row this thread loads (and finishes in pass Z) has finite inputs
      for (int q = gtid, it = 0; q < my_tiles * TC_TILE; q += TC_GROUP_THREADS, ++it) {
        const int r = (2 * (q >> 7) + grp) * TC_TILE + (q & (TC_TILE - 1));
        if (r >= nrows) continue;
        const long long row = row0 + r;
        float* st = state + r;
        float chk = 0.0f;
        const RowIndex ri = pz_row_index(args, row);
        for (int j = 0; j < D; ++j) {
          const float v = pz_input_value(args, ri, j, D);
          st[(INV ? plan->perm_final[j] : j) * CR] = v;
          chk += v * 0.0f;   // NaN iff some input is NaN or infinite
        }
        for (int l = 0; l < n_lev; ++l) st[(D + l) * CR] = 0.0f;
        for (int j = 0; j < C; ++j) {
          const float v = pz_cond_value(args, ri, j, C);
          st[(COL_COND + j) * CR] = v;
          chk += v * 0.0f;
        }
        in_ok |= (chk == 0.0f ? 1u : 0u) << it;
        int level = 0;
        light_steps_cols(plan, INV, 0, first_nsc, level, st, CR);
      }
      for (int q = gtid; q < my_tiles * (TC_TILE / 32); q += TC_GROUP_THREADS)   // this group's tiles' flag words
        sh->ovf[(2 * (q >> 2) + grp) * (TC_TILE / 32) + (q & 3)] = 0u;
      int level = level_after(plan, INV, 0, first_nsc, 0);  // chain depth (row independent)
      named_bar_sync(2 + grp, TC_GROUP_THREADS);

      // ---- the coupling stages ----
      int prev_end = first_nsc;
      bool pending = false;  // the previous stage's two log-det halves still sit in the pending columns
      int pend_level = 0;
      for (int si = first_nsc; si <= last_nsc; ++si) {
        if (plan->steps[INV ? ns - 1 - si : si].kind != STEP_NSC) continue;
        const bool has_light = si > prev_end;
        const int light0 = prev_end;
        const int level_in = level;
        if (has_light) level = level_after(plan, INV, prev_end, si, level);
        prev_end = si + 1;

        // a stage with more than four spline dimensions runs as sub-stages of four (same conditioner inputs, own blob)
        const int n_sub = plan->nsc[plan->steps[INV ? ns - 1 - si : si].idx].n_sub;
        for (int sub = 0; sub < n_sub; ++sub) {
          const bool light_now = has_light && sub == 0;
          // parameter block of this stage (double buffered); previous-stage state writes of the row's other half
          TC_TRACE_S(0);
          mbar_wait(TC_BAR(par_full[0]) + 8u * (seq & 1), (uint32_t)((seq >> 1) & 1));
          TC_TRACE_S(1);
          named_bar_sync(2 + grp, TC_GROUP_THREADS);
          TC_TRACE_S(2);
          const float* par = reinterpret_cast<const float*>(base + sl.par + (seq & 1) * TC_PAR_BYTES);
          const int* pint = reinterpret_cast<const int*>(par + TC_PI);
          const int L = pint[TC_PI_L];
          // The half that evaluates the last spline dimension finishes a tile last; the OTHER half writes the next
          // tile's key operand once it is through with its own dimensions.
          const int kw = L & 1;

          // Between two coupling stages: fold the previous stage's log-det halves into their chain level and apply the
          // non-NSC steps (e.g. an affine inside the chain), for every tile of this group, before any key is read.
          if (pending || light_now) {
            for (int q = gtid; q < my_tiles * TC_TILE; q += TC_GROUP_THREADS) {
              float* st = state + (2 * (q >> 7) + grp) * TC_TILE + (q & (TC_TILE - 1));
              if (pending) st[(D + pend_level) * CR] += st[COL_PEND * CR] + st[(COL_PEND + 1) * CR];
              if (light_now) {
                int lv = level_in;
                light_steps_cols(plan, INV, light0, si, lv, st, CR);
              }
            }
            if (light_now) named_bar_sync(2 + grp, TC_GROUP_THREADS);
          }

          // E0: the <= 15 conditioner inputs of a row and a constant 1 (bias row) as a K = 16 fp16 hi/lo A operand in
          // Y[0,8) / Y[16,24).  One thread per row (half kw) writes it; the tile's first-layer GEMM then runs on the
          // tensor pipe.  For tile k + 1 this happens while tile k's last spline dimension is still being evaluated.
          auto write_keys = [&](int kk) {
            const float* krow = state + (2 * kk + grp) * TC_TILE + lrow;
            const int in_dim = pint[TC_PI_IN];
            float key[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) key[i] = (i < in_dim) ? krow[pint[TC_PI_KEY + i] * CR] : 0.0f;
            float chk = ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + (key[6] + key[7]));
            uint32_t rh[8], rl[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) split2(key[2 * j], key[2 * j + 1], rh[j], rl[j]);
            if (in_dim <= 8) {  // (uniform branch) the usual case: k = 8 is the constant 1 of the bias row, k > 8 unused
              rh[4] = 0x00003C00u;
              rl[4] = 0u;
#pragma unroll
              for (int j = 5; j < 8; ++j) rh[j] = rl[j] = 0u;
            } else {            // inputs 8..14 ride in k = 9..15
#pragma unroll
              for (int i = 0; i < 7; ++i) key[i] = (8 + i < in_dim) ? krow[pint[TC_PI_KEY + 8 + i] * CR] : 0.0f;
              chk += ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + key[6]);
              split2(1.0f, key[0], rh[4], rl[4]);
#pragma unroll
              for (int j = 0; j < 3; ++j) split2(key[1 + 2 * j], key[2 + 2 * j], rh[5 + j], rl[5 + j]);
            }
            tmem_st8(t_key, rh);
            tmem_st8(t_key + 16u, rl);
            // a sum of finite inputs of this size is finite; NaN or infinite inputs make it non-finite (inf - inf -> NaN)
            const uint32_t ok = __ballot_sync(0xffffffffu, fabsf(chk) < INFINITY);
            if (lane == 0) sh->kok[((2 * kk + grp) * TC_TILE + lrow) >> 5] = ok;
            tc_wait_st();
          };
          TC_TRACE_S(3);
          if (half == kw) write_keys(0);
          tc_fence_before();
          warp_arrive(bar_key, lane);
          TC_TRACE_S(4);

          for (int k = 0; k < my_tiles; ++k) {
            const int r = (2 * k + grp) * TC_TILE + lrow;  // row within the chunk (state is allocated for every tile)
            float* srow = state + r;

            // ---- E1: h1 = LeakyReLU(D1) (bias already in the GEMM), rewritten in place over this half's 64 columns of X ----
            TC_TRACE(0);
            TC_HALF_WAIT(bar_d1, ph_d1);
            ph_d1 ^= 1u;
            tc_fence_after();
#pragma unroll 1
            for (int bt = 0; bt < 2; ++bt) {
              float dv[32];
              tmem_ld32(t_x + bt * 32, dv);
              tc_wait_ld();
              uint32_t rh[16], rl[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
#if PZB_TC_STUB & 2   // measurement build only: the LeakyReLU + hi/lo split removed, TMEM traffic kept
                rh[j] = 0x2c002c00u + (__float_as_uint(dv[2 * j]) & 0x00ff00ffu);
                rl[j] = __float_as_uint(dv[2 * j + 1]) & 0x03ff03ffu;
#else
                float h0, h1;
                leaky2(dv[2 * j], dv[2 * j + 1], h0, h1);
                split2(h0, h1, rh[j], rl[j]);
#endif
              }
              tmem_st16(t_x + bt * 32, rh);
              tmem_st16(t_x + bt * 32 + 16, rl);
            }
            tc_wait_st();
            tc_fence_before();
            warp_arrive(bar_ready, lane);
            TC_TRACE(1);

            // ---- E2: h2 = LeakyReLU(D * 2^-e2 + b2), rewritten in place over this half's accumulator columns ----
            {
              const float s2inv = par[TC_PF_SCAL];
#pragma unroll 1
              for (int bt = 0; bt < 2; ++bt) {
                TC_HALF_WAIT(bar_d2 + 8u * bt, ph_d2);  // this quarter of the hidden GEMM has landed
                tc_fence_after();
                if (bt == 0) TC_TRACE(2);
                float dv[32];
                tmem_ld32(t_y + bt * 32, dv);
                tc_wait_ld();
                const float* b2 = par + TC_PF_B2 + half * 64 + bt * 32;
                uint32_t rh[16], rl[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 bb = *reinterpret_cast<const float4*>(b2 + q * 4);
                  const float2 sc = make_float2(s2inv, s2inv);  // packed fp32x2 FMA: the same roundings as two FFMAs
                  const float2 p01 = __ffma2_rn(make_float2(dv[4 * q + 0], dv[4 * q + 1]), sc, make_float2(bb.x, bb.y));
                  const float2 p23 = __ffma2_rn(make_float2(dv[4 * q + 2], dv[4 * q + 3]), sc, make_float2(bb.z, bb.w));
#if PZB_TC_STUB & 2
                  rh[2 * q] = 0x2c002c00u + (__float_as_uint(p01.x) & 0x00ff00ffu);
                  rl[2 * q] = __float_as_uint(p01.y) & 0x03ff03ffu;
                  rh[2 * q + 1] = 0x2c002c00u + (__float_as_uint(p23.x) & 0x00ff00ffu);
                  rl[2 * q + 1] = __float_as_uint(p23.y) & 0x03ff03ffu;
#else
                  float h0, h1, h2, h3;
                  leaky2(p01.x, p01.y, h0, h1);
                  leaky2(p23.x, p23.y, h2, h3);
                  split2(h0, h1, rh[2 * q], rl[2 * q]);
                  split2(h2, h3, rh[2 * q + 1], rl[2 * q + 1]);
#endif
                }
                tmem_st16(t_y + bt * 32, rh);
                tmem_st16(t_y + bt * 32 + 16, rl);
              }
            }
            ph_d2 ^= 1u;
            tc_wait_st();
            tc_fence_before();
            warp_arrive(bar_ready, lane);
            TC_TRACE(3);

            // ---- E3: logits -> knots -> spline; this thread takes the dimensions half, half + 2 ----
            const bool more = k + 1 < my_tiles;
            float ldsum = 0.0f;
#pragma unroll 1
            for (int d = half; d < L; d += 2) {
              TC_HALF_WAIT(bar_d3, ph_d3);
              ph_d3 ^= 1u;
              tc_fence_after();
              TC_TRACE(d < 2 ? 4 : 6);
              float raw[48];
#if PZB_TC_PAIR_PASS
              tmem_ld32(t_x - 16u * (uint32_t)half, raw);          // pair passes: dims 2p, 2p + 1 land in X[0,48) / X[48,96)
              tmem_ld16(t_x - 16u * (uint32_t)half + 32, raw + 32);
#else
              tmem_ld32(t_x, raw);
              tmem_ld16(t_x + 32, raw + 32);
#endif
              tc_wait_ld();
              tc_fence_before();
              warp_arrive(bar_free, lane);  // accumulator buffer drained: the next pass may overwrite it
              if (d + 2 >= L && more && half != kw) warp_arrive(bar_key, lane);  // my last dimension: X is drained on my side
              const int col = pint[TC_PI_LOW + d];
              const float xin = srow[col * CR];
              float y, ldd;
              int idx;
              bool fin;
#if PZB_TC_STUB & 1   // measurement build only (results are garbage): the spline arithmetic removed, its TMEM reads kept
              y = xin + (raw[0] + raw[47]) * 1e-30f;
              ldd = raw[20] * 1e-30f;
              idx = 0;
              fin = true;
#else
              spline_dim<INV>(raw, par + TC_PF_B3 + d * TC_DIM_N, par[TC_PF_SCAL + 1], xin, par[TC_PF_SCAL + 2],
                              par[TC_PF_SCAL + 3], pint[TC_PI_PERIODIC] != 0, y, ldd, idx, fin);
#endif
              if (!fin && ((sh->kok[r >> 5] >> (r & 31)) & 1u))
                atomicOr(&sh->ovf[r >> 5], 1u << (r & 31));   // (rare) finite inputs, non-finite logits: pzb_plan_overflow_rows
              ldsum += ldd;
              srow[col * CR] = y;
              TC_TRACE(d < 2 ? 5 : 7);
              if (args.bins != nullptr && r < nrows)
                // (K = beyond the last knot; the padding bins of a K < 16 stage count as that)
                args.bins[(row0 + r) * plan->n_spline_dims + pint[TC_PI_BINS] + d] = min(idx, pint[TC_PI_K]);
            }
            srow[(COL_PEND + half) * CR] = INV ? -ldsum : ldsum;
            if (half == kw) {
              // every output pass of this tile has completed, so Y is free: write the next tile's key operand
              TC_HALF_WAIT(bar_g3, ph_g3);
              if (more) {
                tc_fence_after();
                write_keys(k + 1);
                tc_fence_before();
                warp_arrive(bar_key, lane);
              }
            }
            ph_g3 ^= 1u;  // g3_done completes once per tile: BOTH halves count it (kw changes with the stage's L)
          }
          warp_arrive(TC_BAR(par_free[0]) + 8u * (seq & 1), lane);
          pending = true;
          pend_level = level;
          ++seq;
        }
      }
      named_bar_sync(2 + grp, TC_GROUP_THREADS);

      // ---- pass Z: trailing steps, latent log-density, outputs of this group's tiles ----
      uint32_t n_bad = 0u;  // rows of finite inputs whose logits came out non-finite: an fp16 overflow in a conditioner (pzb_plan_overflow_rows)
      for (int q = gtid, it = 0; q < my_tiles * TC_TILE; q += TC_GROUP_THREADS, ++it) {
        const int r = (2 * (q >> 7) + grp) * TC_TILE + (q & (TC_TILE - 1));
        if (r >= nrows) continue;
        const long long row = row0 + r;
        float* st = state + r;
        if (pending) st[(D + pend_level) * CR] += st[COL_PEND * CR] + st[(COL_PEND + 1) * CR];
        int lv = level;
        light_steps_cols(plan, INV, prev_end, ns, lv, st, CR);
        n_bad += ((sh->ovf[r >> 5] >> (r & 31)) & (in_ok >> it) & 1u);
        if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
          const float lat = pz_latent_logprob(plan, [&](int j) { return st[plan->perm_final[j] * CR]; });
          const float lpv = pz_nan_to_num_lp(lat + st[D * CR]);
          if (fuse) st[COL_PEND * CR] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;  // reduced below, in shared memory
          else args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
        } else {
          for (int j = 0; j < D; ++j) args.out0[row * D + j] = st[(INV ? j : plan->perm_final[j]) * CR];
          if (args.out1 != nullptr) args.out1[row] = st[D * CR];
        }
      }
      if (overflow_rows != nullptr && __any_sync(0xffffffffu, n_bad != 0u)) {
        for (int o = 16; o > 0; o >>= 1) n_bad += __shfl_xor_sync(0xffffffffu, n_bad, o);
        if (lane == 0) {
          atomicAdd(overflow_rows, (unsigned long long)n_bad);
          if (args.ovf_flag != nullptr) *args.ovf_flag = 1u;
        }
      }
      // (pass A of the next chunk touches a row from the same thread as pass Z did: no barrier needed)
      if (fuse) tc_fused_reduction(lp, args, sl, base, state, CR, COL_PEND, nrows, unit0, sub, (int)tid, warp, lane);
    }
  } else {
    // control warps run converged (all 32 lanes wait on the mbarriers); one elected lane issues
    const uint32_t sb = smem_u32(base);
    const int n_seq = sh->n_seq;
    const int total_seq = my_chunks * n_seq;
    // coupling stage (walk order) and sub-stage of sequence number s, and the tile count of its chunk
    auto stage_of = [&](int s) -> const NscDev& {
      return plan->nsc[__shfl_sync(0xffffffffu, (int)sh->order[s % n_seq], 0)];
    };
    auto sub_of = [&](int s) -> int { return __shfl_sync(0xffffffffu, (int)sh->sub[s % n_seq], 0); };
    auto tiles_of = [&](int s) -> int {
      long long row0, unit0;
      int nrows, sub;
      tc_chunk_span(lp, args, s / n_seq, row0, nrows, unit0, sub);
      return max(2, (nrows + TC_TILE - 1) / TC_TILE);
    };
    if (warp < 18) {
      // =============================== MMA issuer of group g ===============================
      const int g = __shfl_sync(0xffffffffu, warp, 0) - 16;
      const uint32_t t_x = __shfl_sync(0xffffffffu, tbase, 0) + (uint32_t)g * 256u, t_y = t_x + 128u;
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc3 = (1u << 4) | ((uint32_t)(TC_DIM_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bar_ready = TC_BAR(e_ready[0][0]) + 16u * g;
      const uint32_t bar_d2 = TC_BAR(d2_full[0][0]) + 32u * g;
      const uint32_t bar_d3 = TC_BAR(d3_full[0][0]) + 16u * g;
      const uint32_t bar_free = TC_BAR(d3_free[0][0]) + 16u * g;
      const uint32_t w1 = sb + sl.w1, w2 = sb + sl.w2, w3 = sb + sl.w3;
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bar_key = TC_BAR(k_ready[0]) + 8u * g, bar_d1 = TC_BAR(d1_full[0]) + 8u * g;
      uint32_t ph_ready = 0u, ph_key = 0u;
      uint32_t np0 = 0u, np1 = 0u;  // passes issued so far into each accumulator buffer
      for (int s = 0; s < total_seq; ++s) {
        const int L = __shfl_sync(0xffffffffu, tc_sub_dims(stage_of(s).L, sub_of(s)), 0);
        const int my_tiles = (tiles_of(s) - g + 1) / 2;
        for (int k = 0; k < my_tiles; ++k) {
          // ---- first layer: keys (Y[0,24), K = 16) x W1 tile -> X, fp32, bias included ----
          if (k == 0) mbar_wait(TC_BAR(w1_full), (uint32_t)(s & 1));
          mbar_wait(bar_key, ph_key);
          ph_key ^= 1u;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dh = make_b_desc_w1(w1), dl = make_b_desc_w1(w1 + TC_W1_BYTES / 2);
            umma_f16_ts<false>(t_x, t_y + 16u, dh, idesc1);  // lo * hi
            umma_f16_ts<true>(t_x, t_y, dl, idesc1);         // hi * lo
            umma_f16_ts<true>(t_x, t_y, dh, idesc1);         // hi * hi
            umma_commit(bar_d1);
            if (k == my_tiles - 1) umma_commit(TC_BAR(w1_free));
          }
          __syncwarp();
          // ---- hidden GEMM, two N = 64 halves: X (h1) x W2 -> Y ----
          if (k == 0) mbar_wait(TC_BAR(w2_full), (uint32_t)(s & 1));
          TC_TRACE_I(0);
          mbar_wait(bar_ready, ph_ready);
          mbar_wait(bar_ready + 8u, ph_ready);
          ph_ready ^= 1u;
          tc_fence_after();
          TC_TRACE_I(1);
          if (elect_one()) {
#if PZB_TC_G2_HALVES
            // experiment: two N = 64 halves (half the MMA instructions; E2 of a half starts when all its 64 columns are there)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int n0 = h * 64;
              const uint32_t id64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
              issue_gemm<0, 8, true>(t_y + (uint32_t)n0, t_x, w2 + n0 * 128, w2 + TC_W2_BYTES / 2 + n0 * 128, 128 * 128, id64);
              umma_commit(bar_d2 + 8u * (2 * h));
              umma_commit(bar_d2 + 8u * (2 * h + 1));
            }
#else
            // four N = 32 quarters, alternating between the halves so both start their E2 early
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int h = q & 1, bt = q >> 1, n0 = h * 64 + bt * 32;
              issue_gemm<0, 8, true>(t_y + (uint32_t)n0, t_x, w2 + n0 * 128, w2 + TC_W2_BYTES / 2 + n0 * 128, 128 * 128, idesc2);
              umma_commit(bar_d2 + 8u * (2 * h + bt));
            }
#endif
            if (k == my_tiles - 1) umma_commit(TC_BAR(w2_free));  // W2 is free once this group's last hidden GEMM is done
          }
          __syncwarp();
          // ---- output layer, one N = 48 pass per spline dimension: Y (h2) x W3 -> X[0,48) / X[64,112).  Half 0
          // finishes its E2 first (its GEMM half commits first): the first two passes start on its 64 K-elements
          // and take the other 64 when half 1 arrives ----
          TC_TRACE_I(2);
          if (k == 0) mbar_wait(TC_BAR(w3_full), (uint32_t)(s & 1));
          mbar_wait(bar_ready, ph_ready);
          tc_fence_after();
          TC_TRACE_I(3);
#if PZB_TC_PAIR_PASS
          // the output layer as (at most) two passes of TWO spline dimensions each, N = 96 -- half the MMA
          // instructions of one pass per dimension (48 clk each instead of 24); dims 2p / 2p + 1 land in X[0,48) / X[48,96)
          {
            const int ndA = L < 2 ? L : 2, ndB = L - 2;
            const uint32_t idA = (1u << 4) | ((uint32_t)((TC_DIM_N * ndA) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            if (np0 > 0u) mbar_wait(bar_free, (np0 - 1u) & 1u);
            if (ndA == 2 && np1 > 0u) mbar_wait(bar_free + 8u, (np1 - 1u) & 1u);
            tc_fence_after();
            if (elect_one()) issue_gemm<0, 4, true>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idA);
            __syncwarp();
            mbar_wait(bar_ready + 8u, ph_ready);
            ph_ready ^= 1u;
            tc_fence_after();
            if (elect_one()) {
              issue_gemm<4, 8, false>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idA);
              umma_commit(bar_d3);
              if (ndA == 2) umma_commit(bar_d3 + 8u);
              if (L <= 2) umma_commit(TC_BAR(g3_done[0]) + 8u * g);
              if (L <= 2 && k == my_tiles - 1) umma_commit(TC_BAR(w3_free));
            }
            __syncwarp();
            np0 += 1u;
            np1 += ndA == 2 ? 1u : 0u;
            if (ndB > 0) {
              const uint32_t idB = (1u << 4) | ((uint32_t)((TC_DIM_N * ndB) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
              mbar_wait(bar_free, (np0 - 1u) & 1u);
              mbar_wait(bar_free + 8u, (np1 - 1u) & 1u);   // (dim 1's columns [48,96) are overwritten by a two-dim pass; harmless to wait for when ndB == 1)
              tc_fence_after();
              const uint32_t tile = w3 + TC_W3_PASS_BYTES;
              if (elect_one()) {
                issue_gemm<0, 8, true>(t_x, t_y, tile, tile + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idB);
                umma_commit(bar_d3);
                if (ndB == 2) umma_commit(bar_d3 + 8u);
                umma_commit(TC_BAR(g3_done[0]) + 8u * g);
                if (k == my_tiles - 1) umma_commit(TC_BAR(w3_free));
              }
              __syncwarp();
              np0 += 1u;
              np1 += ndB == 2 ? 1u : 0u;
            }
          }
#else
          const int n_early = L < 2 ? L : 2;
          for (int d = 0; d < n_early; ++d) {
            const uint32_t np = d ? np1 : np0;
            if (np > 0u) {  // the buffer's previous pass must have been drained
              mbar_wait(bar_free + 8u * d, (np - 1u) & 1u);
              tc_fence_after();
            }
            const uint32_t tile = w3 + (uint32_t)d * (TC_DIM_N * 128);
            if (elect_one()) issue_gemm<0, 4, true>(t_x + 64u * d, t_y, tile, tile + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idesc3);
            __syncwarp();
          }
          TC_TRACE_I(4);
          mbar_wait(bar_ready + 8u, ph_ready);
          ph_ready ^= 1u;
          tc_fence_after();
          TC_TRACE_I(5);
          for (int d = 0; d < n_early; ++d) {
            const uint32_t tile = w3 + (uint32_t)d * (TC_DIM_N * 128);
            if (elect_one()) {
              issue_gemm<4, 8, false>(t_x + 64u * d, t_y, tile, tile + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idesc3);
              umma_commit(bar_d3 + 8u * d);
              if (d == L - 1) umma_commit(TC_BAR(g3_done[0]) + 8u * g);
              if (d == L - 1 && k == my_tiles - 1) umma_commit(TC_BAR(w3_free));
            }
            __syncwarp();
            np0 += d ? 0u : 1u;
            np1 += d ? 1u : 0u;
          }
          TC_TRACE_I(6);
          for (int d = 2; d < L; ++d) {
            const int b = d & 1;
            const uint32_t np = b ? np1 : np0;
            mbar_wait(bar_free + 8u * b, (np - 1u) & 1u);  // the buffer's previous pass must have been drained
            tc_fence_after();
            if (d == L - 1) TC_TRACE_I(7);
            const uint32_t tile = w3 + (uint32_t)(d >> 1) * TC_W3_PASS_BYTES + (uint32_t)b * (TC_DIM_N * 128);
            if (elect_one()) {
              issue_gemm<0, 8, true>(t_x + 64u * b, t_y, tile, tile + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, idesc3);
              umma_commit(bar_d3 + 8u * b);
              if (d == L - 1) umma_commit(TC_BAR(g3_done[0]) + 8u * g);
              if (d == L - 1 && k == my_tiles - 1) umma_commit(TC_BAR(w3_free));
            }
            __syncwarp();
            np0 += b ? 0u : 1u;
            np1 += b ? 1u : 0u;
          }
#endif
        }
      }
    } else {
      // =============================== weight streamer ===============================
      for (int s = 0; s < total_seq; ++s) {
        const NscDev& st = stage_of(s);
        const int sub = sub_of(s);
        const TcLayout lay = TcLayout::make(tc_sub_dims(st.L, sub));
        const uint8_t* src = static_cast<const uint8_t*>(st.tc[sub].blob);
        {  // parameter block: buffer s & 1 was last used by stage s - 2
          if (s >= 2) mbar_wait(TC_BAR(par_free[0]) + 8u * (s & 1), (uint32_t)(((s >> 1) - 1) & 1));
          const uint32_t bar = TC_BAR(par_full[0]) + 8u * (s & 1);
          if (lane == 0) {
            mbar_expect_tx(bar, TC_PAR_BYTES);
            tma_bulk_g2s(sb + sl.par + (s & 1) * TC_PAR_BYTES, src + lay.par, TC_PAR_BYTES, bar);
          }
        }
        {
          if (s >= 1) mbar_wait(TC_BAR(w1_free), (uint32_t)((s - 1) & 1));
          const uint32_t bar = TC_BAR(w1_full);
          if (lane == 0) {
            mbar_expect_tx(bar, TC_W1_BYTES);
            tma_bulk_g2s(sb + sl.w1, src + lay.w1, TC_W1_BYTES, bar);
          }
        }
        {
          if (s >= 1) mbar_wait(TC_BAR(w2_free), (uint32_t)((s - 1) & 1));
          const uint32_t bar = TC_BAR(w2_full);
          if (lane == 0) {
            mbar_expect_tx(bar, TC_W2_BYTES);
            for (int off = 0; off < TC_W2_BYTES; off += 16384) tma_bulk_g2s(sb + sl.w2 + off, src + off, 16384u, bar);
          }
        }
        {
          if (s >= 1) mbar_wait(TC_BAR(w3_free), (uint32_t)((s - 1) & 1));
          const uint32_t bar = TC_BAR(w3_full);
          const int nb = lay.n_pass * TC_W3_PASS_BYTES;
          if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)nb);
            for (int off = 0; off < nb; off += 16384) tma_bulk_g2s(sb + sl.w3 + off, src + lay.w3 + off, 16384u, bar);
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

#include "nsf_tc_pipe3.cuh"

static int tc_fit_tiles(const TcLaunch& lp, int acc_floats) {
  int fit = TC_MAX_CHUNK_TILES;
  if (const char* cap = getenv("PZB_TC_CHUNK_TILES"))   // experiments only: fewer tiles per chunk than shared memory holds
    fit = std::max(2, std::min(fit, atoi(cap)));
  while (fit > 2 && tc_smem_layout(lp.max_pass, fit, lp.state_cols, acc_floats, lp.key_bytes).total + 1024 > TC_SMEM_MAX) --fit;
  return fit - (fit & 1);  // the two epilogue groups take alternate tiles: keep the count even
}

// Which of the two kernels runs: PZB_TC_PIPE3=1 selects nsf_chain_tc3_kernel (three tiles in flight), 0 the two-tile kernel.
static bool tc_pipe3_enabled() {   // read per launch: a test switches kernels inside one process
  const char* v = getenv("PZB_TC_PIPE3");
  return v != nullptr ? atoi(v) != 0 : PZB_TC_PIPE3_DEFAULT != 0;
}

static void tc_base_launch(const PlanDev& hp, TcLaunch& lp) {
  lp.max_pass = 1;
  for (int s = 0; s < hp.n_nsc; ++s) lp.max_pass = max(lp.max_pass, (min(hp.nsc[s].L, TC_L_MAX) + 1) / 2);
  lp.n_lev = tc_log_det_levels(hp);
  lp.key_bytes = tc_pipe3_enabled() ? TC_W1_BYTES : 0;
  lp.state_cols = hp.D + lp.n_lev + 2 + hp.C + (lp.key_bytes != 0 ? 2 : 0);   // (three-tile kernel: a log-det column per quarter)
  lp.unit_rows = lp.upc = lp.cpu = lp.spc = 0;
  lp.n_units = 0;
  if (lp.key_bytes != 0 && tc_fit_tiles(lp, 0) < 4) {   // the three-tile kernel walks at least four tiles per chunk: a flow
    lp.key_bytes = 0;                                    // whose state is too wide for that takes the two-tile kernel
    lp.state_cols -= 2;
  }
}

// Can the posterior of n_units galaxies (V variants x S error samples x G grid points each) be reduced inside the
// kernel?  0 = no (the caller takes the scratch + reduction-kernel route), 1 = whole galaxies per chunk, 2 = a galaxy
// spans several chunks of one CTA.
int tc_fuse_mode(const PlanDev& hp, int G, int V, int S, long long n_units, int num_sms) {
  if (!tc_plan_supported(hp) || G < 1 || n_units < 1) return 0;
  TcLaunch lp;
  tc_base_launch(hp, lp);
  const long long RU = (long long)(V > 0 ? V : 1) * (S > 0 ? S : 1) * G;
  if (RU <= (long long)tc_fit_tiles(lp, 0) * TC_TILE) return 1;
  // long galaxies are dealt to the CTAs whole: few of them would leave most SMs idle
  const int fit2 = tc_fit_tiles(lp, 2 * G);
  if (lp.key_bytes != 0 && fit2 < 4) return 0;   // (the three-tile kernel needs four tiles of state beside the sums)
  if ((long long)G <= (long long)fit2 * TC_TILE && n_units >= 2ll * num_sms) return 2;
  return 0;
}

cudaError_t launch_tc(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms, void* workspace,
                      cudaStream_t stream) {
  unsigned long long* overflow_rows = static_cast<unsigned long long*>(workspace);  // the plan's counter (may be null)
  TcLaunch lp;
  tc_base_launch(hp, lp);
  long long n_work;  // chunks (or whole long galaxies) to deal out
  int acc_floats = 0;
  if (args.fuse) {
    if (args.mode != MODE_POSTERIOR && !(args.mode == MODE_LOGPROB && args.G == 1)) return cudaErrorInvalidValue;
    const int V = args.V > 0 ? args.V : 1, S = args.S > 0 ? args.S : 1;
    const long long RU = (long long)V * S * args.G;
    lp.n_units = args.N / RU;
    lp.unit_rows = (int)RU;
    const int fit = tc_fit_tiles(lp, 0);
    if (RU <= (long long)fit * TC_TILE) {
      // whole galaxies per chunk: as many as fit, fewer when that leaves SMs idle
      long long upc = (long long)fit * TC_TILE / RU;
      const long long spread = (lp.n_units + num_sms - 1) / num_sms;
      if (upc > spread) upc = spread;
      if (upc < 1) upc = 1;
      lp.upc = (int)upc;
      int tiles = (int)((upc * RU + TC_TILE - 1) / TC_TILE);
      tiles += tiles & 1;
      lp.chunk_tiles = tiles < 2 ? 2 : tiles;
      n_work = (lp.n_units + upc - 1) / upc;
    } else {
      acc_floats = 2 * args.G;
      const int fit2 = tc_fit_tiles(lp, acc_floats);
      if ((long long)args.G > (long long)fit2 * TC_TILE) return cudaErrorInvalidConfiguration;
      lp.spc = fit2 * TC_TILE / args.G;
      const int vs = V * S;
      lp.cpu = (vs + lp.spc - 1) / lp.spc;
      lp.spc = (vs + lp.cpu - 1) / lp.cpu;  // even out the parts
      lp.cpu = (vs + lp.spc - 1) / lp.spc;
      int tiles = (int)(((long long)lp.spc * args.G + TC_TILE - 1) / TC_TILE);
      tiles += tiles & 1;
      lp.chunk_tiles = tiles < 2 ? 2 : tiles;
      n_work = lp.n_units;
    }
  } else {
    // rows per chunk: as many 128-row tiles as shared memory holds, fewer when that leaves SMs idle
    const int fit = tc_fit_tiles(lp, 0);
    const long long tiles_total = (args.N + TC_TILE - 1) / TC_TILE;
    long long want = (tiles_total + num_sms - 1) / num_sms;
    if (want < 2) want = 2;
    lp.chunk_tiles = (int)(want < fit ? want : fit);
    // the two epilogue groups take alternate tiles: an odd tile count leaves one group idle for a whole tile per stage
    if (lp.chunk_tiles > 2 && (lp.chunk_tiles & 1)) lp.chunk_tiles -= (want > fit) ? 1 : -1;
    if (lp.chunk_tiles > fit) lp.chunk_tiles = fit;
    const long long chunk_rows = (long long)lp.chunk_tiles * TC_TILE;
    n_work = (args.N + chunk_rows - 1) / chunk_rows;
  }
  if (lp.key_bytes != 0 && lp.chunk_tiles < 4) lp.chunk_tiles = 4;   // nsf_chain_tc3_kernel walks at least four tiles per chunk
  const TcSmem sl = tc_smem_layout(lp.max_pass, lp.chunk_tiles, lp.state_cols, acc_floats, lp.key_bytes);
  const size_t smem = (size_t)sl.total + 1024;
  if (smem > (size_t)TC_SMEM_MAX) return cudaErrorInvalidConfiguration;
  const bool inv = (args.mode == MODE_INVERSE);
  long long grid = num_sms;
  if (grid > n_work) grid = n_work;
  if (grid < 1) grid = 1;
  if (lp.key_bytes != 0) {
    cudaError_t e3 = inv ? cudaFuncSetAttribute(nsf_chain_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(nsf_chain_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e3 != cudaSuccess) return e3;
    if (inv)
      nsf_chain_tc3_kernel<true><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp, overflow_rows);
    else
      nsf_chain_tc3_kernel<false><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp, overflow_rows);
    return cudaGetLastError();
  }
  cudaError_t e = inv ? cudaFuncSetAttribute(nsf_chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(nsf_chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (inv)
    nsf_chain_tc_kernel<true><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp, overflow_rows);
  else
    nsf_chain_tc_kernel<false><<<(unsigned)grid, TC_THREADS, smem, stream>>>(d_plan, args, lp, overflow_rows);
  return cudaGetLastError();
}

}  // namespace pzb

#ifdef PZB_TC3_TRACE
extern "C" int pzb_debug_tc3_trace(unsigned int* issuer, unsigned int* epilogue) {
  cudaError_t e = cudaMemcpyFromSymbol(issuer, pzb::g_tc3_trace, sizeof(pzb::g_tc3_trace));
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(epilogue, pzb::g_tc3_trace_e, sizeof(pzb::g_tc3_trace_e));
  return (int)e;
}
#endif

#ifdef PZB_TC_TRACE
extern "C" int pzb_debug_tc_trace(unsigned int* out) {
  return (int)cudaMemcpyFromSymbol(out, pzb::g_tc_trace, sizeof(pzb::g_tc_trace));
}
extern "C" int pzb_debug_tc_trace2(unsigned int* out) {
  return (int)cudaMemcpyFromSymbol(out, pzb::g_tc_trace2, sizeof(pzb::g_tc_trace2));
}
#endif
