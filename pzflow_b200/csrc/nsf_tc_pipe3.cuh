// nsf_chain_tc3_kernel: the default-family tcgen05 engine of nsf_tc.cu with THREE tiles in flight per SM (experimental:
// PZB_TC_PIPE3=1; bit-identical to nsf_chain_tc_kernel, 5 % slower -- DESIGN.md section 4.5b says why).
// (Included by nsf_tc.cu inside namespace pzb, after its own kernel: it shares that file's packing, shared-memory
// layout, chunk geometry and the arithmetic of pzb_tc_common.cuh.  Same reference lines as nsf_simt.cu.)
//
// nsf_chain_tc_kernel gives each epilogue group a private X and Y region of TMEM (2 x 256 columns), so two tiles
// are in flight and the tensor pipe idles whenever both groups are in an epilogue phase (pipe ~64 % active).  A tile
// only needs BOTH regions from the start of its hidden GEMM to the end of its last output pass; before (first layer,
// E1) and after (the splines of its last pass) it lives in X alone.  So here TMEM holds three X regions and ONE
// shared Y:
//
//     columns [0,128) [128,256) [256,384): X0 X1 X2 (tile G uses X[G mod 3]);  [384,512): Y
//
// and the tensor pipe runs the tiles' "Y phases" back to back, one PERIOD per tile:
//
//     tensor pipe     G2(G) -> [G1(G+1)] -> G3(G) first pass (dims 0, 1) -> G3(G) second pass (dims 2, 3)
//     quarter 0       keys of tile G+1
//     quarters 0, 1   E2(G) -> E1(G+1) -> spline of their dimension of tile G      (first pass)
//     quarters 2, 3   spline of their dimension of tile G-1 (second pass) -> E2(G) -> E1(G+1)
//
// The 16 epilogue warps work as ONE team with FOUR threads per row: quarter q owns hidden columns [32 q, 32 q + 32) in
// E1 / E2 and spline dimension q in E3, so every phase is half as long as with two threads per row, and nobody waits for
// "its" tile: the tiles simply stream through.  ONE issuer warp orders all MMAs (Y's reuse is then plain program order on
// the tensor pipe).  A warp cannot issue much faster than one tcgen05.mma per ~20 clk next to 16 busy epilogue warps
// (N = 32 / N = 48 instructions, 16 / 24 clk on the pipe, went out at 19 / 32 clk), so the GEMMs are cut into FEWER, WIDER
// instructions than the two-tile kernel's: the hidden GEMM in two N = 64 halves, the output layer in two N = 96 passes of
// two spline dimensions; and the walk's bookkeeping runs one tile ahead, under a wait the issuer has anyway.  The K = 16
// key operand of the first layer moves to shared memory (8 KB, SWIZZLE_NONE, the layout of the W1 tile): there is no
// TMEM column left for it; tcgen05.mma reads it through a descriptor (.ss form).  A row's log-det terms of a stage sit in
// four state columns (one per quarter) and are folded into the chain level by the thread that writes the row's next keys,
// in the association of the two-tile kernel ((d0 + d2) + (d1 + d3)): same bits.
//
// mbarrier phases.  Tile G is the (G / 3)-th user of its X region: every barrier of a region (first layer landed, h1 written,
// each of the two output passes landed / drained -- a stage with one or two spline dimensions commits, and its threads
// drain, an EMPTY second pass) completes once per user, parity (G / 3) & 1.  Barriers of Y and of the key buffer complete
// once per tile: parity G & 1.  (Every waiter has, through the chain of commits it observed before, already seen the
// previous phase of its barrier complete, so a parity never aliases an older phase.)

#ifdef PZB_TC3_TRACE
#ifndef TC3_T0
#define TC3_T0 0
#endif
// debug build only: SM-clock stamps of CTA 0's issuer (16 events x 32 consecutive tiles from tile 12 on) and of one
// epilogue warp per (group, half) (pzb_debug_tc3_trace)
__device__ unsigned int g_tc3_trace[32 * 16];
__device__ unsigned int g_tc3_trace_e[4 * 32 * 16];
#define TC3_TI(ev)                                                                                      \
  do {                                                                                                  \
    if (blockIdx.x == 0 && lane == 0 && G >= TC3_T0 && G < TC3_T0 + 32) g_tc3_trace[(G - TC3_T0) * 16 + (ev)] = (unsigned int)clock64(); \
  } while (0)
#define TC3_TE(ev)                                                                                      \
  do {                                                                                                  \
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0 && G >= TC3_T0 && G < TC3_T0 + 32)                 \
      g_tc3_trace_e[(qt * 32 + (int)(G - TC3_T0)) * 16 + (ev)] = (unsigned int)clock64();     \
  } while (0)
#else
#define TC3_TI(ev) \
  do {             \
  } while (0)
#define TC3_TE(ev) \
  do {             \
  } while (0)
#endif

constexpr int TC3_MIN_TILES = 4;   // a row's stages must be two tile periods apart (its quarters hand over through the state)

#define TC3_BAR(field) (sh_addr + (uint32_t)offsetof(TcShared, p3) + (uint32_t)offsetof(TcPipe3Bars, field))

__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

template <bool INV>
__global__ void __launch_bounds__(TC_THREADS, 1)
nsf_chain_tc3_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, TcLaunch lp,
                     unsigned long long* __restrict__ overflow_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool fuse = lp.unit_rows > 0;
  const TcSmem sl = tc_smem_layout(lp.max_pass, lp.chunk_tiles, lp.state_cols, (fuse && lp.upc == 0) ? 2 * args.G : 0,
                                   TC_W1_BYTES);
  TcShared* sh = reinterpret_cast<TcShared*>(base + sl.shared);
  const uint32_t sh_addr = smem_u32(sh);
  float* state = reinterpret_cast<float*>(base + sl.state);
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  const int D = plan->D, C = plan->C, ns = plan->n_steps, n_lev = lp.n_lev;
  const int CR = lp.chunk_tiles * TC_TILE;
  const int COL_PEND = D + n_lev, COL_COND = D + n_lev + 2;

  if (tid == 0) {
    for (int r = 0; r < 3; ++r) {
      mbar_init(TC3_BAR(d1_full[r]), 1);
      for (int h = 0; h < 2; ++h) {
        mbar_init(TC3_BAR(d3_full[r][h]), 1);
        mbar_init(TC3_BAR(d3_free[r][h]), TC_GROUP_THREADS / 32);   // the two quarters of a pass
      }
      mbar_init(TC3_BAR(e1_ready[r]), TC_EPI_THREADS / 32);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(TC3_BAR(d2_full[h]), 1);
      mbar_init(TC3_BAR(e2_ready[h]), TC_GROUP_THREADS / 32);   // the two quarters of a K-half
      mbar_init(TC_BAR(par_full[h]), 1);
      mbar_init(TC_BAR(par_free[h]), TC_EPI_THREADS / 32);
    }
    mbar_init(TC3_BAR(key_ready), TC_TILE / 32);
    mbar_init(TC3_BAR(key_free), 1);
    mbar_init(TC_BAR(w1_full), 1);
    mbar_init(TC_BAR(w2_full), 1);
    mbar_init(TC_BAR(w3_full), 1);
    mbar_init(TC_BAR(w1_free), 1);   // one issuer
    mbar_init(TC_BAR(w2_free), 1);
    mbar_init(TC_BAR(w3_free), 1);
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
    // per (sub-)stage of the walk: the steps between it and the previous coupling stage [lt0, lt1), the chain depth in front
    // of those steps and at the stage (what the two-tile kernel's stage loop carries along)
    int first = ns;
    for (int si = 0; si < ns; ++si)
      if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
        first = si;
        break;
      }
    int level = level_after(plan, INV, 0, first, 0), prev_end = first, j = 0;
    for (int si = first; si < ns; ++si) {
      const StepDev step = plan->steps[INV ? ns - 1 - si : si];
      if (step.kind != STEP_NSC) continue;
      const bool has_light = si > prev_end;
      const int light0 = prev_end, level_in = level;
      if (has_light) level = level_after(plan, INV, prev_end, si, level);
      prev_end = si + 1;
      for (int sub = 0; sub < plan->nsc[step.idx].n_sub; ++sub, ++j) {
        sh->p3.lt0[j] = (uint16_t)((has_light && sub == 0) ? light0 : 0);
        sh->p3.lt1[j] = (uint16_t)((has_light && sub == 0) ? si : 0);
        sh->p3.lvl_in[j] = (int8_t)level_in;
        sh->p3.lvl[j] = (int8_t)level;
      }
    }
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
  const int n_seq = sh->n_seq;

  if (warp < 16) {
    // ======================================= epilogue warps: one team of 16 =======================================
    // FOUR threads per row: quarter qt = warp >> 2 owns hidden columns [32 qt, 32 qt + 32) in E1 / E2 and spline dimension
    // qt in E3 (dims 0, 1: the tile's first output pass, dims 2, 3: its second).  Per tile G every thread runs one PERIOD
    // (the tile's Y phase on the tensor pipe):
    //     quarter 0      keys of tile G+1 (E0: one thread per row; it also folds the row's log-det terms of the previous
    //                    stage and applies the steps between two coupling stages)
    //     quarters 0, 1  E2(G) -> E1(G+1) -> spline of their dimension of tile G          (first pass)
    //     quarters 2, 3  spline of their dimension of tile G-1 (second pass) -> E2(G) -> E1(G+1)
    const int qt = warp >> 2;                           // quarter of the row's work
    const int lrow = ((warp & 3) << 5) + lane;          // row within the tile == TMEM lane
    const int etid = (int)tid;
    const uint32_t t_row = tbase + ((uint32_t)((warp & 3) << 5) << 16);   // this thread's TMEM lane
    const uint32_t t_q = t_row + 32u * (uint32_t)qt;                       // + 128 * region: its 32 columns of X; + 384: of Y
    const uint32_t t_s = t_row + 48u * (uint32_t)(qt & 1);                 // + 128 * region: its dimension's 48 logits of a pass
    uint8_t* keyb = base + sl.keys;
    const int COL_PEND2 = COL_COND + C;   // log-det terms of quarters 2, 3 (quarters 0, 1: COL_PEND, COL_PEND + 1)
    const int my_pend = qt < 2 ? COL_PEND + qt : COL_PEND2 + (qt - 2);
    int seq0 = 0;                 // coupling (sub-)stages the CTA has walked before this chunk
    int tile_base = 0;            // tiles the CTA has issued before this chunk (X regions rotate through all of them)

    int first_nsc = ns, last_nsc = -1;
    for (int si = 0; si < ns; ++si)
      if (plan->steps[INV ? ns - 1 - si : si].kind == STEP_NSC) {
        if (first_nsc == ns) first_nsc = si;
        last_nsc = si;
      }
    const int level_z = sh->p3.lvl[n_seq - 1];                    // ... and at the last one

    for (int ci = 0; ci < my_chunks; ++ci) {
      long long row0, unit0;
      int nrows, sub;
      tc_chunk_span(lp, args, ci, row0, nrows, unit0, sub);
      const int ntiles = max(TC3_MIN_TILES, (nrows + TC_TILE - 1) / TC_TILE);
      const int T = n_seq * ntiles;   // tile periods of this chunk

      // ---- pass A: rows -> state, steps in front of the first coupling stage ----
      uint32_t in_ok = 0u;
      for (int r = etid, it = 0; r < ntiles * TC_TILE; r += TC_EPI_THREADS, ++it) {
        if (r >= nrows) continue;
        const long long row = row0 + r;
        float* st = state + r;
        float chk = 0.0f;
        const RowIndex ri = pz_row_index(args, row);
        for (int j = 0; j < D; ++j) {
          const float v = pz_input_value(args, ri, j, D);
          st[(INV ? plan->perm_final[j] : j) * CR] = v;
          chk += v * 0.0f;
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
      for (int w = etid; w < ntiles * (TC_TILE / 32); w += TC_EPI_THREADS) sh->ovf[w] = 0u;
      named_bar_sync(1, TC_EPI_THREADS);

      // parameter block of chunk stage j (double buffered over the CTA's stage sequence)
      auto par_of = [&](int j) -> const float* {
        return reinterpret_cast<const float*>(base + sl.par + ((seq0 + j) & 1) * TC_PAR_BYTES);
      };
      auto wait_par = [&](int j) { mbar_wait(TC_BAR(par_full[0]) + 8u * ((seq0 + j) & 1), (uint32_t)(((seq0 + j) >> 1) & 1)); };
      auto free_par = [&](int j) { warp_arrive(TC_BAR(par_free[0]) + 8u * ((seq0 + j) & 1), lane); };

      // E0 of chunk tile g (stage j = g / ntiles, tile t = g % ntiles), by quarter 0's thread of each row
      auto write_keys = [&](int g) {
        const int j = g / ntiles, t = g - j * ntiles;
        const int G = tile_base + g;
        const int r = t * TC_TILE + lrow;
        float* srow = state + r;
        if (t == 0) wait_par(j);
        const float* par = par_of(j);
        const int* pint = reinterpret_cast<const int*>(par + TC_PI);
        if (j > 0)   // the previous stage's four log-det terms of this row, in the association of the two-tile kernel
          srow[(D + sh->p3.lvl[j - 1]) * CR] += (srow[COL_PEND * CR] + srow[COL_PEND2 * CR]) +
                                                (srow[(COL_PEND + 1) * CR] + srow[(COL_PEND2 + 1) * CR]);
        if (sh->p3.lt0[j] < sh->p3.lt1[j]) {   // steps between two coupling stages (e.g. an affine inside the chain)
          int lv = sh->p3.lvl_in[j];
          light_steps_cols(plan, INV, sh->p3.lt0[j], sh->p3.lt1[j], lv, srow, CR);
        }
        if (G >= 1) mbar_wait(TC3_BAR(key_free), (uint32_t)((G - 1) & 1));
        const int in_dim = pint[TC_PI_IN];
        float key[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) key[i] = (i < in_dim) ? srow[pint[TC_PI_KEY + i] * CR] : 0.0f;
        float chk = ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + (key[6] + key[7]));
        uint32_t rh[8], rl[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2(key[2 * i], key[2 * i + 1], rh[i], rl[i]);
        if (in_dim <= 8) {
          rh[4] = 0x00003C00u;
          rl[4] = 0u;
#pragma unroll
          for (int i = 5; i < 8; ++i) rh[i] = rl[i] = 0u;
        } else {
#pragma unroll
          for (int i = 0; i < 7; ++i) key[i] = (8 + i < in_dim) ? srow[pint[TC_PI_KEY + 8 + i] * CR] : 0.0f;
          chk += ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + key[6]);
          split2(1.0f, key[0], rh[4], rl[4]);
#pragma unroll
          for (int i = 0; i < 3; ++i) split2(key[1 + 2 * i], key[2 + 2 * i], rh[5 + i], rl[5 + i]);
        }
        // K-major SWIZZLE_NONE, the W1 tile's layout: 8 x 8 cores of 128 B, K-adjacent cores 128 B apart, 8-row groups 256 B apart
        uint8_t* kp = keyb + (lrow >> 3) * 256 + (lrow & 7) * 16;
        *reinterpret_cast<uint4*>(kp) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
        *reinterpret_cast<uint4*>(kp + 128) = make_uint4(rh[4], rh[5], rh[6], rh[7]);
        *reinterpret_cast<uint4*>(kp + TC_W1_BYTES / 2) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
        *reinterpret_cast<uint4*>(kp + TC_W1_BYTES / 2 + 128) = make_uint4(rl[4], rl[5], rl[6], rl[7]);
        const uint32_t ok = __ballot_sync(0xffffffffu, fabsf(chk) < INFINITY);
        if (lane == 0) sh->kok[r >> 5] = ok;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the MMA's async-proxy reads
        warp_arrive(TC3_BAR(key_ready), lane);
        TC3_TE(1);
      };

      // E1 of chunk tile g: h1 = LeakyReLU(D1) (bias already in the GEMM), this quarter's 32 columns of X, in place
      auto e1 = [&](int g) {
        const int G = tile_base + g, xr = G % 3;
        const uint32_t t_x = t_q + 128u * (uint32_t)xr;
        mbar_wait(TC3_BAR(d1_full[0]) + 8u * xr, (uint32_t)((G / 3) & 1));
        tc_fence_after();
        TC3_TE(3);
        float dv[32];
        tmem_ld32(t_x, dv);
        tc_wait_ld();
        uint32_t rh[16], rl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float h0, h1;
          leaky2(dv[2 * i], dv[2 * i + 1], h0, h1);
          split2(h0, h1, rh[i], rl[i]);
        }
        tmem_st16(t_x, rh);
        tmem_st16(t_x + 16, rl);
        tc_wait_st();
        tc_fence_before();
        warp_arrive(TC3_BAR(e1_ready[0]) + 8u * xr, lane);
        TC3_TE(4);
      };

      // E2 of chunk tile g: h2 = LeakyReLU(D * 2^-e2 + b2), this quarter's 32 columns of Y, in place
      auto e2 = [&](int g) {
        const int j = g / ntiles;
        const int G = tile_base + g;
        const float* par = par_of(j);
        const uint32_t t_y = t_q + 384u;
        mbar_wait(TC3_BAR(d2_full[0]) + 8u * (qt >> 1), (uint32_t)(G & 1));   // this half (N = 64) of the hidden GEMM has landed
        tc_fence_after();
        TC3_TE(5);
        const float s2inv = par[TC_PF_SCAL];
        float dv[32];
        tmem_ld32(t_y, dv);
        tc_wait_ld();
        const float* b2 = par + TC_PF_B2 + qt * 32;
        uint32_t rh[16], rl[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = *reinterpret_cast<const float4*>(b2 + i * 4);
          const float2 sc = make_float2(s2inv, s2inv);
          const float2 p01 = __ffma2_rn(make_float2(dv[4 * i + 0], dv[4 * i + 1]), sc, make_float2(bb.x, bb.y));
          const float2 p23 = __ffma2_rn(make_float2(dv[4 * i + 2], dv[4 * i + 3]), sc, make_float2(bb.z, bb.w));
          float h0, h1, h2, h3;
          leaky2(p01.x, p01.y, h0, h1);
          leaky2(p23.x, p23.y, h2, h3);
          split2(h0, h1, rh[2 * i], rl[2 * i]);
          split2(h2, h3, rh[2 * i + 1], rl[2 * i + 1]);
        }
        tmem_st16(t_y, rh);
        tmem_st16(t_y + 16, rl);
        tc_wait_st();
        tc_fence_before();
        warp_arrive(TC3_BAR(e2_ready[0]) + 8u * (qt >> 1), lane);
        TC3_TE(8);
      };

      // E3 of chunk tile g, this quarter's dimension: wait for its pass (quarters 0, 1: the tile's first, quarters 2, 3: its
      // second), pull the 48 logits into registers, hand the columns back, evaluate.  A stage without that dimension still
      // completes the (empty) pass: the barriers' phases stay two per tile.
      auto e3 = [&](int g) {
        const int j = g / ntiles, t = g - j * ntiles;
        const int G = tile_base + g, xr = G % 3;
        const float* par = par_of(j);
        const int* pint = reinterpret_cast<const int*>(par + TC_PI);
        const int L = pint[TC_PI_L];
        const int r = t * TC_TILE + lrow;
        float* srow = state + r;
        const uint32_t bar_off = 16u * (uint32_t)xr + 8u * (uint32_t)(qt >> 1);   // [region][pass]
        mbar_wait(TC3_BAR(d3_full[0][0]) + bar_off, (uint32_t)((G / 3) & 1));
        TC3_TE(9);
        float ldd = 0.0f;
        if (qt >= L) {
          warp_arrive(TC3_BAR(d3_free[0][0]) + bar_off, lane);
        } else {
          tc_fence_after();
          float raw[48];
          const uint32_t t_x = t_s + 128u * (uint32_t)xr;
          tmem_ld32(t_x, raw);
          tmem_ld16(t_x + 32, raw + 32);
          tc_wait_ld();
          tc_fence_before();
          warp_arrive(TC3_BAR(d3_free[0][0]) + bar_off, lane);   // (with the pass's other quarter) the columns may be overwritten
          const int col = pint[TC_PI_LOW + qt];
          const float xin = srow[col * CR];
          float y;
          int idx;
          bool fin;
          spline_dim<INV>(raw, par + TC_PF_B3 + qt * TC_DIM_N, par[TC_PF_SCAL + 1], xin, par[TC_PF_SCAL + 2],
                          par[TC_PF_SCAL + 3], pint[TC_PI_PERIODIC] != 0, y, ldd, idx, fin);
          if (!fin && ((sh->kok[r >> 5] >> (r & 31)) & 1u)) atomicOr(&sh->ovf[r >> 5], 1u << (r & 31));
          srow[col * CR] = y;
          if (args.bins != nullptr && r < nrows)
            args.bins[(row0 + r) * plan->n_spline_dims + pint[TC_PI_BINS] + qt] = min(idx, pint[TC_PI_K]);
        }
        srow[my_pend * CR] = INV ? -ldd : ldd;
        TC3_TE(10);
        if (t == ntiles - 1) free_par(j);   // this warp's last use of the stage's parameter block
      };

      // ---- the coupling stages: one period per tile ----
      if (qt == 0) write_keys(0);
      wait_par(0);
      e1(0);
      for (int g = 0; g < T; ++g) {
        if (g + 1 < T) {
          if (qt == 0) write_keys(g + 1);
          else if ((g + 1) % ntiles == 0) wait_par((g + 1) / ntiles);   // (quarter 0 waits inside write_keys)
        }
        if (qt >= 2 && g > 0) e3(g - 1);
        e2(g);
        if (g + 1 < T) e1(g + 1);
        if (qt < 2) e3(g);
      }
      if (qt >= 2) e3(T - 1);
      tile_base += T;
      seq0 += n_seq;
      named_bar_sync(1, TC_EPI_THREADS);

      // ---- pass Z: the last stage's log-det terms, trailing steps, latent log-density, outputs ----
      uint32_t n_bad = 0u;
      for (int r = etid, it = 0; r < ntiles * TC_TILE; r += TC_EPI_THREADS, ++it) {
        if (r >= nrows) continue;
        const long long row = row0 + r;
        float* st = state + r;
        st[(D + level_z) * CR] += (st[COL_PEND * CR] + st[COL_PEND2 * CR]) + (st[(COL_PEND + 1) * CR] + st[(COL_PEND2 + 1) * CR]);
        int lv = level_z;
        light_steps_cols(plan, INV, last_nsc + 1, ns, lv, st, CR);
        n_bad += ((sh->ovf[r >> 5] >> (r & 31)) & (in_ok >> it) & 1u);
        if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
          const float lat = pz_latent_logprob(plan, [&](int jj) { return st[plan->perm_final[jj] * CR]; });
          const float lpv = pz_nan_to_num_lp(lat + st[D * CR]);
          if (fuse) st[COL_PEND * CR] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
          else args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
        } else {
          for (int jj = 0; jj < D; ++jj) args.out0[row * D + jj] = st[(INV ? jj : plan->perm_final[jj]) * CR];
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
      if (fuse) tc_fused_reduction(lp, args, sl, base, state, CR, COL_PEND, nrows, unit0, sub, (int)tid, warp, lane);
      else named_bar_sync(1, TC_EPI_THREADS);   // pass A of the next chunk rewrites rows other threads are still reading
    }
  } else if (warp == 16) {
    // =============================== the MMA issuer (all tiles, one program order) ===============================
    const uint32_t sb = smem_u32(base);
    const int total_seq = my_chunks * n_seq;
    auto dims_of = [&](int s) -> int {   // (reads the plan in global memory: once per stage, not per tile)
      const int st_i = __shfl_sync(0xffffffffu, (int)sh->order[s % n_seq], 0);
      const int sub_i = __shfl_sync(0xffffffffu, (int)sh->sub[s % n_seq], 0);
      return __shfl_sync(0xffffffffu, tc_sub_dims(plan->nsc[st_i].L, sub_i), 0);
    };
    auto tiles_of = [&](int s) -> int {
      long long row0, unit0;
      int nrows, sub;
      tc_chunk_span(lp, args, s / n_seq, row0, nrows, unit0, sub);
      return max(TC3_MIN_TILES, (nrows + TC_TILE - 1) / TC_TILE);
    };
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0), t_y = tb + 384u;
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = 128
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // N = 64
    auto idesc3 = [](int dims) -> uint32_t {   // an output pass of one or two spline dimensions: N = 48 or 96
      return (1u << 4) | ((uint32_t)((TC_DIM_N * dims) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    };
    const uint32_t w1 = sb + sl.w1, w2 = sb + sl.w2, w3 = sb + sl.w3, keys = sb + sl.keys;

    // first-layer GEMM of tile Gn (stage sn; first / last tile of that stage): keys (shared memory, K = 16) x W1 -> X
    auto issue_g1 = [&](int Gn, int sn, bool first, bool last) {
      const int xr = Gn % 3;
      if (first) mbar_wait(TC_BAR(w1_full), (uint32_t)(sn & 1));
      mbar_wait(TC3_BAR(key_ready), (uint32_t)(Gn & 1));
      if (Gn >= 3)   // the second pass of the region's previous tile is in its threads' registers
        mbar_wait(TC3_BAR(d3_free[0][0]) + 16u * xr + 8u, (uint32_t)(((Gn / 3) - 1) & 1));   // (Gn / 3: a multiply-shift)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_x = tb + 128u * (uint32_t)xr;
        const uint64_t dh = make_b_desc_w1(w1), dl = make_b_desc_w1(w1 + TC_W1_BYTES / 2);
        const uint64_t ah = make_b_desc_w1(keys), al = make_b_desc_w1(keys + TC_W1_BYTES / 2);
        umma_f16_ss(t_x, al, dh, idesc1, false);  // lo * hi
        umma_f16_ss(t_x, ah, dl, idesc1, true);   // hi * lo
        umma_f16_ss(t_x, ah, dh, idesc1, true);   // hi * hi
        umma_commit(TC3_BAR(d1_full[0]) + 8u * xr);
        umma_commit(TC3_BAR(key_free));
        if (last) umma_commit(TC_BAR(w1_free));
      }
      __syncwarp();
    };

    // The issuer shares its scheduler with four busy epilogue warps: every instruction it runs between two GEMMs is a
    // handful of clocks of dry tensor pipe.  So the walk's bookkeeping (which stage / tile / region comes next) is done one
    // tile ahead, while the issuer waits for the first output pass to be drained anyway.
    struct Tile {
      int s, sq, t, nt, L, xr;   // stage (of the CTA's sequence; sq = s mod n_seq), tile of the stage, tiles, spline dims, X region
      uint32_t upar;             // parity of the region's barriers for this tile
      bool valid, first_of_chunk;
    };
    auto successor = [&](const Tile& c) -> Tile {
      Tile n = c;
      n.first_of_chunk = false;
      n.xr = c.xr == 2 ? 0 : c.xr + 1;
      n.upar = n.xr == 0 ? c.upar ^ 1u : c.upar;
      n.t = c.t + 1;
      if (n.t == c.nt) {
        n.t = 0;
        n.s = c.s + 1;
        n.sq = c.sq + 1 == n_seq ? 0 : c.sq + 1;
        n.valid = n.s < total_seq;
        if (n.valid) {
          n.first_of_chunk = n.sq == 0;
          if (n.first_of_chunk) n.nt = tiles_of(n.s);
          n.L = dims_of(n.s);
        }
      }
      return n;
    };

    if (total_seq > 0) {
      Tile cur;
      cur.s = cur.sq = cur.t = 0;
      cur.nt = tiles_of(0);
      cur.L = dims_of(0);
      cur.xr = 0;
      cur.upar = 0u;
      cur.valid = true;
      cur.first_of_chunk = true;
      Tile nxt = successor(cur);
      issue_g1(0, 0, true, false);
      for (int G = 0;; ++G) {
        const int xr = cur.xr, L = cur.L;
        const uint32_t t_x = tb + 128u * (uint32_t)xr;
        const uint32_t gpar = (uint32_t)(G & 1);
        const bool last_tile = cur.t == cur.nt - 1;
        // the next tile's keys are written at the start of this period when the tile belongs to the same chunk; the first
        // tile of the next chunk gets its keys only after this chunk's last spline
        const bool next_here = nxt.valid && !nxt.first_of_chunk;
        // ---- hidden GEMM, two N = 64 halves (E2 of the first runs under the second): X[xr] (h1) x W2 -> Y ----
        if (cur.t == 0) mbar_wait(TC_BAR(w2_full), (uint32_t)(cur.s & 1));
        TC3_TI(0);
        mbar_wait(TC3_BAR(e1_ready[0]) + 8u * xr, cur.upar);
        tc_fence_after();
        TC3_TI(1);
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n0 = h * 64;
            issue_gemm<0, 8, true>(t_y + (uint32_t)n0, t_x, w2 + n0 * 128, w2 + TC_W2_BYTES / 2 + n0 * 128, 128 * 128, idesc2);
            umma_commit(TC3_BAR(d2_full[0]) + 8u * h);
          }
          if (last_tile) umma_commit(TC_BAR(w2_free));
        }
        __syncwarp();
        TC3_TI(2);
        // ---- the next tile's first layer: its keys were written before this tile's E2, so they are there (or about to be).
        // (Issued later, under the drain of the first output pass, it costs 6 %: E1 of the next tile then ends too late.) ----
        if (next_here) issue_g1(G + 1, nxt.s, nxt.t == 0, nxt.t == nxt.nt - 1);
        TC3_TI(3);
        // ---- output layer, first pass (dims 0, 1; N = 96): starts on the 64 K-elements of quarters 0 / 1, takes the rest when
        // quarters 2 / 3 arrive ----
        if (cur.t == 0) mbar_wait(TC_BAR(w3_full), (uint32_t)(cur.s & 1));
        mbar_wait(TC3_BAR(e2_ready[0]), gpar);
        tc_fence_after();
        TC3_TI(4);
        const uint32_t id_a = idesc3(L < 2 ? L : 2), id_b = idesc3(L > 3 ? 2 : 1);
        if (elect_one()) issue_gemm<0, 4, true>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, id_a);
        __syncwarp();
        TC3_TI(5);
        mbar_wait(TC3_BAR(e2_ready[0]) + 8u, gpar);
        tc_fence_after();
        TC3_TI(6);
        if (elect_one()) {
          issue_gemm<4, 8, false>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, id_a);
          umma_commit(TC3_BAR(d3_full[0][0]) + 16u * xr);
          if (L <= 2 && last_tile) umma_commit(TC_BAR(w3_free));
        }
        __syncwarp();
        TC3_TI(7);
        const Tile nxt2 = nxt.valid ? successor(nxt) : nxt;   // (bookkeeping under the drain)
        // ---- second pass (dims 2, 3) once quarters 0 / 1 have the first in registers; without such dimensions an empty pass ----
        mbar_wait(TC3_BAR(d3_free[0][0]) + 16u * xr, cur.upar);
        tc_fence_after();
        TC3_TI(9);
        if (elect_one()) {
          if (L > 2) issue_gemm<0, 8, true>(t_x, t_y, w3 + TC_W3_PASS_BYTES, w3 + TC_W3_PASS_BYTES + TC_W3_PASS_BYTES / 2,
                                            TC_PASS_N * 128, id_b);
          umma_commit(TC3_BAR(d3_full[0][0]) + 16u * xr + 8u);
          if (L > 2 && last_tile) umma_commit(TC_BAR(w3_free));
        }
        __syncwarp();
        TC3_TI(10);
        if (!nxt.valid) break;
        if (!next_here) issue_g1(G + 1, nxt.s, nxt.t == 0, nxt.t == nxt.nt - 1);   // first tile of the next chunk
        TC3_TI(12);
        cur = nxt;
        nxt = nxt2;
      }
    }
  } else if (warp == 18) {
    // =============================== weight streamer (as in nsf_chain_tc_kernel) ===============================
    const uint32_t sb = smem_u32(base);
    const int total_seq = my_chunks * n_seq;
    for (int s = 0; s < total_seq; ++s) {
      const NscDev& st = plan->nsc[__shfl_sync(0xffffffffu, (int)sh->order[s % n_seq], 0)];
      const int sub = __shfl_sync(0xffffffffu, (int)sh->sub[s % n_seq], 0);
      const TcLayout lay = TcLayout::make(tc_sub_dims(st.L, sub));
      const uint8_t* src = static_cast<const uint8_t*>(st.tc[sub].blob);
      {
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

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}
