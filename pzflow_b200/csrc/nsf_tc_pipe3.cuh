// nsf_chain_tc3_kernel: the default-family tcgen05 engine of nsf_tc.cu with THREE tiles in flight per SM.
// (Included by nsf_tc.cu inside namespace pzb, after its own kernel: it shares that file's packing, shared-memory
// layout, chunk geometry and the arithmetic of pzb_tc_common.cuh.  Same reference lines as nsf_simt.cu.)
//
// nsf_chain_tc_kernel gives each epilogue group a private X and Y region of TMEM (2 x 256 columns), so two tiles
// are in flight and the tensor pipe idles whenever both groups are in an epilogue phase (pipe ~62 % active).  A tile
// only needs BOTH regions from the start of its hidden GEMM to the end of its last output pass; before (first layer,
// E1) and after (the splines of its last passes) it lives in X alone.  So here TMEM holds three X regions and ONE
// shared Y:
//
//     columns [0,128) [128,256) [256,384): X0 X1 X2 (tile G uses X[G mod 3]);  [384,512): Y
//
// and the tensor pipe runs the tiles' "Y phases" back to back:
//
//     Y phase of tile G      G2(G) -> E2(G) -> G3(G) pass A (dims 0, 1) -> G3(G) pass B (dims 2, 3); G1(G+1) slotted in
//     meanwhile, other group  keys of tile G+1, splines of pass B of tile G-1, E1 of tile G+1
//     own group               E2(G), splines of pass A of tile G
//
// One warp issues every MMA, and a warp cannot issue much faster than one tcgen05.mma per ~20 clk next to 16 busy
// epilogue warps (measured: N = 32 / N = 48 instructions, 16 / 24 clk on the pipe, went out at 19 / 32 clk).  So the
// GEMMs are cut into FEWER, WIDER instructions than in nsf_chain_tc_kernel: the hidden GEMM in two N = 64 halves (32 clk
// each instruction), the output layer in two N = 96 passes of two spline dimensions (48 clk).
//
// Tiles alternate between the two epilogue groups (chunk tile t -> group t & 1, as before), so per tile a group runs
//     phase A (under the other group's Y phase):  E3b(previous own tile: dims half + 2), E0 (keys), E1
//     phase B (its own Y phase):                   E2, E3a (dim half)
// ONE issuer warp orders all MMAs (Y's reuse is then plain program order on the tensor pipe).  The K = 16 key operand of
// the first layer moves to shared memory (8 KB, SWIZZLE_NONE, the layout of the W1 tile): there is no TMEM column left
// for it; tcgen05.mma reads it through a descriptor (.ss form).
//
// mbarrier phases.  Tile G is the (G / 3)-th user of its X region: d1_full / e1_ready of the region complete once per
// user, parity (G / 3) & 1.  A region sees exactly TWO output passes per tile -- a stage with one or two spline
// dimensions commits (and its threads drain) an empty second pass -- so d3_full / d3_free always complete
// twice per user: the first pass of a tile is parity 0, the second parity 1.  Barriers of Y and of the key buffer complete
// once per tile: parity G & 1.  (Every waiter has, through the chain of commits it observed before, already seen the
// previous phase complete, so a parity never aliases an older phase.)

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
      g_tc3_trace_e[((2 * grp + half) * 32 + (int)(G - TC3_T0)) * 16 + (ev)] = (unsigned int)clock64();     \
  } while (0)
#else
#define TC3_TI(ev) \
  do {             \
  } while (0)
#define TC3_TE(ev) \
  do {             \
  } while (0)
#endif

#ifndef PZB_TC3_YIELD
#define PZB_TC3_YIELD 0   // nanoseconds phase-A work sleeps per poll while a warp is inside E2 (0: no yielding)
#endif

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
      mbar_init(TC3_BAR(d3_full[r]), 1);
      mbar_init(TC3_BAR(d3_free[r]), TC_GROUP_THREADS / 32);
      for (int h = 0; h < 2; ++h) mbar_init(TC3_BAR(e1_ready[r][h]), TC_TILE / 32);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(TC3_BAR(d2_full[h]), 1);
      mbar_init(TC3_BAR(e2_ready[h]), TC_TILE / 32);
      mbar_init(TC_BAR(par_full[h]), 1);
      mbar_init(TC_BAR(par_free[h]), TC_EPI_THREADS / 32);
    }
    sh->p3.e2_busy = 0u;
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
    // ======================================= epilogue warps =======================================
    const int grp = (warp >> 3) & 1;
    const int half = (warp >> 2) & 1;
    const int lrow = ((warp & 3) << 5) + lane;
    const int gtid = (int)(tid & (TC_GROUP_THREADS - 1));
    const uint32_t t_row = tbase + ((uint32_t)((warp & 3) << 5) << 16);   // this thread's TMEM lane
    const uint32_t t_lane = t_row + (uint32_t)half * 64u;                  // + 128 * region: this half of X
    const uint32_t t_y = t_lane + 384u;                                                           // this half of Y
    uint8_t* keyb = base + sl.keys;
    int seq = 0;                  // coupling (sub-)stages processed so far by this CTA
    int tile_base = 0;            // tiles issued by the CTA before this chunk (X regions rotate through all of them)

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
      uint32_t in_ok = 0u;
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
      for (int q = gtid; q < my_tiles * (TC_TILE / 32); q += TC_GROUP_THREADS)
        sh->ovf[(2 * (q >> 2) + grp) * (TC_TILE / 32) + (q & 3)] = 0u;
      int level = level_after(plan, INV, 0, first_nsc, 0);
      named_bar_sync(2 + grp, TC_GROUP_THREADS);

      // ---- the coupling stages, software-pipelined over this group's tiles ----
      // the deferred half of the previous own tile: its dims half + 2 (passes 2 / 3), evaluated under the NEXT Y phase
      bool have_prev = false;
      int pv_k = 0, pv_xr = 0, pv_L = 0, pv_seq = 0;
      const float* pv_par = nullptr;

      // E2 sits on the tensor pipe's critical path (G2 -> E2 -> G3 of a tile are serial on the one Y), everything a group
      // does in phase A does not: phase-A work sleeps while some warp is inside E2, leaving it the issue slots.
      volatile uint32_t* e2_busy = &sh->p3.e2_busy;
      auto yield_to_e2 = [&]() {
#if PZB_TC3_YIELD
        while (*e2_busy != 0u) __nanosleep(PZB_TC3_YIELD);
#endif
      };
      // pass `second` (0: dims 0 / 1, 1: dims 2 / 3) of tile k (region xr): wait for it, pull this half's 48 logits
      // (columns [48 half, 48 half + 48) of the region) into registers, hand them back, evaluate dimension d = half + 2 * second.
      // A stage without that dimension still completes the (empty) pass: the barriers' phases stay two per tile.
      auto spline_pass = [&](int second, int L, int k, int xr, const float* par, const int* pint) -> float {
        const int d = half + 2 * second;
        const int r = (2 * k + grp) * TC_TILE + lrow;
        float* srow = state + r;
        mbar_wait(TC3_BAR(d3_full[0]) + 8u * xr, (uint32_t)second);
        if (d >= L) {
          warp_arrive(TC3_BAR(d3_free[0]) + 8u * xr, lane);
          return 0.0f;
        }
        tc_fence_after();
        float raw[48];
        const uint32_t t_x = t_row + 128u * (uint32_t)xr + 48u * (uint32_t)half;   // this half's dimension of the pass
        tmem_ld32(t_x, raw);
        tmem_ld16(t_x + 32, raw + 32);
        tc_wait_ld();
        tc_fence_before();
        warp_arrive(TC3_BAR(d3_free[0]) + 8u * xr, lane);   // (with the other half's arrivals) the pass may be overwritten
        if (second) yield_to_e2();
        const int col = pint[TC_PI_LOW + d];
        const float xin = srow[col * CR];
        float y, ldd;
        int idx;
        bool fin;
        spline_dim<INV>(raw, par + TC_PF_B3 + d * TC_DIM_N, par[TC_PF_SCAL + 1], xin, par[TC_PF_SCAL + 2],
                        par[TC_PF_SCAL + 3], pint[TC_PI_PERIODIC] != 0, y, ldd, idx, fin);
        if (!fin && ((sh->kok[r >> 5] >> (r & 31)) & 1u)) atomicOr(&sh->ovf[r >> 5], 1u << (r & 31));
        srow[col * CR] = y;
        if (args.bins != nullptr && r < nrows)
          args.bins[(row0 + r) * plan->n_spline_dims + pint[TC_PI_BINS] + d] = min(idx, pint[TC_PI_K]);
        return ldd;
      };
      auto finish_prev = [&]() {
        if (!have_prev) return;
        have_prev = false;
        const float ldd = spline_pass(1, pv_L, pv_k, pv_xr, pv_par, reinterpret_cast<const int*>(pv_par + TC_PI));
        float* pend = state + (2 * pv_k + grp) * TC_TILE + lrow + (COL_PEND + half) * CR;
        *pend += INV ? -ldd : ldd;
      };

      int prev_end = first_nsc;
      bool pending = false;
      int pend_level = 0;
      int q_chunk = 0;  // sub-stages of this chunk walked so far
      for (int si = first_nsc; si <= last_nsc; ++si) {
        if (plan->steps[INV ? ns - 1 - si : si].kind != STEP_NSC) continue;
        const bool has_light = si > prev_end;
        const int light0 = prev_end;
        const int level_in = level;
        if (has_light) level = level_after(plan, INV, prev_end, si, level);
        prev_end = si + 1;

        const int n_sub = plan->nsc[plan->steps[INV ? ns - 1 - si : si].idx].n_sub;
        for (int sb_i = 0; sb_i < n_sub; ++sb_i, ++q_chunk) {
          const bool light_now = has_light && sb_i == 0;
          // the last tile of the previous stage still owes its dims half + 2; then that stage's parameter block is free
          if (have_prev) {
            finish_prev();
            warp_arrive(TC_BAR(par_free[0]) + 8u * (pv_seq & 1), lane);
          }
          mbar_wait(TC_BAR(par_full[0]) + 8u * (seq & 1), (uint32_t)((seq >> 1) & 1));
          named_bar_sync(2 + grp, TC_GROUP_THREADS);   // the other half's state writes of the previous stage
          const float* par = reinterpret_cast<const float*>(base + sl.par + (seq & 1) * TC_PAR_BYTES);
          const int* pint = reinterpret_cast<const int*>(par + TC_PI);
          const int L = pint[TC_PI_L];

          if (pending || light_now) {
            for (int q = gtid; q < my_tiles * TC_TILE; q += TC_GROUP_THREADS) {
              float* st = state + (2 * (q >> 7) + grp) * TC_TILE + (q & (TC_TILE - 1));
              if (pending) st[(D + pend_level) * CR] += st[COL_PEND * CR] + st[(COL_PEND + 1) * CR];
              if (light_now) {
                int lv = level_in;
                light_steps_cols(plan, INV, light0, si, lv, st, CR);
              }
            }
            named_bar_sync(2 + grp, TC_GROUP_THREADS);
          }

          for (int k = 0; k < my_tiles; ++k) {
            const int G = tile_base + q_chunk * ntiles + (2 * k + grp);   // this tile's number on the CTA
            const int xr = (int)(G % 3);
            const uint32_t gpar = (uint32_t)(G & 1), upar = (uint32_t)((G / 3) & 1);
            const int r = (2 * k + grp) * TC_TILE + lrow;
            float* srow = state + r;
            const uint32_t t_x = t_lane + 128u * (uint32_t)xr;

            // ---- phase A (the other group's Y phase): keys first -- the tile's first-layer GEMM can then be slotted early
            // into that Y phase --, then what the previous own tile still owes, then E1 ----
            // E0: the <= 15 conditioner inputs of a row and a constant 1 (bias row) as a K = 16 fp16 hi/lo operand in
            // shared memory, K-major SWIZZLE_NONE (the W1 tile's layout: 8 x 8 cores of 128 B, K-adjacent cores 128 B
            // apart, 8-row groups 256 B apart).  One thread per row (half 0).
            TC3_TE(0);
            if (half == 0) {
              if (G >= 1) mbar_wait(TC3_BAR(key_free), (uint32_t)((G - 1) & 1));
              const int in_dim = pint[TC_PI_IN];
              float key[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) key[i] = (i < in_dim) ? srow[pint[TC_PI_KEY + i] * CR] : 0.0f;
              float chk = ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + (key[6] + key[7]));
              uint32_t rh[8], rl[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) split2(key[2 * j], key[2 * j + 1], rh[j], rl[j]);
              if (in_dim <= 8) {
                rh[4] = 0x00003C00u;
                rl[4] = 0u;
#pragma unroll
                for (int j = 5; j < 8; ++j) rh[j] = rl[j] = 0u;
              } else {
#pragma unroll
                for (int i = 0; i < 7; ++i) key[i] = (8 + i < in_dim) ? srow[pint[TC_PI_KEY + 8 + i] * CR] : 0.0f;
                chk += ((key[0] + key[1]) + (key[2] + key[3])) + ((key[4] + key[5]) + key[6]);
                split2(1.0f, key[0], rh[4], rl[4]);
#pragma unroll
                for (int j = 0; j < 3; ++j) split2(key[1 + 2 * j], key[2 + 2 * j], rh[5 + j], rl[5 + j]);
              }
              uint8_t* kp = keyb + (lrow >> 3) * 256 + (lrow & 7) * 16;
              *reinterpret_cast<uint4*>(kp) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
              *reinterpret_cast<uint4*>(kp + 128) = make_uint4(rh[4], rh[5], rh[6], rh[7]);
              *reinterpret_cast<uint4*>(kp + TC_W1_BYTES / 2) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
              *reinterpret_cast<uint4*>(kp + TC_W1_BYTES / 2 + 128) = make_uint4(rl[4], rl[5], rl[6], rl[7]);
              const uint32_t ok = __ballot_sync(0xffffffffu, fabsf(chk) < INFINITY);
              if (lane == 0) sh->kok[r >> 5] = ok;
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the MMA's async-proxy reads
              warp_arrive(TC3_BAR(key_ready), lane);
            }
            TC3_TE(1);
            if (k > 0) finish_prev();
            TC3_TE(2);

            // ---- E1: h1 = LeakyReLU(D1), rewritten in place over this half's 64 columns of X[xr] ----
            mbar_wait(TC3_BAR(d1_full[0]) + 8u * xr, upar);
            TC3_TE(3);
            tc_fence_after();
#pragma unroll 1
            for (int bt = 0; bt < 2; ++bt) {
              yield_to_e2();
              float dv[32];
              tmem_ld32(t_x + bt * 32, dv);
              tc_wait_ld();
              uint32_t rh[16], rl[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float h0, h1;
                leaky2(dv[2 * j], dv[2 * j + 1], h0, h1);
                split2(h0, h1, rh[j], rl[j]);
              }
              tmem_st16(t_x + bt * 32, rh);
              tmem_st16(t_x + bt * 32 + 16, rl);
            }
            tc_wait_st();
            tc_fence_before();
            warp_arrive(TC3_BAR(e1_ready[0][0]) + 16u * xr + 8u * half, lane);
            TC3_TE(4);

            // ---- phase B (this tile's Y phase).  E2: h2 = LeakyReLU(D * 2^-e2 + b2), in place in this half of Y ----
            {
              const float s2inv = par[TC_PF_SCAL];
              mbar_wait(TC3_BAR(d2_full[0]) + 8u * half, gpar);   // this half (N = 64) of the hidden GEMM has landed
              tc_fence_after();
#if PZB_TC3_YIELD
              if (lane == 0) atomicAdd(const_cast<uint32_t*>(e2_busy), 1u);
#endif
              TC3_TE(5);
#pragma unroll 1
              for (int bt = 0; bt < 2; ++bt) {
                float dv[32];
                tmem_ld32(t_y + bt * 32, dv);
                tc_wait_ld();
                const float* b2 = par + TC_PF_B2 + half * 64 + bt * 32;
                uint32_t rh[16], rl[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 bb = *reinterpret_cast<const float4*>(b2 + q * 4);
                  const float2 sc = make_float2(s2inv, s2inv);
                  const float2 p01 = __ffma2_rn(make_float2(dv[4 * q + 0], dv[4 * q + 1]), sc, make_float2(bb.x, bb.y));
                  const float2 p23 = __ffma2_rn(make_float2(dv[4 * q + 2], dv[4 * q + 3]), sc, make_float2(bb.z, bb.w));
                  float h0, h1, h2, h3;
                  leaky2(p01.x, p01.y, h0, h1);
                  leaky2(p23.x, p23.y, h2, h3);
                  split2(h0, h1, rh[2 * q], rl[2 * q]);
                  split2(h2, h3, rh[2 * q + 1], rl[2 * q + 1]);
                }
                tmem_st16(t_y + bt * 32, rh);
                tmem_st16(t_y + bt * 32 + 16, rl);
                TC3_TE(6 + 2 * bt);   // 6, 8
              }
            }
            tc_wait_st();
            tc_fence_before();
            warp_arrive(TC3_BAR(e2_ready[0]) + 8u * half, lane);
#if PZB_TC3_YIELD
            if (lane == 0) atomicSub(const_cast<uint32_t*>(e2_busy), 1u);
#endif
            TC3_TE(9);

            // ---- E3a: this half's first dimension (pass `half`); dims half + 2 wait for the next phase A ----
            const float ldsum = spline_pass(0, L, k, xr, par, pint);
            TC3_TE(10);
            srow[(COL_PEND + half) * CR] = INV ? -ldsum : ldsum;
            have_prev = true;
            pv_k = k;
            pv_xr = xr;
            pv_L = L;
            pv_par = par;
            pv_seq = seq;
          }
          pending = true;
          pend_level = level;
          ++seq;
        }
      }
      if (have_prev) {   // drain: the chunk's last tile of this group
        finish_prev();
        warp_arrive(TC_BAR(par_free[0]) + 8u * (pv_seq & 1), lane);
      }
      tile_base += n_seq * ntiles;
      named_bar_sync(2 + grp, TC_GROUP_THREADS);

      // ---- pass Z: trailing steps, latent log-density, outputs of this group's tiles ----
      uint32_t n_bad = 0u;
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
          if (fuse) st[COL_PEND * CR] = (args.mode == MODE_POSTERIOR) ? expf(lpv) : lpv;
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
      if (fuse) tc_fused_reduction(lp, args, sl, base, state, CR, COL_PEND, nrows, unit0, sub, (int)tid, warp, lane);
    }
  } else if (warp == 16) {
    // =============================== the MMA issuer (all tiles, one program order) ===============================
    const uint32_t sb = smem_u32(base);
    const int total_seq = my_chunks * n_seq;
    auto dims_of = [&](int s) -> int {
      const int st_i = __shfl_sync(0xffffffffu, (int)sh->order[s % n_seq], 0);
      const int sub_i = __shfl_sync(0xffffffffu, (int)sh->sub[s % n_seq], 0);
      return __shfl_sync(0xffffffffu, tc_sub_dims(plan->nsc[st_i].L, sub_i), 0);
    };
    auto tiles_of = [&](int s) -> int {
      long long row0, unit0;
      int nrows, sub;
      tc_chunk_span(lp, args, s / n_seq, row0, nrows, unit0, sub);
      return max(2, (nrows + TC_TILE - 1) / TC_TILE);
    };
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0), t_y = tb + 384u;
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = 128
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // N = 64
    auto idesc3 = [](int dims) -> uint32_t {   // an output pass of one or two spline dimensions: N = 48 or 96
      return (1u << 4) | ((uint32_t)((TC_DIM_N * dims) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    };
    const uint32_t w1 = sb + sl.w1, w2 = sb + sl.w2, w3 = sb + sl.w3, keys = sb + sl.keys;

    // first-layer GEMM of tile Gn (stage sn; first / last tile of that stage): keys (shared memory, K = 16) x W1 -> X
    auto issue_g1 = [&](int Gn, int sn, bool first, bool last, bool keys_seen) {
      const int xr = (int)(Gn % 3);
      if (first) mbar_wait(TC_BAR(w1_full), (uint32_t)(sn & 1));
      if (!keys_seen) mbar_wait(TC3_BAR(key_ready), (uint32_t)(Gn & 1));
      if (Gn >= 3) mbar_wait(TC3_BAR(d3_free[0]) + 8u * xr, 1u);   // the second pass of the region's previous tile is in registers
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

    if (total_seq > 0) {
      int s = 0, t = 0, nt = tiles_of(0);
      int L = dims_of(0);   // (a read of the plan in global memory: once per stage, not per tile)
      issue_g1(0, 0, true, nt == 1, false);
      for (int G = 0;; ++G) {
        const int xr = (int)(G % 3);
        const uint32_t t_x = tb + 128u * (uint32_t)xr;
        const uint32_t gpar = (uint32_t)(G & 1);
        const bool last_tile = t == nt - 1;
        // ---- hidden GEMM: X[xr] (h1) x W2 -> Y ----
        if (t == 0) mbar_wait(TC_BAR(w2_full), (uint32_t)(s & 1));
        TC3_TI(0);
        mbar_wait(TC3_BAR(e1_ready[0][0]) + 16u * xr, (uint32_t)((G / 3) & 1));
        mbar_wait(TC3_BAR(e1_ready[0][0]) + 16u * xr + 8u, (uint32_t)((G / 3) & 1));
        tc_fence_after();
        TC3_TI(1);
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {   // two N = 64 halves: E2 of half 0 runs under the GEMM of half 1
            const int n0 = h * 64;
            issue_gemm<0, 8, true>(t_y + (uint32_t)n0, t_x, w2 + n0 * 128, w2 + TC_W2_BYTES / 2 + n0 * 128, 128 * 128, idesc2);
            umma_commit(TC3_BAR(d2_full[0]) + 8u * h);
          }
          if (last_tile) umma_commit(TC_BAR(w2_free));
        }
        __syncwarp();
        // the next tile: its first-layer GEMM is slotted in as soon as its keys are there
        int s2 = s, t2 = t + 1, nt2 = nt;
        if (t2 == nt) {
          s2 = s + 1;
          t2 = 0;
          nt2 = s2 < total_seq ? tiles_of(s2) : 0;
        }
        const bool has_next = s2 < total_seq;
        bool g1_issued = !has_next;
        auto try_g1 = [&]() {
          if (g1_issued) return;
          if (__shfl_sync(0xffffffffu, (int)mbar_test(TC3_BAR(key_ready), (uint32_t)((G + 1) & 1)), 0)) {
            issue_g1(G + 1, s2, t2 == 0, t2 == nt2 - 1, true);
            g1_issued = true;
          }
        };
        TC3_TI(2);
        try_g1();
        TC3_TI(3);
        // ---- output layer, pass A (dims 0, 1; N = 96): starts on half 0's 64 K-elements of Y, takes half 1's when it arrives ----
        if (t == 0) mbar_wait(TC_BAR(w3_full), (uint32_t)(s & 1));
        mbar_wait(TC3_BAR(e2_ready[0]), gpar);
        tc_fence_after();
        TC3_TI(4);
        const uint32_t id_a = idesc3(L < 2 ? L : 2), id_b = idesc3(L > 3 ? 2 : 1);
        if (elect_one()) issue_gemm<0, 4, true>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, id_a);
        __syncwarp();
        TC3_TI(5);
        try_g1();
        mbar_wait(TC3_BAR(e2_ready[0]) + 8u, gpar);
        tc_fence_after();
        TC3_TI(6);
        if (elect_one()) {
          issue_gemm<4, 8, false>(t_x, t_y, w3, w3 + TC_W3_PASS_BYTES / 2, TC_PASS_N * 128, id_a);
          umma_commit(TC3_BAR(d3_full[0]) + 8u * xr);
          if (L <= 2 && last_tile) umma_commit(TC_BAR(w3_free));
        }
        __syncwarp();
        TC3_TI(7);
        try_g1();
        TC3_TI(8);
        // ---- pass B (dims 2, 3) once both halves have pass A in registers; without such dimensions an empty pass ----
        mbar_wait(TC3_BAR(d3_free[0]) + 8u * xr, 0u);
        tc_fence_after();
        TC3_TI(9);
        if (elect_one()) {
          if (L > 2) issue_gemm<0, 8, true>(t_x, t_y, w3 + TC_W3_PASS_BYTES, w3 + TC_W3_PASS_BYTES + TC_W3_PASS_BYTES / 2,
                                            TC_PASS_N * 128, id_b);
          umma_commit(TC3_BAR(d3_full[0]) + 8u * xr);
          if (L > 2 && last_tile) umma_commit(TC_BAR(w3_free));
        }
        __syncwarp();
        TC3_TI(10);
        TC3_TI(11);
        if (!has_next) break;
        if (!g1_issued) issue_g1(G + 1, s2, t2 == 0, t2 == nt2 - 1, false);   // (blocking: its keys come after this tile's splines)
        TC3_TI(12);
        if (s2 != s) L = dims_of(s2);
        s = s2;
        t = t2;
        nt = nt2;
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
