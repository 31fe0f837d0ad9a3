// SIMT (fp32 FFMA) engine: one fused kernel per bijector chain.
//
// A CTA owns a tile of TM = 64 rows and walks the whole chain with the row
// state resident in shared memory: affine stages, then for every
// NeuralSplineCoupling stage the conditioner MLP (register-tiled FFMA GEMMs,
// weights streamed L2 -> smem with a cp.async double buffer), the softmax /
// softplus knot construction, the bin search, the rational-quadratic spline
// and the log-determinant, and finally the latent log-density.  Nothing but
// the input rows and the result ever touches HBM.
//
// Follows: pzflow/flow.py:397-407 (_log_prob), pzflow/bijectors.py:155-170
// (Chain), 480-520 (NeuralSplineCoupling), pzflow/utils.py:45-48 (MLP),
// 64-227 (RationalQuadraticSpline), pzflow/distributions.py:165-194, 414-441.
#include "pzb_common.cuh"

namespace pzb {

// Activations and logits live in shared memory as [col][TM] with the 16-byte
// row groups XOR-swizzled by the column, so both the GEMM operand loads
// (LDS.128, broadcast) and the GEMM epilogue stores (STS.128) are conflict free.
__device__ __forceinline__ int swz(int col, int row) { return col * TM + (row ^ (col & 28)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Two FMAs sharing the multiplicand a as ONE packed fp32x2 instruction (sm_100 FFMA2): the same roundings as two
// FFMAs at half the issue slots, which is what bounds this register-tiled GEMM (74 % of its instructions are FMAs).
__device__ __forceinline__ void fma_pair(float a, float b0, float b1, float& c0, float& c1) {
  const float2 r = __ffma2_rn(make_float2(a, a), make_float2(b0, b1), make_float2(c0, c1));
  c0 = r.x;
  c1 = r.y;
}

// acc[4][4*NCJ] += A[in_dim][TM] (smem, swizzled) x W[in_dim][NC] (global, chunk-major).
// Thread (ty, tx) owns rows 4ty..4ty+3 and columns 64j + 4tx .. +3, j < NCJ.
template <int NCJ>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, int in_dim,
                                          const float* __restrict__ Wg, float* stage,
                                          float (&acc)[4][4 * NCJ], int tid, int ty, int tx) {
  constexpr int NC = 64 * NCJ;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4 * NCJ; ++c) acc[i][c] = 0.0f;

  const int nk = (in_dim + KC - 1) / KC;
  auto prefetch = [&](int kc) {
    const int rows = min(KC, in_dim - kc * KC);
    const int nvec = rows * (NC / 4);
    const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)kc * KC * NC);
    float4* dst = reinterpret_cast<float4*>(stage + (kc & 1) * (KC * MAX_NC));
    for (int i = tid; i < nvec; i += NTHREADS) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  if (nk > 0) prefetch(0);
  for (int kc = 0; kc < nk; ++kc) {
    if (kc + 1 < nk) {
      prefetch(kc + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Ws = stage + (kc & 1) * (KC * MAX_NC);
    const float* Ak = A + kc * KC * TM;
    const int rows = min(KC, in_dim - kc * KC);
    if (rows == KC) {
#pragma unroll 8
      for (int k = 0; k < KC; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(Ak + k * TM + ((4 * ty) ^ (k & 28)));
#pragma unroll
        for (int j = 0; j < NCJ; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(Ws + k * NC + 64 * j + 4 * tx);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; c += 2) fma_pair(av[i], bv[c], bv[c + 1], acc[i][4 * j + c], acc[i][4 * j + c + 1]);
        }
      }
    } else {
      for (int k = 0; k < rows; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(Ak + k * TM + ((4 * ty) ^ (k & 28)));
#pragma unroll
        for (int j = 0; j < NCJ; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(Ws + k * NC + 64 * j + 4 * tx);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; c += 2) fma_pair(av[i], bv[c], bv[c + 1], acc[i][4 * j + c], acc[i][4 * j + c + 1]);
        }
      }
    }
    __syncthreads();
  }
}

// bias (+ LeakyReLU) epilogue: out[col0 + col][TM] swizzled, STS.128 over the 4 rows.
template <int NCJ, bool LEAKY>
__device__ __forceinline__ void store_tile(float* __restrict__ out, int col0,
                                           const float* __restrict__ bias,
                                           const float (&acc)[4][4 * NCJ], int ty, int tx) {
#pragma unroll
  for (int j = 0; j < NCJ; ++j) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + 64 * j + 4 * tx));
    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = col0 + 64 * j + 4 * tx + c;
      float4 v;
      v.x = acc[0][4 * j + c] + bv[c];
      v.y = acc[1][4 * j + c] + bv[c];
      v.z = acc[2][4 * j + c] + bv[c];
      v.w = acc[3][4 * j + c] + bv[c];
      if (LEAKY) {
        v.x = pz_leaky(v.x);
        v.y = pz_leaky(v.y);
        v.z = pz_leaky(v.z);
        v.w = pz_leaky(v.w);
      }
      *reinterpret_cast<float4*>(out + col * TM + ((4 * ty) ^ (col & 28))) = v;
    }
  }
}

// One (row, spline dimension) pair: softmax knots, bin search, RQS, log-det.
// Lg = logits of the current column chunk, [col][TM] swizzled; c0 = first logit
// column of this dimension.  pzflow/bijectors.py:488-491 + pzflow/utils.py:113-227.
__device__ __forceinline__ void spline_pair(float* Lg, int row, int c0, int K, bool periodic,
                                            float B, bool inverse, float x, float& out, float& ld,
                                            int& idx_out) {
  const float twoB = 2.0f * B;
  const int cw = c0, ch = c0 + K, cd = c0 + 2 * K;
  // softmax numerators in place, denominators in registers
  float mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) {
    mw = fmaxf(mw, Lg[swz(cw + j, row)]);
    mh = fmaxf(mh, Lg[swz(ch + j, row)]);
  }
  float sw = 0.0f, sh = 0.0f;
  for (int j = 0; j < K; ++j) {
    const float ew = expf(Lg[swz(cw + j, row)] - mw);
    const float eh = expf(Lg[swz(ch + j, row)] - mh);
    Lg[swz(cw + j, row)] = ew;
    Lg[swz(ch + j, row)] = eh;
    sw += ew;
    sh += eh;
  }
  const bool oob = (x <= -B) || (x >= B);
  const float mx = oob ? fabsf(x) - B : x;

  // knots: pad(-B + cumsum(2B * softmax), left = -B); bin = (#knots <= mx) - 1
  const int cs_ = inverse ? ch : cw;  // searched axis
  const int co_ = inverse ? cw : ch;  // other axis
  const float ss = inverse ? sh : sw, so = inverse ? sw : sh;
  int cnt = (-B <= mx) ? 1 : 0;
  float cum = 0.0f, ksel = -B;
  for (int j = 0; j < K; ++j) {
    const float v = twoB * (Lg[swz(cs_ + j, row)] / ss);
    Lg[swz(cs_ + j, row)] = v;
    cum += v;
    const float knot = -B + cum;
    if (knot <= mx) {
      ++cnt;
      ksel = knot;
    }
  }
  const int idx = cnt - 1;
  float cum2 = 0.0f, osel = -B;
  for (int j = 0; j < K; ++j) {
    const float v = twoB * (Lg[swz(co_ + j, row)] / so);
    Lg[swz(co_ + j, row)] = v;
    cum2 += v;
    if (j + 1 == idx) osel = -B + cum2;
  }
  idx_out = idx;

  SplineKnotSel s;
  s.idx = idx;
  const bool bad = (idx < 0) || (idx >= K);  // take_along_axis "fill" -> NaN
  const float qnan = __int_as_float(0x7fc00000);
  s.wk = bad ? qnan : Lg[swz(cw + (bad ? 0 : idx), row)];
  s.hk = bad ? qnan : Lg[swz(ch + (bad ? 0 : idx), row)];
  s.xk = inverse ? osel : ksel;
  s.yk = inverse ? ksel : osel;
  if (bad) {
    s.dk = qnan;
    s.dkp1 = qnan;
  } else if (periodic) {
    // dk = pad(D, (1, 0), mode="wrap"): dk[t] = D[(t - 1) mod K]
    s.dk = pz_softplus(Lg[swz(cd + (idx + K - 1) % K, row)]);
    s.dkp1 = pz_softplus(Lg[swz(cd + idx, row)]);
  } else {
    // dk = pad(D, (1, 1), constant 1)
    s.dk = (idx == 0) ? 1.0f : pz_softplus(Lg[swz(cd + idx - 1, row)]);
    s.dkp1 = (idx == K - 1) ? 1.0f : pz_softplus(Lg[swz(cd + idx, row)]);
  }
  pz_rqs_eval(s, x, mx, oob, inverse, periodic, out, ld);
}

struct SmemLayout {
  int xs, cs, ldacc, ldp, act0, act1, stage, total;  // float offsets
};

__host__ __device__ inline SmemLayout make_layout(int D, int C, int Lmax, int act_rows, bool pingpong) {
  SmemLayout s;
  int o = 0;
  s.xs = o;
  o += D * TM;
  s.cs = o;
  o += (C > 0 ? C : 0) * TM;
  s.ldacc = o;
  o += MAX_CHAIN_DEPTH * TM;
  s.ldp = o;
  o += Lmax * TM;
  s.act0 = o;
  o += act_rows * TM;
  s.act1 = o;
  o += pingpong ? act_rows * TM : 0;
  s.stage = o;
  o += 2 * KC * MAX_NC;
  s.total = o;
  return s;
}

__device__ void run_nsc(const NscDev& st, bool inverse, const LaunchArgs& args, long long row0,
                        float* Xs, const float* Cs, float* ldlev, float* ldp, float* act0,
                        float* act1, float* stage, int n_spline_dims) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int ty = (warp >> 1) * 4 + (lane >> 3);
  const int tx = (warp & 1) * 8 + (lane & 7);
  const int U = st.U, C = st.C, L = st.L;

  // key = hstack(upper, conditions)[:, :U + C]  (pzflow/bijectors.py:481-483)
  for (int e = tid; e < (U + C) * TM; e += NTHREADS) {
    const int k = e / TM, r = e % TM;
    act0[swz(k, r)] = (k < U) ? Xs[st.perm[k] * TM + r] : Cs[(k - U) * TM + r];
  }
  __syncthreads();

  float* ain = act0;
  const bool pingpong = (act1 != act0);
  // hidden layers: Dense + LeakyReLU (pzflow/utils.py:45-48)
  for (int li = 0; li + 1 < st.n_dense; ++li) {
    const DenseDev& dl = st.dense[li];
    float* aout = pingpong ? (ain == act0 ? act1 : act0) : ain;
    for (int c = 0; c < dl.n_chunks; ++c) {
      float acc[4][8];
      gemm_tile<2>(ain, dl.in_dim, dl.W + (size_t)c * dl.in_dim * 128, stage, acc, tid, ty, tx);
      store_tile<2, true>(aout, c * 128, dl.b + c * 128, acc, ty, tx);
    }
    __syncthreads();
    ain = aout;
  }

  // output layer, one chunk of whole spline dimensions at a time
  const DenseDev& dl = st.dense[st.n_dense - 1];
  float* Lg = stage;  // logits alias the (idle) weight stage buffers
  for (int c = 0; c < dl.n_chunks; ++c) {
    const float* Wc = dl.W + (size_t)c * dl.in_dim * dl.NC;
    const float* bc = dl.b + c * dl.NC;
    if (dl.NC == 64) {
      float acc[4][4];
      gemm_tile<1>(ain, dl.in_dim, Wc, stage, acc, tid, ty, tx);
      store_tile<1, false>(Lg, 0, bc, acc, ty, tx);
    } else if (dl.NC == 128) {
      float acc[4][8];
      gemm_tile<2>(ain, dl.in_dim, Wc, stage, acc, tid, ty, tx);
      store_tile<2, false>(Lg, 0, bc, acc, ty, tx);
    } else {
      float acc[4][12];
      gemm_tile<3>(ain, dl.in_dim, Wc, stage, acc, tid, ty, tx);
      store_tile<3, false>(Lg, 0, bc, acc, ty, tx);
    }
    __syncthreads();
    const int d0 = c * st.dpc;
    const int nd = min(st.dpc, L - d0);
    for (int p = tid; p < nd * TM; p += NTHREADS) {
      const int dl_ = p / TM, r = p % TM;
      const int d = d0 + dl_;
      const int phys = st.perm[U + d];
      const float x = Xs[phys * TM + r];
      float y, ld;
      int idx;
      spline_pair(Lg, r, dl_ * st.cpd, st.K, st.periodic != 0, st.B, inverse, x, y, ld, idx);
      Xs[phys * TM + r] = y;
      ldp[d * TM + r] = ld;
      if (args.bins != nullptr && row0 + r < args.N)
        args.bins[(row0 + r) * n_spline_dims + st.bins_off + d] = idx;
    }
    __syncthreads();
  }
  // log_det.sum(axis=1) (negated for the inverse), then the Chain's running sum
  for (int r = tid; r < TM; r += NTHREADS) {
    float s = 0.0f;
    for (int d = 0; d < L; ++d) s += ldp[d * TM + r];
    ldlev[r] += inverse ? -s : s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 2)
nsf_chain_simt_kernel(const PlanDev* __restrict__ plan, LaunchArgs args, int Lmax, int act_rows,
                      int pingpong) {
  extern __shared__ __align__(16) float smem[];
  const int D = plan->D, C = plan->C;
  const SmemLayout lay = make_layout(D, C, Lmax, act_rows, pingpong != 0);
  float* Xs = smem + lay.xs;
  float* Cs = smem + lay.cs;
  float* ldacc = smem + lay.ldacc;
  float* ldp = smem + lay.ldp;
  float* act0 = smem + lay.act0;
  float* act1 = pingpong ? smem + lay.act1 : act0;
  float* stage = smem + lay.stage;
  const int tid = threadIdx.x;
  const bool inverse = (args.mode == MODE_INVERSE);
  const long long n_tiles = (args.N + TM - 1) / TM;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    // ---- load the row tile (physical = logical order at the chain entrance; the
    // inverse enters at the chain exit, so it scatters through perm_final) ----
    for (int e = tid; e < D * TM; e += NTHREADS) {
      const int r = e / D, j = e % D;
      const long long row = row0 + r;
      const float v = row < args.N ? pz_input_value(args, pz_row_index(args, row), j, D) : 0.0f;
      const int phys = inverse ? plan->perm_final[j] : j;
      Xs[phys * TM + r] = v;
    }
    for (int e = tid; e < C * TM; e += NTHREADS) {
      const int r = e / C, j = e % C;
      const long long row = row0 + r;
      const float v = row < args.N ? pz_cond_value(args, pz_row_index(args, row), j, C) : 0.0f;
      Cs[j * TM + r] = v;
    }
    for (int e = tid; e < MAX_CHAIN_DEPTH * TM; e += NTHREADS) ldacc[e] = 0.0f;
    __syncthreads();

    // ---- walk the chain (pzflow/bijectors.py:155-170) ----
    int level = 0;
    const int ns = plan->n_steps;
    for (int si = 0; si < ns; ++si) {
      const StepDev step = plan->steps[inverse ? ns - 1 - si : si];
      int kind = step.kind;
      if (inverse && kind == STEP_CHAIN_BEGIN) kind = STEP_CHAIN_END;
      else if (inverse && kind == STEP_CHAIN_END) kind = STEP_CHAIN_BEGIN;
      if (kind == STEP_CHAIN_BEGIN) {
        ++level;
        for (int r = tid; r < TM; r += NTHREADS) ldacc[level * TM + r] = 0.0f;
        __syncthreads();
      } else if (kind == STEP_CHAIN_END) {
        for (int r = tid; r < TM; r += NTHREADS) ldacc[(level - 1) * TM + r] += ldacc[level * TM + r];
        --level;
        __syncthreads();
      } else if (kind == STEP_AFFINE) {
        const AffineDev& af = plan->affine[step.idx];
        if (af.kind == AFF_INV_SOFTPLUS) {
          // data-dependent log-det: one thread walks a row's columns (sum in column order)
          for (int r = tid; r < TM; r += NTHREADS) {
            float lds = 0.0f;
            for (int j = 0; j < D; ++j) {
              float ld;
              Xs[j * TM + r] = pz_affine_elem(af, inverse, j, Xs[j * TM + r], ld);
              lds += ld;
            }
            ldacc[level * TM + r] += lds;
          }
        } else {
          for (int e = tid; e < D * TM; e += NTHREADS) {
            float ld;
            Xs[e] = pz_affine_elem(af, inverse, e / TM, Xs[e], ld);
          }
          const float ldc = inverse ? af.logdet_inv : af.logdet_fwd;
          for (int r = tid; r < TM; r += NTHREADS) ldacc[level * TM + r] += ldc;
        }
        __syncthreads();
      } else if (kind == STEP_COLOR) {
        for (int r = tid; r < TM; r += NTHREADS) pz_color_row(plan->color[step.idx], inverse, Xs + r, TM);
        __syncthreads();
      } else {
        run_nsc(plan->nsc[step.idx], inverse, args, row0, Xs, Cs, ldacc + level * TM, ldp, act0,
                act1, stage, plan->n_spline_dims);
      }
    }

    // ---- finish: latent log-density / outputs ----
    if (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR) {
      for (int r = tid; r < TM; r += NTHREADS) {
        const long long row = row0 + r;
        if (row >= args.N) continue;
        const float lat = pz_latent_logprob(plan, [&](int j) { return Xs[plan->perm_final[j] * TM + r]; });
        float lp = pz_nan_to_num_lp(lat + ldacc[r]);  // pzflow/flow.py:404-406
        args.out0[row] = (args.mode == MODE_POSTERIOR) ? expf(lp) : lp;  // flow.py:720
      }
    } else {
      for (int e = tid; e < D * TM; e += NTHREADS) {
        const int r = e / D, j = e % D;
        const long long row = row0 + r;
        if (row >= args.N) continue;
        const int phys = inverse ? j : plan->perm_final[j];
        args.out0[row * D + j] = Xs[phys * TM + r];
      }
      if (args.out1 != nullptr)
        for (int r = tid; r < TM; r += NTHREADS)
          if (row0 + r < args.N) args.out1[row0 + r] = ldacc[r];
    }
    __syncthreads();
  }
}

// Chains without any NeuralSplineCoupling stage never reach the GEMM code, but
// they run through the same kernel (the affine steps are handled above).

cudaError_t launch_simt(const PlanDev* d_plan, const LaunchArgs& args, int D, int C, int Lmax,
                        int act_rows, bool pingpong, int num_sms, cudaStream_t stream) {
  const SmemLayout lay = make_layout(D, C, Lmax, act_rows, pingpong);
  const size_t smem = (size_t)lay.total * sizeof(float);
  // function attributes are per DEVICE (one host thread may drive plans on several): set on every launch, like launch_tc
  cudaError_t e = cudaFuncSetAttribute(nsf_chain_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nsf_chain_simt_kernel, NTHREADS, smem);
  if (occ < 1) occ = 1;
  const long long n_tiles = (args.N + TM - 1) / TM;
  long long grid = (long long)num_sms * occ;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  nsf_chain_simt_kernel<<<(unsigned)grid, NTHREADS, smem, stream>>>(d_plan, args, Lmax, act_rows,
                                                                     pingpong ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace pzb
