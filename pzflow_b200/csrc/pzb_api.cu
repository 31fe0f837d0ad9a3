// C-ABI of libpzflow_b200.so (see include/pzflow_b200.h): plan compilation
// (Bijector_Info flattening -> packed device plan) and the launch entry points.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>
#include <sched.h>

#include "../../include/pzflow_b200.h"
#include "pzb_common.cuh"

namespace pzb {
cudaError_t launch_simt(const PlanDev* d_plan, const LaunchArgs& args, int D, int C, int Lmax,
                        int act_rows, bool pingpong, int num_sms, cudaStream_t stream);
bool tc_plan_supported(const PlanDev& hp);
int tc_stage_subs(const NscDev& st);
size_t tc_pack_bytes(const NscDev& st, int sub);
void tc_pack_stage(const PlanDev& hp, const NscDev& st, int sub, const float* const* weights, void* dst);
cudaError_t launch_tc(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms,
                      void* workspace, cudaStream_t stream);
int tc_fuse_mode(const PlanDev& hp, int G, int V, int S, long long n_units, int num_sms);
bool wide_plan_supported(const PlanDev& hp);
size_t wide_pack_bytes(const NscDev& st);
void wide_pack_stage(const PlanDev& hp, const NscDev& st, const float* const* weights, void* dst);
cudaError_t launch_wide(const PlanDev* d_plan, const PlanDev& hp, const LaunchArgs& args, int num_sms,
                        unsigned long long* overflow_rows, cudaStream_t stream);
struct TrainOffsets {
  long long w[MAX_DENSE], b[MAX_DENSE];
};
struct TrainBatch {
  const long long* row_idx;
  const float* Xerr;
  const float* cond_err;
  unsigned long long seed, noise_base;
};
bool train_step_supported(const PlanDev& hp);
size_t train_step_workspace_floats(const PlanDev& hp, long long N, long long n_params);
cudaError_t launch_train_step(const PlanDev* d_plan, const PlanDev& hp, const TrainOffsets* d_offs, const float* params,
                              const float* X, const float* cond, const float* w, long long N, float inv_n, long long n_params,
                              float* workspace, float* flat_grad, float* loss, int num_sms, const TrainBatch& tb,
                              cudaStream_t stream);
cudaError_t launch_spline_train_fwd(const float* logits, const float* x, long long n_elems, int K, float B, int periodic,
                                    float* y, float* ld, cudaStream_t stream);
cudaError_t launch_spline_train_bwd(const float* logits, const float* x, const float* gy, const float* gld,
                                    long long n_elems, int K, float B, int periodic, float* glogits, float* gx,
                                    cudaStream_t stream);
cudaError_t launch_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, int* step, cudaStream_t stream);
}  // namespace pzb

using namespace pzb;

static thread_local std::string g_err = "";
static thread_local bool g_force_simt = false;  // run_host's fp16 guard: re-evaluate this thread's job on the SIMT engine
// pzb_set_error_stream: which stream the NEXT error-sample calls of this host thread draw their eps from, and how many
// units (rows / galaxies) the caller's WHOLE call has (jax's original layout pairs element g with g + n_total / 2)
static thread_local int g_eps_layout = 0;
static thread_local long long g_eps_units = 0;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(PZB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Streams, events and device staging buffers of the host-buffer entry points (pzb_*_host): a
// three-deep ring so chunk i's H2D copy, chunk i-1's kernel and chunk i-2's D2H copy overlap.
struct HostPipe {
  static constexpr int NBUF = 3;
  std::mutex mu;
  bool ready = false;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[NBUF] = {}, ev_k[NBUF] = {}, ev_out[NBUF] = {};
  void* d_in[NBUF] = {};
  void* d_cond[NBUF] = {};
  void* d_out0[NBUF] = {};
  void* d_out1[NBUF] = {};
  void* d_grid = nullptr;
  void* d_ens[NBUF] = {};  // ensemble form: [n_flows, chunk] per-flow log-probabilities of a chunk
  size_t cap_in = 0, cap_cond = 0, cap_out0 = 0, cap_out1 = 0, cap_grid = 0, cap_ens = 0;
  // page-locked host staging for the column-source form (float64 DataFrame columns -> float32 rows)
  void* h_in[NBUF] = {};
  void* h_cond[NBUF] = {};
  size_t cap_hin = 0, cap_hcond = 0;
  // page-locked host staging for results whose destination is PAGEABLE memory (a fresh NumPy array): the D2H copy
  // lands here at PCIe speed and the host cores move it on while the GPU works on the next chunks
  void* h_out0[NBUF] = {};
  void* h_out1[NBUF] = {};
  size_t cap_hout0 = 0, cap_hout1 = 0;
  // the fp16 guard: one device word per ring slot that a tensor-core launch sets when a row overflowed, copied to a
  // page-locked per-chunk array behind the chunk's results
  unsigned int* d_flag = nullptr;
  unsigned int* h_flags = nullptr;
  size_t cap_flags = 0;
};

struct pzb_plan {
  int device = 0;
  int num_sms = 0;
  PlanDev host;  // host copy (device pointers inside)
  PlanDev* d_plan = nullptr;
  void* d_weights = nullptr;
  size_t weight_bytes = 0;
  int Lmax = 1, act_rows = 1;
  bool pingpong = false;
  double flops_per_row = 0.0;
  bool tc_ok = false;
  bool wide_ok = false;
  unsigned long long* d_overflow = nullptr;  // rows whose logits came out non-finite on a tensor-core engine (fp16 range)
  mutable std::atomic<unsigned long long> overflow_seen{0};  // value of *d_overflow the host-buffer guard has already acted on
  TrainOffsets* d_train_offs = nullptr;      // pzb_plan_set_train_layout: where each dense layer sits in the flat parameter buffer
  long long train_params = 0;
  int engine = PZB_ENGINE_AUTO;
  bool latent_ready = true;
  mutable HostPipe pipe;
};

// pzflow/distributions.py: every latent of the path as segments of the plan (host copy; caller uploads)
static int fill_latent(PlanDev& hp, const pzb_latent_desc* segs, int n_segs) {
  if (segs == nullptr || n_segs < 1 || n_segs > MAX_LAT)
    return fail(PZB_ERR_UNSUPPORTED, "latent: between 1 and %d segments are supported", MAX_LAT);
  int start = 0;
  for (int i = 0; i < n_segs; ++i) {
    const pzb_latent_desc& d = segs[i];
    LatentSeg& sg = hp.lat[i];
    memset(&sg, 0, sizeof(sg));
    sg.kind = d.kind;
    sg.start = start;
    sg.n = d.dim;
    sg.B = d.B;
    if (d.dim < 1 || start + d.dim > hp.D) return fail(PZB_ERR_INVALID, "latent: segment %d does not fit input_dim %d", i, hp.D);
    switch (d.kind) {
      case PZB_LATENT_UNIFORM:
        sg.c0 = (float)(-d.dim) * logf(2.0f * d.B);  // -input_dim * log(2B)
        break;
      case PZB_LATENT_CENTBETA13: {
        const float g13 = (float)lgamma(13.0), g26 = (float)lgamma(26.0);
        sg.c0 = (g13 + g13) - g26;
        break;
      }
      case PZB_LATENT_NORMAL:
        break;
      case PZB_LATENT_CENTBETA: {
        if (d.params != nullptr && d.n_params != 2 * d.dim)
          return fail(PZB_ERR_INVALID, "latent: CentBeta segment %d expects %d parameters", i, 2 * d.dim);
        for (int j = 0; j < d.dim; ++j) {
          // a = exp(params[i][0]), b = exp(params[i][1]); default (0, 0) (distributions.py:66, 84-85)
          const float a = d.params ? expf(d.params[2 * j]) : 1.0f, b = d.params ? expf(d.params[2 * j + 1]) : 1.0f;
          hp.latent_a[start + j] = a;
          hp.latent_b[start + j] = b;
          hp.latent_betaln[start + j] =
              ((float)lgamma((double)a) + (float)lgamma((double)b)) - (float)lgamma((double)a + (double)b);
        }
        break;
      }
      case PZB_LATENT_TDIST: {
        if (d.params != nullptr && d.n_params != 1) return fail(PZB_ERR_INVALID, "latent: Tdist expects one parameter (log nu)");
        const float nu = d.params ? expf(d.params[0]) : 30.0f;  // default log(30) (distributions.py:326)
        const float t = 0.5f * (nu + (float)d.dim);
        sg.c0 = ((float)lgamma((double)t) - (float)lgamma(0.5 * (double)nu)) - ((float)d.dim / 2.0f) * logf(nu * 3.14159265358979323846f);
        sg.c1 = t;
        sg.c2 = 1.0f / nu;
        break;
      }
      default:
        return fail(PZB_ERR_UNSUPPORTED, "latent: kind %d cannot be a segment", d.kind);
    }
    start += d.dim;
  }
  if (start != hp.D) return fail(PZB_ERR_INVALID, "latent: segments cover %d of %d dimensions", start, hp.D);
  hp.n_lat = n_segs;
  return PZB_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// float32 rows [n, w] -> w column arrays (rows r0.. of each), on all host cores
static void scatter_rows_to_cols(const float* src, int w, int64_t r0, int64_t n, float* const* cols, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t blk = 0; blk < (n + 4095) / 4096; ++blk) {
    const int64_t a = blk * 4096, b = std::min<int64_t>(n, a + 4096);
    for (int j = 0; j < w; ++j) {
      float* dst = cols[j] + r0;
      for (int64_t r = a; r < b; ++r) dst[r] = src[r * w + j];
    }
  }
}

// float64 / float32 columns -> float32 rows [n, w] on all host cores (explicit thread count: torchrun exports OMP_NUM_THREADS=1)
template <typename T>
static void gather_cast_rows(const void* const* cols, int w, int64_t r0, int64_t n, float* dst, const float* mean,
                             const float* stdv, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t blk = 0; blk < (n + 4095) / 4096; ++blk) {
    const int64_t a = blk * 4096, b = std::min<int64_t>(n, a + 4096);
    for (int j = 0; j < w; ++j) {
      const T* src = static_cast<const T*>(cols[j]) + r0;
      if (mean != nullptr) {
        const float m = mean[j], sd = stdv[j];
        for (int64_t r = a; r < b; ++r) dst[r * w + j] = ((float)src[r] - m) / sd;
      } else {
        for (int64_t r = a; r < b; ++r) dst[r * w + j] = (float)src[r];
      }
    }
  }
}

extern "C" {

int pzb_abi_version(void) { return PZB_ABI_VERSION; }
const char* pzb_last_error(void) { return g_err.c_str(); }
long long pzb_launch_count_ll(void) { return g_launches.load(); }
int64_t pzb_launch_count(void) { return (int64_t)g_launches.load(); }

int pzb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int pzb_plan_create(const pzb_stage_desc* stages, int32_t n_stages, int32_t input_dim,
                    int32_t latent_kind, float latent_B, int32_t device, pzb_plan** out) {
  if (out == nullptr) return fail(PZB_ERR_INVALID, "plan_create: out is NULL");
  *out = nullptr;
  if (n_stages < 0 || (n_stages > 0 && stages == nullptr))
    return fail(PZB_ERR_INVALID, "plan_create: bad stage list");
  const int D = input_dim;
  if (D < 1 || D > MAX_D)
    return fail(PZB_ERR_UNSUPPORTED, "plan_create: input_dim %d outside [1, %d]", D, MAX_D);
  if (latent_kind < PZB_LATENT_UNIFORM || latent_kind > PZB_LATENT_JOINT)
    return fail(PZB_ERR_UNSUPPORTED, "plan_create: latent kind %d is outside the hot path", latent_kind);
  if (pzb_device_count() <= 0)
    return fail(PZB_ERR_NO_DEVICE, "plan_create: no CUDA device (this library has no CPU fallback)");

  pzb_plan* plan = new (std::nothrow) pzb_plan();
  if (!plan) return fail(PZB_ERR_INVALID, "plan_create: out of host memory");
  PlanDev& hp = plan->host;
  memset(&hp, 0, sizeof(hp));
  hp.D = D;
  hp.C = 0;
  hp.latent_kind = latent_kind;
  hp.latent_B = latent_B;
  {
    // one default segment over all D dimensions (a Joint latent gets its segments from pzb_plan_set_latent)
    pzb_latent_desc one{};
    one.kind = latent_kind == PZB_LATENT_JOINT ? PZB_LATENT_UNIFORM : latent_kind;
    one.dim = D;
    one.B = latent_B;
    one.params = nullptr;
    one.n_params = 0;
    int lrc = fill_latent(hp, &one, 1);
    if (lrc != PZB_OK) {
      delete plan;
      return lrc;
    }
    plan->latent_ready = latent_kind != PZB_LATENT_JOINT;
  }

  // ---- pass 1: flatten, fold Roll into permutations, size the weight blob ----
  int cur[MAX_D];
  for (int j = 0; j < D; ++j) cur[j] = j;
  int depth = 0;
  size_t blob = 0;
  struct Pending {
    int nsc_idx;
    const float* const* weights;
  };
  std::vector<Pending> pend;
  int rc = PZB_OK;
  for (int si = 0; si < n_stages && rc == PZB_OK; ++si) {
    const pzb_stage_desc& sd = stages[si];
    if (hp.n_steps >= MAX_STEPS) {
      rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d steps", MAX_STEPS);
      break;
    }
    switch (sd.kind) {
      case 5: {  // chain begin
        if (++depth >= MAX_CHAIN_DEPTH) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: Chain nesting deeper than %d", MAX_CHAIN_DEPTH - 1);
          break;
        }
        hp.steps[hp.n_steps++] = {STEP_CHAIN_BEGIN, 0};
        break;
      }
      case 6: {  // chain end
        if (--depth < 0) {
          rc = fail(PZB_ERR_INVALID, "plan_create: unbalanced chain markers");
          break;
        }
        hp.steps[hp.n_steps++] = {STEP_CHAIN_END, 0};
        break;
      }
      case PZB_STAGE_ROLL: {
        // out[:, j] = in[:, (j - shift) mod D]  (jnp.roll, pzflow/bijectors.py:586)
        int nxt[MAX_D];
        for (int j = 0; j < D; ++j) {
          int src = ((j - sd.shift) % D + D) % D;
          nxt[j] = cur[src];
        }
        memcpy(cur, nxt, sizeof(int) * D);
        break;
      }
      case PZB_STAGE_REVERSE: {
        // out[:, j] = in[:, D - 1 - j]  (pzflow/bijectors.py:541-549)
        int nxt[MAX_D];
        for (int j = 0; j < D; ++j) nxt[j] = cur[D - 1 - j];
        memcpy(cur, nxt, sizeof(int) * D);
        break;
      }
      case PZB_STAGE_SHUFFLE: {
        // out[:, j] = in[:, perm[j]]  (pzflow/bijectors.py:800-804); perm comes from the caller's jax key schedule
        if (sd.vec_a == nullptr || sd.n_vec != D) {
          rc = fail(PZB_ERR_INVALID, "plan_create: Shuffle needs a permutation of %d columns", D);
          break;
        }
        int nxt[MAX_D];
        bool seen[MAX_D] = {false};
        for (int j = 0; j < D && rc == PZB_OK; ++j) {
          const int src = (int)sd.vec_a[j];
          if (src < 0 || src >= D || seen[src]) {
            rc = fail(PZB_ERR_INVALID, "plan_create: Shuffle's column list is not a permutation");
            break;
          }
          seen[src] = true;
          nxt[j] = cur[src];
        }
        if (rc != PZB_OK) break;
        memcpy(cur, nxt, sizeof(int) * D);
        break;
      }
      case PZB_STAGE_DEQUANTIZE: {
        // UniformDequantizer as the reference executes it (pzflow/bijectors.py:892-907): forward identity, inverse floor
        if (hp.n_affine >= MAX_AFFINE) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d elementwise stages", MAX_AFFINE);
          break;
        }
        if (sd.vec_a == nullptr || sd.n_vec < 1) {
          rc = fail(PZB_ERR_INVALID, "plan_create: UniformDequantizer needs column indices");
          break;
        }
        AffineDev& af = hp.affine[hp.n_affine];
        memset(&af, 0, sizeof(af));
        af.kind = AFF_DEQUANTIZE;
        for (int v = 0; v < sd.n_vec; ++v) {
          int c = (int)sd.vec_a[v];
          if (c < 0) c += D;  // numpy-style negative index
          if (c < 0 || c >= D) {
            rc = fail(PZB_ERR_INVALID, "plan_create: UniformDequantizer column %d outside [0, %d)", (int)sd.vec_a[v], D);
            break;
          }
          af.b[cur[c]] = 1.0f;
        }
        if (rc != PZB_OK) break;
        hp.steps[hp.n_steps++] = {STEP_AFFINE, hp.n_affine++};
        break;
      }
      case PZB_STAGE_SCALE:
      case PZB_STAGE_INV_SOFTPLUS: {
        if (hp.n_affine >= MAX_AFFINE) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d elementwise stages", MAX_AFFINE);
          break;
        }
        AffineDev& af = hp.affine[hp.n_affine];
        memset(&af, 0, sizeof(af));
        if (sd.kind == PZB_STAGE_SCALE) {
          af.kind = AFF_SCALE;
          af.B = sd.B;
          // jnp.log(scale ** input_dim) (pzflow/bijectors.py:692-695)
          af.logdet_fwd = logf(powf(sd.B, (float)D));
          af.logdet_inv = -af.logdet_fwd;
        } else {
          af.kind = AFF_INV_SOFTPLUS;
          if (sd.vec_a == nullptr || sd.vec_b == nullptr || sd.n_vec < 1) {
            rc = fail(PZB_ERR_INVALID, "plan_create: InvSoftplus needs column indices and sharpness");
            break;
          }
          for (int v = 0; v < sd.n_vec; ++v) {
            int c = (int)sd.vec_a[v];
            if (c < 0) c += D;  // numpy-style negative index
            if (c < 0 || c >= D) {
              rc = fail(PZB_ERR_INVALID, "plan_create: InvSoftplus column %d outside [0, %d)", (int)sd.vec_a[v], D);
              break;
            }
            af.b[cur[c]] = sd.vec_b[v];
          }
          if (rc != PZB_OK) break;
        }
        hp.steps[hp.n_steps++] = {STEP_AFFINE, hp.n_affine++};
        break;
      }
      case PZB_STAGE_COLOR_TRANSFORM: {
        if (hp.n_color >= MAX_COLOR) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d ColorTransform stages", MAX_COLOR);
          break;
        }
        const int n = sd.n_vec;
        if (sd.vec_a == nullptr || n < 1 || n > D) {
          rc = fail(PZB_ERR_INVALID, "plan_create: ColorTransform needs 1..%d magnitude indices", D);
          break;
        }
        ColorDev& cd = hp.color[hp.n_color];
        memset(&cd, 0, sizeof(cd));
        cd.n_mag = n;
        cd.ref = -1;
        bool is_mag[MAX_D] = {false};
        int mag[MAX_D];
        for (int i = 0; i < n; ++i) {
          mag[i] = (int)sd.vec_a[i];
          if (mag[i] < 0 || mag[i] >= D || is_mag[mag[i]]) {
            rc = fail(PZB_ERR_INVALID, "plan_create: bad ColorTransform magnitude index %d", mag[i]);
            break;
          }
          is_mag[mag[i]] = true;
          if (mag[i] == sd.shift) cd.ref = i;
          cd.slot[i] = (int16_t)cur[mag[i]];
        }
        if (rc != PZB_OK) break;
        if (cd.ref < 0) {
          rc = fail(PZB_ERR_INVALID, "ref_idx must be in mag_idx.");
          break;
        }
        // new logical order: [front columns (ascending), reference magnitude, colours]  (bijectors.py:236-262)
        int nxt[MAX_D], k = 0;
        for (int j = 0; j < D; ++j)
          if (!is_mag[j]) nxt[k++] = cur[j];
        nxt[k++] = cd.slot[n - 1];
        for (int i = 0; i + 1 < n; ++i) nxt[k++] = cd.slot[i];
        memcpy(cur, nxt, sizeof(int) * D);
        hp.steps[hp.n_steps++] = {STEP_COLOR, hp.n_color++};
        break;
      }
      case PZB_STAGE_SHIFT_BOUNDS:
      case PZB_STAGE_STANDARD_SCALER: {
        if (hp.n_affine >= MAX_AFFINE) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d affine stages", MAX_AFFINE);
          break;
        }
        if (sd.vec_a == nullptr || sd.vec_b == nullptr || (sd.n_vec != 1 && sd.n_vec != D)) {
          rc = fail(PZB_ERR_INVALID, "plan_create: affine stage %d needs 1 or %d values", si, D);
          break;
        }
        AffineDev& af = hp.affine[hp.n_affine];
        af.B = sd.B;
        if (sd.kind == PZB_STAGE_SHIFT_BOUNDS) {
          af.kind = AFF_SHIFT_BOUNDS;
          float pf = 1.0f, pi = 1.0f;
          for (int v = 0; v < sd.n_vec; ++v) {
            if (sd.vec_a[v] > sd.vec_b[v]) {
              rc = fail(PZB_ERR_INVALID, "All mins must be less than maxes.");
              break;
            }
            const float hr = (sd.vec_b[v] - sd.vec_a[v]) / 2.0f;
            pf = pf * (sd.B / hr);  // jnp.prod(B / half_range), over the n_vec entries
            pi = pi * (hr / sd.B);
          }
          if (rc != PZB_OK) break;
          af.logdet_fwd = logf(pf);
          af.logdet_inv = logf(pi);
          for (int j = 0; j < D; ++j) {
            const int v = sd.n_vec == 1 ? 0 : j;
            af.a[cur[j]] = (sd.vec_b[v] + sd.vec_a[v]) / 2.0f;
            af.b[cur[j]] = (sd.vec_b[v] - sd.vec_a[v]) / 2.0f;
          }
        } else {
          af.kind = AFF_STANDARD_SCALER;
          if (sd.n_vec != D) {
            rc = fail(PZB_ERR_INVALID, "plan_create: StandardScaler needs %d means/stds", D);
            break;
          }
          float p = 1.0f;
          for (int j = 0; j < D; ++j) {
            p = p * sd.vec_b[j];
            af.a[cur[j]] = sd.vec_a[j];
            af.b[cur[j]] = sd.vec_b[j];
          }
          af.logdet_fwd = logf(1.0f / p);
          af.logdet_inv = logf(p);
        }
        hp.steps[hp.n_steps++] = {STEP_AFFINE, hp.n_affine++};
        break;
      }
      case PZB_STAGE_NSC: {
        if (hp.n_nsc >= MAX_NSC) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: more than %d coupling stages", MAX_NSC);
          break;
        }
        NscDev& st = hp.nsc[hp.n_nsc];
        const int L = sd.transformed_dim < 0 ? D - D / 2 : sd.transformed_dim;
        const int U = D - L;
        if (L < 1 || U < 0) {
          rc = fail(PZB_ERR_INVALID, "plan_create: transformed_dim %d invalid for input_dim %d", L, D);
          break;
        }
        if (sd.K < 1 || sd.hidden_layers < 0 || sd.hidden_layers + 1 > MAX_DENSE ||
            sd.hidden_dim < 1 || sd.n_conditions < 0 || sd.weights == nullptr) {
          rc = fail(PZB_ERR_INVALID, "plan_create: bad NeuralSplineCoupling arguments at stage %d", si);
          break;
        }
        if (U + sd.n_conditions > MAX_KEY) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: conditioner input wider than %d", MAX_KEY);
          break;
        }
        const int cpd = 3 * sd.K - 1 + (sd.periodic ? 1 : 0);
        if (cpd > MAX_NC) {
          rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: K = %d needs more than %d logits per dim", sd.K, MAX_NC);
          break;
        }
        // every stage keys on hstack(upper, conditions)[:, :U + n_conditions] (bijectors.py:481-483):
        // stages may use different prefixes of the same conditions array.
        if (sd.n_conditions > hp.C) hp.C = sd.n_conditions;
        st.U = U;
        st.L = L;
        st.C = sd.n_conditions;
        st.K = sd.K;
        st.cpd = cpd;
        st.dpc = MAX_NC / cpd < L ? MAX_NC / cpd : L;
        st.periodic = sd.periodic ? 1 : 0;
        st.n_dense = sd.hidden_layers + 1;
        st.H = sd.hidden_dim;
        st.Hpad = (int)align_up(sd.hidden_dim, 128);
        st.B = sd.B;
        st.bins_off = hp.n_spline_dims;
        hp.n_spline_dims += L;
        for (int j = 0; j < D; ++j) st.perm[j] = (int16_t)cur[j];
        int in_dim = U + sd.n_conditions;
        for (int li = 0; li < st.n_dense; ++li) {
          DenseDev& dl = st.dense[li];
          const bool last = (li == st.n_dense - 1);
          dl.in_dim = in_dim;
          dl.out_dim = last ? cpd * L : sd.hidden_dim;
          if (last) {
            dl.n_chunks = (L + st.dpc - 1) / st.dpc;
            dl.NC = (int)align_up((size_t)st.dpc * cpd, 64);
          } else {
            dl.n_chunks = st.Hpad / 128;
            dl.NC = 128;
          }
          blob += align_up((size_t)dl.n_chunks * dl.in_dim * dl.NC * sizeof(float), 256);
          blob += align_up((size_t)dl.n_chunks * dl.NC * sizeof(float), 256);
          in_dim = sd.hidden_dim;
          if (!last && dl.n_chunks > 1) plan->pingpong = true;
          if (sd.weights[2 * li] == nullptr || sd.weights[2 * li + 1] == nullptr) {
            rc = fail(PZB_ERR_INVALID, "plan_create: NULL weight pointer at stage %d layer %d", si, li);
            break;
          }
        }
        if (rc != PZB_OK) break;
        if (L > plan->Lmax) plan->Lmax = L;
        int ar = U + sd.n_conditions;
        if (st.n_dense > 1 && st.Hpad > ar) ar = st.Hpad;
        if (ar > plan->act_rows) plan->act_rows = ar;
        plan->flops_per_row += 2.0 * ((double)(U + sd.n_conditions) * (st.n_dense > 1 ? sd.hidden_dim : cpd * L) +
                                      (st.n_dense > 1 ? (double)(sd.hidden_layers - 1) * sd.hidden_dim * sd.hidden_dim +
                                                            (double)sd.hidden_dim * cpd * L
                                                      : 0.0));
        pend.push_back({hp.n_nsc, sd.weights});
        hp.steps[hp.n_steps++] = {STEP_NSC, hp.n_nsc++};
        break;
      }
      default:
        rc = fail(PZB_ERR_UNSUPPORTED, "plan_create: stage kind %d is outside the hot path", sd.kind);
    }
  }
  if (rc == PZB_OK && depth != 0) rc = fail(PZB_ERR_INVALID, "plan_create: unbalanced chain markers");
  if (rc != PZB_OK) {
    delete plan;
    return rc;
  }
  for (int j = 0; j < D; ++j) hp.perm_final[j] = (int16_t)cur[j];

  // tensor-core packing is appended to the same blob
  plan->tc_ok = tc_plan_supported(hp);
  size_t tc_off = align_up(blob, 1024);
  size_t tc_total = 0;
  if (plan->tc_ok) {
    for (int s = 0; s < hp.n_nsc; ++s)
      for (int sub = 0; sub < tc_stage_subs(hp.nsc[s]); ++sub) tc_total += align_up(tc_pack_bytes(hp.nsc[s], sub), 1024);
  }
  plan->wide_ok = wide_plan_supported(hp);
  const size_t wide_off = align_up(tc_off + tc_total, 1024);
  size_t wide_total = 0;
  if (plan->wide_ok)
    for (int s = 0; s < hp.n_nsc; ++s) wide_total += align_up(wide_pack_bytes(hp.nsc[s]), 1024);
  const size_t total = wide_off + wide_total;

  // ---- pass 2: pack on the host, upload once ----
  plan->device = device;
  DeviceGuard guard(device);
  if (!guard.ok) {
    delete plan;
    return fail(PZB_ERR_CUDA, "plan_create: cannot select device %d", device);
  }
  cudaDeviceProp prop;
  cudaError_t ce = cudaGetDeviceProperties(&prop, device);
  if (ce != cudaSuccess) {
    delete plan;
    return cuda_fail(ce, "cudaGetDeviceProperties");
  }
  plan->num_sms = prop.multiProcessorCount;

  std::vector<unsigned char> hblob(total > 0 ? total : 256, 0);
  ce = cudaMalloc(&plan->d_weights, hblob.size());  // (never empty: a chain without weights still gets 256 bytes)
  if (ce != cudaSuccess) {
    delete plan;
    return cuda_fail(ce, "cudaMalloc(weights)");
  }
  plan->weight_bytes = hblob.size();
  size_t off = 0;
  unsigned char* dbase = static_cast<unsigned char*>(plan->d_weights);
  for (const Pending& pe : pend) {
    NscDev& st = hp.nsc[pe.nsc_idx];
    for (int li = 0; li < st.n_dense; ++li) {
      DenseDev& dl = st.dense[li];
      const bool last = (li == st.n_dense - 1);
      const float* W = pe.weights[2 * li];
      const float* b = pe.weights[2 * li + 1];
      float* Wp = reinterpret_cast<float*>(hblob.data() + off);
      dl.W = reinterpret_cast<const float*>(dbase + off);
      off += align_up((size_t)dl.n_chunks * dl.in_dim * dl.NC * sizeof(float), 256);
      float* bp = reinterpret_cast<float*>(hblob.data() + off);
      dl.b = reinterpret_cast<const float*>(dbase + off);
      off += align_up((size_t)dl.n_chunks * dl.NC * sizeof(float), 256);
      const int cols_per_chunk = last ? st.dpc * st.cpd : 128;
      for (int c = 0; c < dl.n_chunks; ++c) {
        for (int n = 0; n < dl.NC; ++n) {
          const int col = c * cols_per_chunk + n;
          const bool valid = (n < cols_per_chunk) && (col < dl.out_dim);
          bp[c * dl.NC + n] = valid ? b[col] : 0.0f;
          for (int k = 0; k < dl.in_dim; ++k)
            Wp[((size_t)c * dl.in_dim + k) * dl.NC + n] = valid ? W[(size_t)k * dl.out_dim + col] : 0.0f;
        }
      }
    }
  }
  if (plan->tc_ok) {
    size_t o = tc_off;
    for (const Pending& pe : pend) {
      NscDev& st = hp.nsc[pe.nsc_idx];
      st.n_sub = tc_stage_subs(st);
      for (int sub = 0; sub < st.n_sub; ++sub) {
        const size_t nb = tc_pack_bytes(st, sub);
        tc_pack_stage(hp, st, sub, pe.weights, hblob.data() + o);
        st.tc[sub].blob = dbase + o;
        st.tc[sub].blob_bytes = (int)nb;
        o += align_up(nb, 1024);
      }
    }
  }
  if (plan->wide_ok) {
    size_t o = wide_off;
    for (const Pending& pe : pend) {
      NscDev& st = hp.nsc[pe.nsc_idx];
      const size_t nb = wide_pack_bytes(st);
      wide_pack_stage(hp, st, pe.weights, hblob.data() + o);
      st.wide_blob = dbase + o;
      st.wide_bytes = (int)nb;
      o += align_up(nb, 1024);
    }
  }
  ce = cudaMemcpy(plan->d_weights, hblob.data(), hblob.size(), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&plan->d_overflow), sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(plan->d_overflow, 0, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc(&plan->d_plan, sizeof(PlanDev));
  if (ce == cudaSuccess) ce = cudaMemcpy(plan->d_plan, &hp, sizeof(PlanDev), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    int r = cuda_fail(ce, "plan upload");
    pzb_plan_destroy(plan);
    return r;
  }
  *out = plan;
  return PZB_OK;
}

int pzb_plan_set_latent(pzb_plan* plan, const pzb_latent_desc* segs, int32_t n_segs) {
  if (!plan) return fail(PZB_ERR_INVALID, "set_latent: plan is NULL");
  PlanDev& hp = plan->host;
  int rc = fill_latent(hp, segs, n_segs);
  if (rc != PZB_OK) return rc;
  DeviceGuard guard(plan->device);
  cudaError_t ce = cudaMemcpy(plan->d_plan, &hp, sizeof(PlanDev), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return cuda_fail(ce, "set_latent upload");
  plan->latent_ready = true;
  return PZB_OK;
}

void pzb_plan_destroy(pzb_plan* plan) {
  if (!plan) return;
  DeviceGuard guard(plan->device);
  if (plan->d_plan) cudaFree(plan->d_plan);
  if (plan->d_weights) cudaFree(plan->d_weights);
  if (plan->d_overflow) cudaFree(plan->d_overflow);
  if (plan->d_train_offs) cudaFree(plan->d_train_offs);
  HostPipe& hp = plan->pipe;
  if (hp.ready) {
    cudaStreamSynchronize(hp.s_in);
    cudaStreamSynchronize(hp.s_k);
    cudaStreamSynchronize(hp.s_out);
    for (int b = 0; b < HostPipe::NBUF; ++b) {
      cudaFree(hp.d_in[b]);
      cudaFree(hp.d_cond[b]);
      cudaFree(hp.d_out0[b]);
      cudaFree(hp.d_out1[b]);
      cudaFree(hp.d_ens[b]);
      cudaEventDestroy(hp.ev_in[b]);
      cudaEventDestroy(hp.ev_k[b]);
      cudaEventDestroy(hp.ev_out[b]);
    }
    cudaFree(hp.d_grid);
    if (hp.d_flag) cudaFree(hp.d_flag);
    if (hp.h_flags) cudaFreeHost(hp.h_flags);
    for (int b = 0; b < HostPipe::NBUF; ++b) {
      if (hp.h_in[b]) cudaFreeHost(hp.h_in[b]);
      if (hp.h_cond[b]) cudaFreeHost(hp.h_cond[b]);
      if (hp.h_out0[b]) cudaFreeHost(hp.h_out0[b]);
      if (hp.h_out1[b]) cudaFreeHost(hp.h_out1[b]);
    }
    cudaStreamDestroy(hp.s_in);
    cudaStreamDestroy(hp.s_k);
    cudaStreamDestroy(hp.s_out);
  }
  delete plan;
}

int pzb_plan_input_dim(const pzb_plan* plan) { return plan ? plan->host.D : -1; }
int pzb_plan_n_conditions(const pzb_plan* plan) { return plan ? plan->host.C : -1; }
int pzb_plan_n_spline_dims(const pzb_plan* plan) { return plan ? plan->host.n_spline_dims : -1; }
double pzb_plan_flops_per_row(const pzb_plan* plan) { return plan ? plan->flops_per_row : 0.0; }
int pzb_plan_tc_supported(const pzb_plan* plan) { return plan && plan->tc_ok ? 1 : 0; }
int pzb_plan_wide_supported(const pzb_plan* plan) { return plan && plan->wide_ok ? 1 : 0; }

int pzb_plan_overflow_rows(const pzb_plan* plan, int32_t reset, int64_t* count) {
  if (!plan || !count) return fail(PZB_ERR_INVALID, "overflow_rows: NULL argument");
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(PZB_ERR_CUDA, "cannot select device %d", plan->device);
  unsigned long long v = 0;
  cudaError_t ce = cudaMemcpy(&v, plan->d_overflow, sizeof(v), cudaMemcpyDeviceToHost);  // (synchronises the device)
  if (ce == cudaSuccess && reset) {
    ce = cudaMemset(plan->d_overflow, 0, sizeof(v));
    plan->overflow_seen.store(0);
  }
  if (ce != cudaSuccess) return cuda_fail(ce, "overflow_rows");
  *count = (int64_t)v;
  return PZB_OK;
}

int pzb_plan_set_engine(pzb_plan* plan, int32_t engine) {
  if (!plan) return fail(PZB_ERR_INVALID, "set_engine: plan is NULL");
  if (engine != PZB_ENGINE_AUTO && engine != PZB_ENGINE_SIMT && engine != PZB_ENGINE_TC && engine != PZB_ENGINE_WIDE)
    return fail(PZB_ERR_INVALID, "set_engine: unknown engine %d", engine);
  if (engine == PZB_ENGINE_TC && !plan->tc_ok)
    return fail(PZB_ERR_UNSUPPORTED, "set_engine: the tensor-core engine does not support this plan's shapes");
  if (engine == PZB_ENGINE_WIDE && !plan->wide_ok)
    return fail(PZB_ERR_UNSUPPORTED, "set_engine: the wide tensor-core engine does not support this plan's shapes");
  plan->engine = engine;
  return PZB_OK;
}

int pzb_plan_engine(const pzb_plan* plan) {
  if (!plan) return -1;
  if (g_force_simt) return PZB_ENGINE_SIMT;
  if (plan->engine == PZB_ENGINE_AUTO)
    return plan->tc_ok ? PZB_ENGINE_TC : (plan->wide_ok ? PZB_ENGINE_WIDE : PZB_ENGINE_SIMT);
  return plan->engine;
}

static int launch_chain(const pzb_plan* plan, const LaunchArgs& args, void* stream) {
  if (args.N == 0) return PZB_OK;
  if (!plan->latent_ready && (args.mode == MODE_LOGPROB || args.mode == MODE_POSTERIOR))
    return fail(PZB_ERR_INVALID, "this plan has a Joint latent: call pzb_plan_set_latent before evaluating densities");
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(PZB_ERR_CUDA, "cannot select device %d", plan->device);
  cudaError_t ce;
  const int engine = pzb_plan_engine(plan);
  if (engine == PZB_ENGINE_TC) {
    ce = launch_tc(plan->d_plan, plan->host, args, plan->num_sms, plan->d_overflow, (cudaStream_t)stream);
  } else if (engine == PZB_ENGINE_WIDE) {
    ce = launch_wide(plan->d_plan, plan->host, args, plan->num_sms, plan->d_overflow, (cudaStream_t)stream);
  } else {
    ce = launch_simt(plan->d_plan, args, plan->host.D, plan->host.C, plan->Lmax, plan->act_rows,
                     plan->pingpong, plan->num_sms, (cudaStream_t)stream);
  }
  if (ce != cudaSuccess) return cuda_fail(ce, "chain kernel launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// Posterior launches whose reduction (mean over error samples, sum over variants, trapezoid normalisation, NaN -> 0)
// can happen inside the tensor-core kernel: the pdfs then cross HBM once.  PZB_NO_FUSE=1 forces the scratch route.
static bool posterior_fuses(const pzb_plan* plan, int G, int V, int S, int64_t Ng) {
  const char* env = getenv("PZB_NO_FUSE");
  if ((env != nullptr && atoi(env) != 0) || pzb_plan_engine(plan) != PZB_ENGINE_TC) return false;
  if (V <= 0 && S <= 1) {
    // A plain posterior has nothing to reduce but the trapezoid: fused, a galaxy of G = 1000 points fills 7.8 row
    // tiles (24 idle rows) and both epilogue groups meet once per galaxy -- measured 6 % slower than the chain kernel
    // followed by normalize_pdfs_kernel, whose two extra passes over the pdfs cost 0.1 % at 6.5 TB/s.  So the
    // single-pass form is opt-in (PZB_FUSE_POSTERIOR=1) and bench.py reports it beside the default.
    const char* on = getenv("PZB_FUSE_POSTERIOR");
    if (on == nullptr || atoi(on) == 0) return false;
  }
  return tc_fuse_mode(plan->host, G, V, S, Ng, plan->num_sms) != 0;
}

static int check_common(const pzb_plan* plan, const void* in, const float* cond, int64_t N,
                        const void* outp, const char* who) {
  if (!plan) return fail(PZB_ERR_INVALID, "%s: plan is NULL", who);
  if (N < 0) return fail(PZB_ERR_INVALID, "%s: negative row count", who);
  if (N > 0 && (in == nullptr || outp == nullptr)) return fail(PZB_ERR_INVALID, "%s: NULL buffer", who);
  if (N > 0 && plan->host.C > 0 && cond == nullptr)
    return fail(PZB_ERR_INVALID, "%s: this plan is conditional (C = %d) but conditions are NULL", who, plan->host.C);
  return PZB_OK;
}

int pzb_log_prob(const pzb_plan* plan, const float* X, const float* cond, int64_t N, float* out,
                 int32_t* bins, void* stream) {
  int rc = check_common(plan, X, cond, N, out, "log_prob");
  if (rc) return rc;
  LaunchArgs a{};
  a.mode = MODE_LOGPROB;
  a.X = X;
  a.cond = cond;
  a.N = N;
  a.out0 = out;
  a.bins = bins;
  return launch_chain(plan, a, stream);
}

int pzb_forward(const pzb_plan* plan, const float* X, const float* cond, int64_t N, float* U,
                float* log_det, int32_t* bins, void* stream) {
  int rc = check_common(plan, X, cond, N, U, "forward");
  if (rc) return rc;
  LaunchArgs a{};
  a.mode = MODE_FORWARD;
  a.X = X;
  a.cond = cond;
  a.N = N;
  a.out0 = U;
  a.out1 = log_det;
  a.bins = bins;
  return launch_chain(plan, a, stream);
}

int pzb_inverse(const pzb_plan* plan, const float* U, const float* cond, int64_t N, float* X,
                float* log_det, int32_t* bins, void* stream) {
  int rc = check_common(plan, U, cond, N, X, "inverse");
  if (rc) return rc;
  LaunchArgs a{};
  a.mode = MODE_INVERSE;
  a.X = U;
  a.cond = cond;
  a.N = N;
  a.out0 = X;
  a.out1 = log_det;
  a.bins = bins;
  return launch_chain(plan, a, stream);
}

// ---- small memory-bound helper kernels ----

// SM count of the CURRENT device (the helper kernels below are not tied to a plan), cached per device
static int current_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ float nan_to_num0(float v) {
  if (isnan(v)) return 0.0f;
  if (v == INFINITY) return 3.4028234663852886e38f;
  if (v == -INFINITY) return -3.4028234663852886e38f;
  return v;
}

// pdfs / trapezoid(pdfs, grid) per row, then nan_to_num(nan=0)  (pzflow/flow.py:732-737).
// Optionally averages n_flows stacked inputs first (pzflow/flowEnsemble.py:346-351).
__global__ void __launch_bounds__(256)
normalize_pdfs_kernel(const float* in, int n_flows, long long Ng, int G,
                      const float* __restrict__ grid, int normalize, int nan_to_zero,
                      float* out) {  // in == out for the in-place form: neither may be __restrict__
  __shared__ float red[8];
  for (long long g = blockIdx.x; g < Ng; g += gridDim.x) {
    float part = 0.0f;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
      float y;
      if (n_flows > 1) {
        float s = 0.0f;
        for (int f = 0; f < n_flows; ++f) s += in[((size_t)f * Ng + g) * G + i];
        y = s / (float)n_flows;
        out[(size_t)g * G + i] = y;
      } else {
        y = in[(size_t)g * G + i];
      }
      if (normalize && i + 1 < G) {
        float y1;
        if (n_flows > 1) {
          float s = 0.0f;
          for (int f = 0; f < n_flows; ++f) s += in[((size_t)f * Ng + g) * G + i + 1];
          y1 = s / (float)n_flows;
        } else {
          y1 = in[(size_t)g * G + i + 1];
        }
        part += (grid[i + 1] - grid[i]) * (y1 + y);
      }
    }
    float integral = 1.0f;
    if (normalize) integral = 0.5f * block_sum(part, red);
    __syncthreads();
    if (normalize || nan_to_zero || n_flows == 1) {
      for (int i = threadIdx.x; i < G; i += blockDim.x) {
        float y = (n_flows > 1) ? out[(size_t)g * G + i] : in[(size_t)g * G + i];
        if (normalize) y = y / integral;
        if (nan_to_zero) y = nan_to_num0(y);
        out[(size_t)g * G + i] = y;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
logmeanexp_kernel(const float* __restrict__ lp, int n_flows, long long N, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int f = 0; f < n_flows; ++f) s += expf(lp[(size_t)f * N + i]);
    out[i] = logf(s / (float)n_flows);  // naive log(mean(exp)), pzflow/flowEnsemble.py:243
  }
}

__global__ void __launch_bounds__(256)
sample_uniform_kernel(float* __restrict__ U, long long n, float B, uint64_t seed, long long q0) {
  // q0: index of U's first element in the caller's whole array, over 4 (a sharded call draws what one big call would)
  const long long nq = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq;
       q += (long long)gridDim.x * blockDim.x) {
    const unsigned long long qq = (unsigned long long)(q + q0);
    uint32_t c[4] = {(uint32_t)qq, (uint32_t)(qq >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      pz_philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const long long i = q * 4 + t;
      if (i >= n) break;
      // 23 random mantissa bits -> [1, 2) - 1 -> [0, 1); scale to [-B, B) like jax.random.uniform
      const float u01 = __uint_as_float((c[t] >> 9) | 0x3f800000u) - 1.0f;
      U[i] = fmaxf(-B, u01 * (2.0f * B) + (-B));
    }
  }
}

// jax.random.uniform(key, (total_rows, D), minval, maxval) -- elements [e0, e0 + n) of its n_total, in jax's stream
// layout (pzflow/distributions.py:464-469).  Original layout: the count vector iota(n_total) is cut into two halves
// that ride through the block function together (element g < h pairs with g + h, a zero pads an odd count); the
// partitionable layout gives every element its own block (hi32(g), lo32(g)) and xors the two output words.
__global__ void __launch_bounds__(256)
sample_uniform_threefry_kernel(float* __restrict__ U, long long n, long long e0, long long n_total, uint32_t k0, uint32_t k1,
                               float lo, float hi, int partitionable) {
  const long long h = (n_total + 1) / 2;
  const float span = __fsub_rn(hi, lo);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long g = e + e0;
    uint32_t x0, x1, bits;
    if (partitionable) {
      x0 = (uint32_t)((unsigned long long)g >> 32);
      x1 = (uint32_t)g;
      pz_threefry2x32(k0, k1, x0, x1);
      bits = x0 ^ x1;
    } else if (g < h) {
      x0 = (uint32_t)g;
      x1 = (g + h < n_total) ? (uint32_t)(g + h) : 0u;
      pz_threefry2x32(k0, k1, x0, x1);
      bits = x0;
    } else {
      x0 = (uint32_t)(g - h);
      x1 = (uint32_t)g;
      pz_threefry2x32(k0, k1, x0, x1);
      bits = x1;
    }
    // 23 random mantissa bits -> [1, 2) - 1 -> [0, 1); floats * (maxval - minval) + minval, clamped below (jax's order, no FMA)
    const float u01 = __fsub_rn(__uint_as_float((bits >> 9) | 0x3f800000u), 1.0f);
    U[e] = fmaxf(lo, __fadd_rn(__fmul_rn(u01, span), lo));
  }
}

int pzb_normalize_pdfs(float* pdfs, int64_t Ng, int32_t G, const float* grid, int32_t normalize,
                       int32_t nan_to_zero, void* stream) {
  if (Ng < 0 || G < 1) return fail(PZB_ERR_INVALID, "normalize_pdfs: bad shape");
  if (Ng == 0) return PZB_OK;
  if (!pdfs || !grid) return fail(PZB_ERR_INVALID, "normalize_pdfs: NULL buffer");
  if (!normalize && !nan_to_zero) return PZB_OK;
  const long long cap = (long long)current_num_sms() * 64;
  const unsigned blocks = (unsigned)(Ng < cap ? Ng : cap);
  normalize_pdfs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pdfs, 1, Ng, G, grid, normalize,
                                                                  nan_to_zero, pdfs);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "normalize_pdfs launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_posterior(const pzb_plan* plan, const float* rest, const float* cond, int64_t Ng,
                  int32_t col_idx, const float* grid, int32_t G, int32_t normalize,
                  int32_t nan_to_zero, float* pdfs, void* stream) {
  int rc = check_common(plan, plan && plan->host.D > 1 ? (const void*)rest : (const void*)grid, cond, Ng, pdfs, "posterior");
  if (rc) return rc;
  if (G < 1 || grid == nullptr) return fail(PZB_ERR_INVALID, "posterior: empty grid");
  if (col_idx < 0 || col_idx >= plan->host.D) return fail(PZB_ERR_INVALID, "posterior: column index %d out of range", col_idx);
  LaunchArgs a{};
  a.mode = MODE_POSTERIOR;
  a.X = rest;
  a.cond = cond;
  a.grid = grid;
  a.G = G;
  a.col_idx = col_idx;
  a.N = Ng * (int64_t)G;
  a.out0 = pdfs;
  if (posterior_fuses(plan, G, 0, 0, Ng)) {
    a.fuse = 1;
    a.normalize = normalize;
    a.nan_to_zero = nan_to_zero;
    return launch_chain(plan, a, stream);
  }
  rc = launch_chain(plan, a, stream);
  if (rc) return rc;
  DeviceGuard guard(plan->device);
  return pzb_normalize_pdfs(pdfs, Ng, G, grid, normalize, nan_to_zero, stream);
}

int pzb_ensemble_logmeanexp(const float* lp, int32_t n_flows, int64_t N, float* out, void* stream) {
  if (n_flows < 1 || N < 0) return fail(PZB_ERR_INVALID, "ensemble_logmeanexp: bad shape");
  if (N == 0) return PZB_OK;
  if (!lp || !out) return fail(PZB_ERR_INVALID, "ensemble_logmeanexp: NULL buffer");
  long long blocks = (N + 255) / 256;
  if (blocks > (long long)current_num_sms() * 32) blocks = (long long)current_num_sms() * 32;
  logmeanexp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(lp, n_flows, N, out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "logmeanexp launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_ensemble_posterior_mean(const float* pdfs, int32_t n_flows, int64_t Ng, int32_t G,
                                const float* grid, int32_t normalize, float* out, void* stream) {
  if (n_flows < 1 || Ng < 0 || G < 1) return fail(PZB_ERR_INVALID, "ensemble_posterior_mean: bad shape");
  if (Ng == 0) return PZB_OK;
  if (!pdfs || !out || !grid) return fail(PZB_ERR_INVALID, "ensemble_posterior_mean: NULL buffer");
  const long long cap = (long long)current_num_sms() * 64;
  const unsigned blocks = (unsigned)(Ng < cap ? Ng : cap);
  // n_flows == 1 degenerates to a copy (+ normalisation)
  normalize_pdfs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pdfs, n_flows, Ng, G, grid, normalize, 0, out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "ensemble_posterior_mean launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_sample_uniform(float* U, int64_t N, int32_t D, float B, uint64_t seed, int64_t row_offset, void* stream) {
  if (N < 0 || D < 1 || row_offset < 0) return fail(PZB_ERR_INVALID, "sample_uniform: bad shape");
  if ((row_offset * (int64_t)D) % 4 != 0)
    return fail(PZB_ERR_INVALID, "sample_uniform: row_offset * D must be a multiple of 4 (draws come in blocks of four)");
  if (N == 0) return PZB_OK;
  if (!U) return fail(PZB_ERR_INVALID, "sample_uniform: NULL buffer");
  const long long n = N * (long long)D;
  long long blocks = ((n + 3) / 4 + 255) / 256;
  if (blocks > (long long)current_num_sms() * 32) blocks = (long long)current_num_sms() * 32;
  sample_uniform_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(U, n, B, seed, row_offset * (long long)D / 4);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_uniform launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// ---------------------------------------------------------------------------------------------
// Latent draws other than Uniform (pzflow/distributions.py:101-135 CentBeta, 196-229 CentBeta13, 277-303 Normal,
// 360-392 Tdist): counter-based Philox streams, one per element, so a draw depends on (seed, element index) only.
// Not jax's threefry: like the uniform sampler, sample parity is defined on the map u -> x.
// ---------------------------------------------------------------------------------------------
static unsigned grid_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > (long long)current_num_sms() * 32) b = (long long)current_num_sms() * 32;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int pzb_sample_uniform_threefry(float* U, int64_t N, int32_t D, float minval, float maxval, uint32_t key0, uint32_t key1,
                                int64_t total_rows, int64_t row_offset, int32_t partitionable, void* stream) {
  if (N < 0 || D < 1 || row_offset < 0 || total_rows < row_offset + N)
    return fail(PZB_ERR_INVALID, "sample_uniform_threefry: bad shape");
  if (N == 0) return PZB_OK;
  if (!U) return fail(PZB_ERR_INVALID, "sample_uniform_threefry: NULL buffer");
  const long long n_total = total_rows * (long long)D;
  // jax cuts original-layout requests beyond 2^32 - 2 words into separately keyed blocks: not reproduced
  if (!partitionable && n_total >= 0xFFFFFFFFLL)
    return fail(PZB_ERR_UNSUPPORTED, "sample_uniform_threefry: %lld elements exceed one threefry call of jax's original layout",
                n_total);
  const long long n = N * (long long)D;
  sample_uniform_threefry_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(U, n, row_offset * (long long)D, n_total, key0,
                                                                               key1, minval, maxval, partitionable ? 1 : 0);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_uniform_threefry launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

struct BetaShapes {
  float a[MAX_D], b[MAX_D];
};

// uniform in (0, 1) from 24 random bits
__device__ __forceinline__ float pz_u01(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-8f; }

// Gamma(shape, 1), Marsaglia-Tsang squeeze (shape >= 1; smaller shapes through Gamma(shape + 1) * U^(1/shape)).
// Attempt t of element `elem` on stream `sid` takes the Philox block (elem, t, sid): two words -> one normal (Box-Muller),
// one word -> the acceptance uniform, one word -> the boost uniform of shape < 1.
__device__ float pz_gamma_draw(float shape, unsigned long long seed, unsigned long long elem, uint32_t sid) {
  const bool small = shape < 1.0f;
  const float a = small ? shape + 1.0f : shape;
  const float d = a - 0.33333334f, c = rsqrtf(9.0f * d);
  float boost = 1.0f, out = d;
  for (uint32_t t = 0; t < 64u; ++t) {
    uint32_t w[4];
    pz_philox4(seed, elem, t, sid, w);
    if (t == 0u && small) boost = powf(pz_u01(w[3]), 1.0f / shape);
    const float r = sqrtf(-2.0f * logf(pz_u01(w[0])));
    float sn, cs;
    sincospif(2.0f * pz_u01(w[1]), &sn, &cs);
    const float x = r * cs;
    float v = 1.0f + c * x;
    if (v <= 0.0f) continue;
    v = v * v * v;
    const float u = pz_u01(w[2]);
    const float x2 = x * x;
    if (u < 1.0f - 0.0331f * x2 * x2 || logf(u) < 0.5f * x2 + d * (1.0f - v + logf(v))) {
      out = d * v;
      break;
    }
  }
  return out * boost;
}

__global__ void __launch_bounds__(256)
sample_centbeta_kernel(float* __restrict__ U, long long n, int D, BetaShapes sh, float B, unsigned long long seed,
                       long long e0) {
  // 2B (Beta(a_j, b_j) - 0.5) with Beta = Ga / (Ga + Gb)  (distributions.py:120-135, 217-229)
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % D);
    const float ga = pz_gamma_draw(sh.a[j], seed, (unsigned long long)(e + e0), 2u);
    const float gb = pz_gamma_draw(sh.b[j], seed, (unsigned long long)(e + e0), 3u);
    U[e] = (2.0f * B) * (ga / (ga + gb) - 0.5f);
  }
}

__global__ void __launch_bounds__(256)
sample_normal_kernel(float* __restrict__ U, long long n, unsigned long long seed, long long q0) {
  const long long nq = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    uint32_t w[4];
    pz_philox4(seed, (unsigned long long)(q + q0), 0u, 4u, w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float r = sqrtf(-2.0f * logf(pz_u01(w[2 * h])));
      float sn, cs;
      sincospif(2.0f * pz_u01(w[2 * h + 1]), &sn, &cs);
      const long long i = q * 4 + 2 * h;
      if (i < n) U[i] = r * cs;
      if (i + 1 < n) U[i + 1] = r * sn;
    }
  }
}

__global__ void __launch_bounds__(256)
sample_tdist_kernel(float* __restrict__ U, long long N, int D, float nu, unsigned long long seed, long long r0) {
  // z / sqrt(chi2(nu) / nu) with one chi2 = 2 Gamma(nu / 2) per row (distributions.py:377-392, unit scale matrix)
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
    const float chi2 = 2.0f * pz_gamma_draw(0.5f * nu, seed, (unsigned long long)(r + r0), 5u);
    const float scale = rsqrtf(chi2 / nu);
    for (int j0 = 0; j0 < D; j0 += 4) {
      uint32_t w[4];
      pz_philox4(seed, (unsigned long long)(r + r0), (uint32_t)(j0 >> 2), 6u, w);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float rr = sqrtf(-2.0f * logf(pz_u01(w[2 * h])));
        float sn, cs;
        sincospif(2.0f * pz_u01(w[2 * h + 1]), &sn, &cs);
        const int j = j0 + 2 * h;
        if (j < D) U[r * D + j] = rr * cs * scale;
        if (j + 1 < D) U[r * D + j + 1] = rr * sn * scale;
      }
    }
  }
}

int pzb_sample_centbeta(float* U, int64_t N, int32_t D, const float* a, const float* b, float B, uint64_t seed,
                        int64_t row_offset, void* stream) {
  if (N < 0 || D < 1 || D > MAX_D || row_offset < 0) return fail(PZB_ERR_INVALID, "sample_centbeta: bad shape (1 <= D <= %d)", MAX_D);
  if (N == 0) return PZB_OK;
  if (!U || !a || !b) return fail(PZB_ERR_INVALID, "sample_centbeta: NULL buffer");
  BetaShapes sh;
  for (int j = 0; j < D; ++j) {
    if (!(a[j] > 0.0f) || !(b[j] > 0.0f)) return fail(PZB_ERR_INVALID, "sample_centbeta: shapes must be positive");
    sh.a[j] = a[j];
    sh.b[j] = b[j];
  }
  sample_centbeta_kernel<<<grid_for(N * (long long)D), 256, 0, (cudaStream_t)stream>>>(U, N * (long long)D, D, sh, B, seed,
                                                                                   row_offset * (long long)D);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_centbeta launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_sample_normal(float* U, int64_t N, int32_t D, uint64_t seed, int64_t row_offset, void* stream) {
  if (N < 0 || D < 1 || row_offset < 0) return fail(PZB_ERR_INVALID, "sample_normal: bad shape");
  if ((row_offset * (int64_t)D) % 4 != 0)
    return fail(PZB_ERR_INVALID, "sample_normal: row_offset * D must be a multiple of 4 (draws come in blocks of four)");
  if (N == 0) return PZB_OK;
  if (!U) return fail(PZB_ERR_INVALID, "sample_normal: NULL buffer");
  const long long n = N * (long long)D;
  sample_normal_kernel<<<grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(U, n, seed, row_offset * (long long)D / 4);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_normal launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_sample_tdist(float* U, int64_t N, int32_t D, float nu, uint64_t seed, int64_t row_offset, void* stream) {
  if (N < 0 || D < 1 || !(nu > 0.0f) || row_offset < 0) return fail(PZB_ERR_INVALID, "sample_tdist: bad shape or degrees of freedom");
  if (N == 0) return PZB_OK;
  if (!U) return fail(PZB_ERR_INVALID, "sample_tdist: NULL buffer");
  sample_tdist_kernel<<<grid_for(N), 256, 0, (cudaStream_t)stream>>>(U, N, D, nu, seed, row_offset);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_tdist launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// ---------------------------------------------------------------------------------------------
// Gaussian error convolution (pzflow/flow.py:339-395, 448-464, 675-724): err_samples draws per row,
// generated inside the chain kernel from a counter-based stream, reduced by two small kernels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logmeanexp_samples_kernel(const float* __restrict__ lp, long long N, int S, float* __restrict__ out) {
  // out[i] = log(mean_s exp(lp[i, s])), the naive form of pzflow/flow.py:463-464
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < S; ++k) s += expf(lp[(size_t)i * S + k]);
    out[i] = logf(s / (float)S);
  }
}

__global__ void __launch_bounds__(256)
mean_samples_kernel(const float* __restrict__ p, long long Ng, int S, int G, float* __restrict__ out) {
  // out[g, i] = mean_s p[g, s, i] (pzflow/flow.py:722-724)
  const long long n = Ng * (long long)G;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long g = e / G;
    const int i = (int)(e - g * G);
    float s = 0.0f;
    for (int k = 0; k < S; ++k) s += p[((size_t)g * S + k) * G + i];
    out[e] = s / (float)S;
  }
}

__global__ void __launch_bounds__(256)
error_eps_kernel(float* __restrict__ eps, long long n, int W, uint32_t stream_id, unsigned long long seed) {
  const long long total = n * (long long)W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long es = e / W;
    eps[e] = pz_err_normal(seed, stream_id, (unsigned long long)es, (int)(e - es * W));
  }
}


// RationalQuadraticSpline as the stand-alone function the reference exports (pzflow/utils.py:64-227): inputs [N, L],
// bin widths / heights W, H [N, L, K] and derivatives D [N, L, K - 1] (K if periodic) given, not produced by a
// conditioner.  One thread per row walks its L dimensions in order (the log-det sum keeps the reference's order);
// knots by the sequential fp32 cumulative sum, bin = (#knots <= x) - 1, gathers out of range -> NaN, same op order as
// the SIMT engine's spline (pz_rqs_eval).
__global__ void __launch_bounds__(128)
rqs_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ H, const float* __restrict__ D,
           long long N, int L, int K, float B, int periodic, int inverse, float* __restrict__ Y, float* __restrict__ LD) {
  const int nd = K - 1 + (periodic ? 1 : 0);
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < N; row += (long long)gridDim.x * blockDim.x) {
    float lds = 0.0f;
    for (int l = 0; l < L; ++l) {
      const float* w = W + (row * L + l) * (long long)K;
      const float* h = H + (row * L + l) * (long long)K;
      const float* d = D + (row * L + l) * (long long)nd;
      const float x = X[row * L + l];
      const bool oob = (x <= -B) || (x >= B);
      const float mx = oob ? fabsf(x) - B : x;
      const float* srch = inverse ? h : w;
      int idx = (-B <= mx) ? 0 : -1;
      float c = 0.0f;
      for (int k = 0; k < K; ++k) {
        c = (k == 0) ? srch[0] : c + srch[k];
        idx += ((-B + c) <= mx) ? 1 : 0;
      }
      const bool bad = idx < 0 || idx >= K;
      const int ic = min(max(idx, 0), K - 1);
      float cw = 0.0f, ch = 0.0f;
      for (int k = 0; k < ic; ++k) {
        cw = (k == 0) ? w[0] : cw + w[k];
        ch = (k == 0) ? h[0] : ch + h[k];
      }
      const float qnan = __int_as_float(0x7fc00000);
      SplineKnotSel s;
      s.idx = idx;
      s.wk = bad ? qnan : w[ic];
      s.hk = bad ? qnan : h[ic];
      s.xk = bad ? qnan : (ic == 0 ? -B : -B + cw);
      s.yk = bad ? qnan : (ic == 0 ? -B : -B + ch);
      if (periodic) {
        s.dk = bad ? qnan : (ic == 0 ? d[K - 1] : d[ic - 1]);
        s.dkp1 = bad ? qnan : d[ic];
      } else {
        s.dk = bad ? qnan : (ic == 0 ? 1.0f : d[ic - 1]);
        s.dkp1 = bad ? qnan : (ic == K - 1 ? 1.0f : d[ic]);
      }
      float out, ld;
      pz_rqs_eval<false>(s, x, mx, oob, inverse != 0, periodic != 0, out, ld);
      Y[row * L + l] = out;
      lds += ld;
    }
    LD[row] = inverse ? -lds : lds;
  }
}

int pzb_rational_quadratic_spline(const float* inputs, const float* W, const float* H, const float* D, int64_t N, int32_t L,
                                  int32_t K, float B, int32_t periodic, int32_t inverse, float* outputs, float* log_det,
                                  void* stream) {
  if (N < 0 || L < 1 || K < 1) return fail(PZB_ERR_INVALID, "rational_quadratic_spline: bad shape");
  if (N == 0) return PZB_OK;
  if (!inputs || !W || !H || (!D && K - 1 + (periodic ? 1 : 0) > 0) || !outputs || !log_det)
    return fail(PZB_ERR_INVALID, "rational_quadratic_spline: NULL buffer");
  long long blocks = (N + 127) / 128;
  if (blocks > (long long)current_num_sms() * 32) blocks = (long long)current_num_sms() * 32;
  rqs_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(inputs, W, H, D, N, L, K, B, periodic, inverse, outputs, log_det);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "rational_quadratic_spline launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_error_eps(float* eps, int64_t n, int32_t W, int32_t stream_id, uint64_t seed, void* stream) {
  if (n < 0 || W < 1 || (stream_id != 0 && stream_id != 1)) return fail(PZB_ERR_INVALID, "error_eps: bad arguments");
  if (n == 0) return PZB_OK;
  if (!eps) return fail(PZB_ERR_INVALID, "error_eps: NULL buffer");
  error_eps_kernel<<<grid_for(n * W), 256, 0, (cudaStream_t)stream>>>(eps, n, W, (uint32_t)stream_id, seed);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "error_eps launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// Error samples in jax's own stream (pzb_set_error_stream): fills the launch's eps fields, or reports why it cannot.
static int eps_stream_args(const pzb_plan* plan, LaunchArgs& a, uint64_t seed, int64_t row_offset, int64_t units,
                           int32_t S, const char* who) {
  if (g_eps_layout == 0) return PZB_OK;
  if (g_eps_units < row_offset + units)
    return fail(PZB_ERR_INVALID, "%s: pzb_set_error_stream announced %lld units, the call covers [%lld, %lld)", who,
                g_eps_units, (long long)row_offset, (long long)(row_offset + units));
  const long long widest = std::max(plan->host.D, plan->host.C);
  if (g_eps_layout == 1 && g_eps_units * (long long)S * widest >= 0xFFFFFFFFLL)
    return fail(PZB_ERR_UNSUPPORTED, "%s: %lld x %d x %lld draws exceed one threefry call of jax's original layout", who,
                g_eps_units, S, widest);
  a.eps_layout = g_eps_layout;
  a.eps_k0 = 0u;                      // PRNGKey(seed) with x64 disabled: (0, seed mod 2^32)
  a.eps_k1 = (uint32_t)seed;
  a.eps_units = g_eps_units;
  return PZB_OK;
}

int pzb_set_error_stream(int32_t layout, int64_t total_units) {
  if (layout < 0 || layout > 2 || total_units < 0) return fail(PZB_ERR_INVALID, "set_error_stream: bad arguments");
  g_eps_layout = layout;
  g_eps_units = total_units;
  return PZB_OK;
}

// jax.random.normal(key, (total_rows, W)) -- rows [row_offset, row_offset + N) of it -- in jax's stream layout: the
// Normal latent's draws (pzflow/distributions.py:297-303: multivariate_normal with the identity covariance IS normal)
// and the default error model's eps (pzflow/utils.py:59) for callers that want them materialised.
__global__ void __launch_bounds__(256)
sample_normal_threefry_kernel(float* __restrict__ out, long long n, long long e0, long long n_total, uint32_t k0, uint32_t k1,
                              int partitionable) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    out[e] = pz_jax_normal(pz_jax_bits(k0, k1, e + e0, n_total, partitionable != 0));
}

int pzb_sample_normal_threefry(float* out, int64_t N, int32_t W, uint32_t key0, uint32_t key1, int64_t total_rows,
                               int64_t row_offset, int32_t partitionable, void* stream) {
  if (N < 0 || W < 1 || row_offset < 0 || total_rows < row_offset + N)
    return fail(PZB_ERR_INVALID, "sample_normal_threefry: bad shape");
  if (N == 0) return PZB_OK;
  if (!out) return fail(PZB_ERR_INVALID, "sample_normal_threefry: NULL buffer");
  const long long n_total = total_rows * (long long)W;
  if (!partitionable && n_total >= 0xFFFFFFFFLL)
    return fail(PZB_ERR_UNSUPPORTED, "sample_normal_threefry: %lld elements exceed one threefry call of jax's original layout",
                n_total);
  const long long n = N * (long long)W;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)current_num_sms() * 32) blocks = (long long)current_num_sms() * 32;
  sample_normal_threefry_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, row_offset * (long long)W, n_total,
                                                                                   key0, key1, partitionable ? 1 : 0);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "sample_normal_threefry launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_log_prob_err(const pzb_plan* plan, const float* X, const float* Xerr, const float* cond,
                     const float* cond_err, int64_t N, int32_t err_samples, uint64_t seed, int64_t row_offset,
                     float* out, float* scratch, void* stream) {
  int rc = check_common(plan, X, cond, N, out, "log_prob_err");
  if (rc) return rc;
  if (err_samples < 1) return fail(PZB_ERR_INVALID, "log_prob_err: err_samples must be positive");
  if (N == 0) return PZB_OK;
  const bool fused = posterior_fuses(plan, 1, 0, err_samples, N);
  if (!fused && scratch == nullptr) return fail(PZB_ERR_INVALID, "log_prob_err: scratch [N * err_samples] is NULL");
  LaunchArgs a{};
  a.mode = MODE_LOGPROB;
  a.X = X;
  a.Xerr = Xerr;
  a.cond = cond;
  a.cond_err = plan->host.C > 0 ? cond_err : nullptr;
  a.S = err_samples;
  a.seed = seed;
  a.unit0 = row_offset;
  a.N = N * (int64_t)err_samples;
  a.out0 = scratch;
  rc = eps_stream_args(plan, a, seed, row_offset, N, err_samples, "log_prob_err");
  if (rc) return rc;
  if (fused) {  // log(mean_s exp(lp)) taken in shared memory: no [N, S] scratch
    a.fuse = 1;
    a.G = 1;
    a.out0 = out;
    return launch_chain(plan, a, stream);
  }
  rc = launch_chain(plan, a, stream);
  if (rc) return rc;
  DeviceGuard guard(plan->device);
  logmeanexp_samples_kernel<<<grid_for(N), 256, 0, (cudaStream_t)stream>>>(scratch, N, err_samples, out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "logmeanexp_samples launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_posterior_err(const pzb_plan* plan, const float* rest, const float* rest_err, const float* cond,
                      const float* cond_err, int64_t Ng, int32_t col_idx, const float* grid, int32_t G,
                      int32_t err_samples, uint64_t seed, int64_t row_offset, int32_t normalize, int32_t nan_to_zero,
                      float* pdfs, float* scratch, void* stream) {
  int rc = check_common(plan, plan && plan->host.D > 1 ? (const void*)rest : (const void*)grid, cond, Ng, pdfs,
                        "posterior_err");
  if (rc) return rc;
  if (G < 1 || grid == nullptr) return fail(PZB_ERR_INVALID, "posterior_err: empty grid");
  if (col_idx < 0 || col_idx >= plan->host.D)
    return fail(PZB_ERR_INVALID, "posterior_err: column index %d out of range", col_idx);
  if (err_samples < 1) return fail(PZB_ERR_INVALID, "posterior_err: err_samples must be positive");
  if (Ng == 0) return PZB_OK;
  const bool fused = posterior_fuses(plan, G, 0, err_samples, Ng);
  if (!fused && scratch == nullptr) return fail(PZB_ERR_INVALID, "posterior_err: scratch [Ng * err_samples * G] is NULL");
  LaunchArgs a{};
  a.mode = MODE_POSTERIOR;
  a.X = rest;
  a.Xerr = plan->host.D > 1 ? rest_err : nullptr;
  a.cond = cond;
  a.cond_err = plan->host.C > 0 ? cond_err : nullptr;
  a.grid = grid;
  a.G = G;
  a.col_idx = col_idx;
  a.S = err_samples;
  a.seed = seed;
  a.unit0 = row_offset;
  a.N = Ng * (int64_t)err_samples * (int64_t)G;
  a.out0 = scratch;
  rc = eps_stream_args(plan, a, seed, row_offset, Ng, err_samples, "posterior_err");
  if (rc) return rc;
  if (fused) {
    a.fuse = 1;
    a.normalize = normalize;
    a.nan_to_zero = nan_to_zero;
    a.out0 = pdfs;
    return launch_chain(plan, a, stream);
  }
  rc = launch_chain(plan, a, stream);
  if (rc) return rc;
  DeviceGuard guard(plan->device);
  mean_samples_kernel<<<grid_for(Ng * (long long)G), 256, 0, (cudaStream_t)stream>>>(scratch, Ng, err_samples, G, pdfs);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "mean_samples launch");
  g_launches.fetch_add(1);
  return pzb_normalize_pdfs(pdfs, Ng, G, grid, normalize, nan_to_zero, stream);
}

// ---------------------------------------------------------------------------------------------
// Posterior marginalised over missing columns (pzflow/flow.py:560-664): the variants of a galaxy (its row with the
// missing columns replaced by every point of their marginal grids) are a second expansion axis generated inside the
// chain kernel; this kernel takes the mean over the error samples of a variant (flow.py:722-724), maps NaN to 0 per
// variant (flow.py:735-737 of the inner call) and sums the variants (flow.py:651-654).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
marg_reduce_kernel(const float* __restrict__ p, long long Ng, int V, int S, int G, int nan_to_zero, float* __restrict__ out) {
  const long long n = Ng * (long long)G;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long g = e / G;
    const int i = (int)(e - g * G);
    float total = 0.0f;
    for (int v = 0; v < V; ++v) {
      const float* pv = p + ((size_t)(g * V + v) * S) * G + i;
      float y = pv[0];
      if (S > 1) {
        float s = 0.0f;
        for (int k = 0; k < S; ++k) s += pv[(size_t)k * G];
        y = s / (float)S;
      }
      if (nan_to_zero) y = nan_to_num0(y);
      total = v == 0 ? y : total + y;
    }
    out[e] = total;
  }
}

int pzb_posterior_marg(const pzb_plan* plan, const float* rest, const float* rest_err, const float* cond,
                       const float* cond_err, int64_t Ng, int32_t col_idx, const float* grid, int32_t G,
                       int32_t n_mcols, const int32_t* mcols, const float* mvals, int32_t V, int32_t err_samples,
                       uint64_t seed, int64_t row_offset, int32_t nan_to_zero, float* pdf_sums, float* scratch,
                       void* stream) {
  int rc = check_common(plan, plan && plan->host.D > 1 ? (const void*)rest : (const void*)grid, cond, Ng, pdf_sums,
                        "posterior_marg");
  if (rc) return rc;
  if (G < 1 || grid == nullptr) return fail(PZB_ERR_INVALID, "posterior_marg: empty grid");
  if (col_idx < 0 || col_idx >= plan->host.D) return fail(PZB_ERR_INVALID, "posterior_marg: column index %d out of range", col_idx);
  if (n_mcols < 1 || n_mcols > 4 || mcols == nullptr || mvals == nullptr || V < 1)
    return fail(PZB_ERR_INVALID, "posterior_marg: between 1 and 4 marginalised columns with V >= 1 variants each are supported");
  if (err_samples < 0) return fail(PZB_ERR_INVALID, "posterior_marg: negative err_samples");
  for (int q = 0; q < n_mcols; ++q) {
    if (mcols[q] < 0 || mcols[q] >= plan->host.D || mcols[q] == col_idx)
      return fail(PZB_ERR_INVALID, "posterior_marg: marginalised column %d is not one of the other data columns", mcols[q]);
    for (int r = 0; r < q; ++r)
      if (mcols[r] == mcols[q]) return fail(PZB_ERR_INVALID, "posterior_marg: column %d listed twice", mcols[q]);
  }
  if (Ng == 0) return PZB_OK;
  const bool fused = posterior_fuses(plan, G, V, err_samples, Ng);
  if (!fused && scratch == nullptr) return fail(PZB_ERR_INVALID, "posterior_marg: scratch [Ng * V * max(S, 1) * G] is NULL");
  const int S = err_samples > 0 ? err_samples : 1;
  LaunchArgs a{};
  a.mode = MODE_POSTERIOR;
  a.X = rest;
  a.Xerr = (err_samples > 0 && plan->host.D > 1) ? rest_err : nullptr;
  a.cond = cond;
  a.cond_err = (err_samples > 0 && plan->host.C > 0) ? cond_err : nullptr;
  a.grid = grid;
  a.G = G;
  a.col_idx = col_idx;
  a.S = err_samples > 0 ? err_samples : 0;
  a.seed = seed;
  a.unit0 = row_offset;
  a.mvals = mvals;
  a.V = V;
  a.n_mcols = n_mcols;
  for (int q = 0; q < 4; ++q) a.mcol[q] = q < n_mcols ? mcols[q] : -1;
  a.N = Ng * (int64_t)V * (int64_t)S * (int64_t)G;
  a.out0 = scratch;
  if (fused) {
    a.fuse = 1;
    a.normalize = 0;
    a.nan_to_zero = nan_to_zero;
    a.out0 = pdf_sums;
    return launch_chain(plan, a, stream);
  }
  rc = launch_chain(plan, a, stream);
  if (rc) return rc;
  DeviceGuard guard(plan->device);
  marg_reduce_kernel<<<grid_for(Ng * (long long)G), 256, 0, (cudaStream_t)stream>>>(scratch, Ng, V, S, G, nan_to_zero, pdf_sums);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return cuda_fail(ce, "marg_reduce launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// ---------------------------------------------------------------------------------------------
// Training support: the coupling stage's spline, forward and backward (nsf_train.cu)
// ---------------------------------------------------------------------------------------------
int pzb_spline_train_forward(const float* logits, const float* x, int64_t N, int32_t L, int32_t K, float B,
                             int32_t periodic, float* y, float* log_det, void* stream) {
  if (N < 0 || L < 1 || K < 2 || K > 64) return fail(PZB_ERR_INVALID, "spline_train_forward: bad shape (K in [2, 64])");
  if (N > 0 && (!logits || !x || !y || !log_det)) return fail(PZB_ERR_INVALID, "spline_train_forward: NULL buffer");
  cudaError_t ce = launch_spline_train_fwd(logits, x, N * (long long)L, K, B, periodic, y, log_det, (cudaStream_t)stream);
  if (ce != cudaSuccess) return cuda_fail(ce, "spline_train_forward launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

int pzb_spline_train_backward(const float* logits, const float* x, const float* grad_y, const float* grad_log_det,
                              int64_t N, int32_t L, int32_t K, float B, int32_t periodic, float* grad_logits,
                              float* grad_x, void* stream) {
  if (N < 0 || L < 1 || K < 2 || K > 64) return fail(PZB_ERR_INVALID, "spline_train_backward: bad shape (K in [2, 64])");
  if (N > 0 && (!logits || !x || !grad_y || !grad_log_det || !grad_logits || !grad_x))
    return fail(PZB_ERR_INVALID, "spline_train_backward: NULL buffer");
  cudaError_t ce = launch_spline_train_bwd(logits, x, grad_y, grad_log_det, N * (long long)L, K, B, periodic, grad_logits,
                                           grad_x, (cudaStream_t)stream);
  if (ce != cudaSuccess) return cuda_fail(ce, "spline_train_backward launch");
  g_launches.fetch_add(1);
  return PZB_OK;
}

// ---- the whole gradient of one optimisation step as one kernel (nsf_train_step.cu) ----
int pzb_plan_train_supported(const pzb_plan* plan) { return plan && train_step_supported(plan->host) ? 1 : 0; }

int pzb_plan_set_train_layout(pzb_plan* plan, const int64_t* w_offsets, const int64_t* b_offsets, int32_t n_layers_total,
                              int64_t n_params) {
  if (!plan || !w_offsets || !b_offsets) return fail(PZB_ERR_INVALID, "set_train_layout: NULL argument");
  if (!train_step_supported(plan->host))
    return fail(PZB_ERR_UNSUPPORTED, "set_train_layout: this chain is outside the fused training step's shapes");
  const PlanDev& hp = plan->host;
  int total = 0;
  for (int s = 0; s < hp.n_nsc; ++s) total += hp.nsc[s].n_dense;
  if (total != n_layers_total) return fail(PZB_ERR_INVALID, "set_train_layout: expected %d dense layers, got %d", total, n_layers_total);
  std::vector<TrainOffsets> offs(hp.n_nsc);
  long long covered = 0;
  int q = 0;
  for (int s = 0; s < hp.n_nsc; ++s) {
    const NscDev& st = hp.nsc[s];
    int in_dim = st.U + st.C;
    for (int li = 0; li < st.n_dense; ++li, ++q) {
      const int out_dim = li == st.n_dense - 1 ? st.cpd * st.L : st.H;
      const long long wn = (long long)in_dim * out_dim;
      if (w_offsets[q] < 0 || b_offsets[q] < 0 || w_offsets[q] + wn > n_params || b_offsets[q] + out_dim > n_params)
        return fail(PZB_ERR_INVALID, "set_train_layout: layer %d does not fit the flat parameter buffer", q);
      offs[s].w[li] = w_offsets[q];
      offs[s].b[li] = b_offsets[q];
      covered += wn + out_dim;
      in_dim = st.H;
    }
  }
  if (covered != n_params)
    return fail(PZB_ERR_INVALID, "set_train_layout: the dense layers cover %lld of %lld parameters", covered, (long long)n_params);
  DeviceGuard guard(plan->device);
  if (!plan->d_train_offs) {
    cudaError_t ce = cudaMalloc(reinterpret_cast<void**>(&plan->d_train_offs), sizeof(TrainOffsets) * MAX_NSC);
    if (ce != cudaSuccess) return cuda_fail(ce, "set_train_layout");
  }
  cudaError_t ce = cudaMemcpy(plan->d_train_offs, offs.data(), sizeof(TrainOffsets) * hp.n_nsc, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return cuda_fail(ce, "set_train_layout upload");
  plan->train_params = n_params;
  return PZB_OK;
}

int64_t pzb_train_workspace_floats(const pzb_plan* plan, int64_t N) {
  if (!plan || N < 0 || plan->train_params <= 0) return -1;
  return (int64_t)train_step_workspace_floats(plan->host, N, plan->train_params);
}

int pzb_train_step(const pzb_plan* plan, const float* params, const float* X, const float* cond, const float* weights,
                   int64_t N, int64_t n_global, const int64_t* row_idx, const float* Xerr, const float* cond_err,
                   uint64_t seed, uint64_t noise_base, float* workspace, float* flat_grad, float* loss, void* stream) {
  if (!plan || !plan->d_train_offs) return fail(PZB_ERR_INVALID, "train_step: call pzb_plan_set_train_layout first");
  if (N < 0 || n_global < 1) return fail(PZB_ERR_INVALID, "train_step: bad row counts");
  if (!params || !flat_grad || (N > 0 && (!X || !weights || !workspace)))
    return fail(PZB_ERR_INVALID, "train_step: NULL buffer");
  if (N > 0 && plan->host.C > 0 && !cond) return fail(PZB_ERR_INVALID, "train_step: this plan is conditional but conditions are NULL");
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(PZB_ERR_CUDA, "cannot select device %d", plan->device);
  TrainBatch tb;
  tb.row_idx = reinterpret_cast<const long long*>(row_idx);
  tb.Xerr = Xerr;
  tb.cond_err = plan->host.C > 0 ? cond_err : nullptr;
  tb.seed = seed;
  tb.noise_base = noise_base;
  cudaError_t ce = launch_train_step(plan->d_plan, plan->host, plan->d_train_offs, params, X, cond, weights, N,
                                     1.0f / (float)n_global, plan->train_params, workspace, flat_grad, loss, plan->num_sms,
                                     tb, (cudaStream_t)stream);
  if (ce != cudaSuccess) return cuda_fail(ce, "train_step launch");
  g_launches.fetch_add(N > 0 ? 2 : 0);
  return PZB_OK;
}

int pzb_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float b1, float b2,
                  float eps, int32_t* step_dev, void* stream) {
  if (n < 0 || !step_dev || (n > 0 && (!params || !grads || !m || !v)))
    return fail(PZB_ERR_INVALID, "adam_step: NULL buffer or negative size");
  cudaError_t ce = launch_adam_step(params, grads, m, v, n, lr, b1, b2, eps, step_dev, (cudaStream_t)stream);
  if (ce != cudaSuccess) return cuda_fail(ce, "adam_step launch");
  g_launches.fetch_add(2);
  return PZB_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry points: the same evaluations for callers whose arrays live in HOST memory
// (the reference's NumPy / DataFrame world, pzflow/flow.py:440-446).  Rows are cut into chunks
// that flow through a three-deep ring of device staging buffers on three streams, so the H2D
// copy of chunk i, the chain kernel of chunk i-1 and the D2H copy of chunk i-2 overlap; HBM
// never holds more than three chunks.  Blocking: results are in the host buffers on return.
// Pinned (page-locked) host arrays are copied from / into directly; pageable ones (plain NumPy arrays) of 8 MB or
// more go through plan-owned page-locked staging buffers, filled / emptied by all host cores chunk by chunk inside
// the same pipeline, so they overlap too.
// ---------------------------------------------------------------------------------------------
struct HostJob {
  int mode = MODE_LOGPROB;
  const float* in = nullptr;   // [units, in_w]
  bool in_on_device = false;   // `in` / `cond` are DEVICE arrays (latent draws made on the device): no H2D stage
  int in_w = 0;
  const float* cond = nullptr; // [units, C] or NULL
  float* out0 = nullptr;       // [units, out0_w]
  float* const* out_cols = nullptr;  // instead of out0: out0_w HOST column arrays of `units` values each (DataFrame layout)
  int out0_w = 1;
  float* out1 = nullptr;       // [units] or NULL
  const float* grid = nullptr; // posterior: host [G]
  int G = 1, col_idx = 0, normalize = 0, nan_to_zero = 0;
  int64_t units = 0;           // rows, or galaxies for the posterior
  // column-source form: in_w (and C) float64 column pointers instead of `in` / `cond`; conditions are standard-scaled
  // here, (float(c) - mean) / std in float32 like Flow._get_conditions (pzflow/flow.py:330-336)
  const void* const* in_cols = nullptr;
  const void* const* cond_cols = nullptr;
  bool cols_f32 = false;       // the columns are float32 (else float64)
  // ensemble form (log_prob only): every chunk is evaluated by all n_flows plans (plan == flows[0]) and reduced with
  // log(mean(exp)) (pzflow/flowEnsemble.py:243) before it leaves the device
  const pzb_plan* const* flows = nullptr;
  int n_flows = 0;
  const float* cond_mean = nullptr;
  const float* cond_std = nullptr;
  int64_t base = 0;            // unit offset into the COLUMN arrays (in_cols / cond_cols / out_cols) of a sliced job
};

static cudaError_t host_reserve(void** p, size_t* cap, size_t want, int n) {
  if (want <= *cap) return cudaSuccess;
  for (int b = 0; b < n; ++b) {
    if (p[b]) cudaFreeHost(p[b]);
    p[b] = nullptr;
  }
  *cap = 0;
  for (int b = 0; b < n; ++b) {
    cudaError_t e = cudaHostAlloc(&p[b], want, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
  }
  *cap = want;
  return cudaSuccess;
}


static cudaError_t pipe_reserve(void** p, size_t* cap, size_t want, int n) {
  if (want <= *cap) return cudaSuccess;
  for (int b = 0; b < n; ++b) {
    if (p[b]) cudaFree(p[b]);
    p[b] = nullptr;
  }
  *cap = 0;
  for (int b = 0; b < n; ++b) {
    cudaError_t e = cudaMalloc(&p[b], want);
    if (e != cudaSuccess) return e;
  }
  *cap = want;
  return cudaSuccess;
}

// Host threads for the casts / copies of one pipeline: the CPUs this process may run on (its affinity mask, which
// sharding.bind_to_gpu_numa_node narrows), shared fairly when torchrun runs several ranks on the box, at most 16.
static thread_local int g_host_threads = 0;  // pzb_set_host_threads: this host thread's budget for its pipelines
static int host_thread_count() {
  if (g_host_threads > 0) return g_host_threads;
  int n = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (const char* lws = getenv("LOCAL_WORLD_SIZE")) {
    const int w = atoi(lws);
    if (w > 1) n /= w;
  }
  return std::max(1, std::min(16, n));
}

// memcpy on all host cores (first-touch page faults of a fresh destination included)
static void parallel_copy(void* dst, const void* src, size_t bytes, int threads) {
  const size_t blk = 1u << 20;
  const int64_t nblk = (int64_t)((bytes + blk - 1) / blk);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < nblk; ++i) {
    const size_t a = (size_t)i * blk, n = std::min(blk, bytes - a);
    memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, n);
  }
}

// true if the CUDA driver knows p as page-locked host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory)
static bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

struct UnitSpan {
  int64_t u0, nu;
};
static int run_host_once(const pzb_plan* plan, const HostJob& job, const char* who, std::vector<UnitSpan>* flagged);

// units [u0, u0 + nu) of a job as a job of its own
static HostJob job_slice(const HostJob& job, int C, int64_t u0, int64_t nu) {
  HostJob s = job;
  if (s.in) s.in += u0 * s.in_w;
  if (s.cond) s.cond += u0 * C;
  if (s.out0) s.out0 += u0 * s.out0_w;
  if (s.out1) s.out1 += u0;
  s.base = job.base + u0;
  s.units = nu;
  return s;
}

// The fp16 guard of an AUTO-engine plan: the tensor-core engines carry activations as fp16 hi/lo pairs, so a
// conditioner pre-activation beyond +-65504 turns a row into NaN where the reference's fp32 stays finite (catalogue
// placeholders like -9999 are enough).  Such rows (finite inputs, non-finite logits) are counted on the device and every
// tensor-core launch of the pipeline reports whether ITS chunk held one; the flagged chunks -- not the whole call --
// are evaluated again on the SIMT engine (fp32 throughout), so the caller gets the reference's answer either way and
// a few bad rows in 10 M cost a few chunks, not a tenfold slower call.
static int run_host(const pzb_plan* plan, const HostJob& job, const char* who) {
  if (job.units == 0) return PZB_OK;
  const bool guarded = plan->engine == PZB_ENGINE_AUTO && !g_force_simt && pzb_plan_engine(plan) != PZB_ENGINE_SIMT &&
                       job.n_flows == 0;
  if (!guarded) return run_host_once(plan, job, who, nullptr);
  std::vector<UnitSpan> flagged;
  int rc = run_host_once(plan, job, who, &flagged);
  if (rc != PZB_OK || flagged.empty()) return rc;
  g_force_simt = true;
  for (size_t k = 0; k < flagged.size() && rc == PZB_OK; ++k) {
    UnitSpan sp = flagged[k];
    while (k + 1 < flagged.size() && flagged[k + 1].u0 == sp.u0 + sp.nu) sp.nu += flagged[++k].nu;  // adjacent chunks together
    rc = run_host_once(plan, job_slice(job, plan->host.C, sp.u0, sp.nu), who, nullptr);
  }
  g_force_simt = false;
  unsigned long long now = 0;
  DeviceGuard guard(plan->device);
  if (cudaMemcpy(&now, plan->d_overflow, sizeof(now), cudaMemcpyDeviceToHost) == cudaSuccess) plan->overflow_seen.store(now);
  return rc;
}

static int run_host_once(const pzb_plan* plan, const HostJob& job, const char* who, std::vector<UnitSpan>* flagged) {
  HostPipe& hp = plan->pipe;
  std::lock_guard<std::mutex> lock(hp.mu);
  DeviceGuard guard(plan->device);
  if (!guard.ok) return fail(PZB_ERR_CUDA, "%s: cannot select device %d", who, plan->device);
  cudaError_t ce = cudaSuccess;
  if (!hp.ready) {
    ce = cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&hp.s_k, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking);
    for (int b = 0; b < HostPipe::NBUF && ce == cudaSuccess; ++b) {
      ce = cudaEventCreateWithFlags(&hp.ev_in[b], cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&hp.ev_k[b], cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&hp.ev_out[b], cudaEventDisableTiming);
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "host pipeline setup");
    hp.ready = true;
  }
  const int C = plan->host.C;
  const int host_threads = host_thread_count();
  const int rows_per_unit = (job.mode == MODE_POSTERIOR) ? job.G : 1;
  // chunk: ~1.2 M chain rows (8 waves of 148 CTAs x 1024 rows), staging buffers <= 64 MB each
  int64_t chunk = (int64_t)plan->num_sms * 1024 * 8 / rows_per_unit;
  const int64_t widest = (int64_t)sizeof(float) * std::max<int64_t>(std::max(job.in_w, C), job.out0_w);
  if (widest > 0 && chunk * widest > (64ll << 20)) chunk = (64ll << 20) / widest;
  if (chunk < 1) chunk = 1;
  if (chunk > job.units) chunk = job.units;
  // every ring buffer may still be in flight from this call only (the previous call drained)
  if (!job.in_on_device)
    ce = pipe_reserve(hp.d_in, &hp.cap_in, (size_t)chunk * std::max(job.in_w, 1) * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && C > 0 && !job.in_on_device)
    ce = pipe_reserve(hp.d_cond, &hp.cap_cond, (size_t)chunk * C * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess)
    ce = pipe_reserve(hp.d_out0, &hp.cap_out0, (size_t)chunk * job.out0_w * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && job.out1)
    ce = pipe_reserve(hp.d_out1, &hp.cap_out1, (size_t)chunk * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && job.n_flows > 0)
    ce = pipe_reserve(hp.d_ens, &hp.cap_ens, (size_t)chunk * job.out0_w * job.n_flows * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && job.grid) {
    ce = pipe_reserve(&hp.d_grid, &hp.cap_grid, (size_t)job.G * sizeof(float), 1);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(hp.d_grid, job.grid, (size_t)job.G * sizeof(float), cudaMemcpyHostToDevice, hp.s_k);
  }
  // float32 rows in pageable memory (a plain NumPy array) are staged like the float64 columns are: copied into
  // page-locked buffers by all host cores, chunk by chunk, so that the H2D copies run at PCIe speed and overlap
  const bool stage_in = job.in != nullptr && job.in_w > 0 && !job.in_on_device &&
                        (size_t)job.units * job.in_w * sizeof(float) >= (8u << 20) && !host_pointer_is_pinned(job.in);
  if (ce == cudaSuccess && (job.in_cols || stage_in))
    ce = host_reserve(hp.h_in, &hp.cap_hin, (size_t)chunk * std::max(job.in_w, 1) * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && C > 0 && (job.cond_cols || (stage_in && job.cond)))
    ce = host_reserve(hp.h_cond, &hp.cap_hcond, (size_t)chunk * C * sizeof(float), HostPipe::NBUF);
  // results bound for pageable memory go through page-locked staging (small results are not worth the detour)
  const size_t out_total = (size_t)job.units * job.out0_w * sizeof(float);
  const bool stage_out = job.out_cols != nullptr || (out_total >= (8u << 20) && !host_pointer_is_pinned(job.out0));
  if (ce == cudaSuccess && stage_out)
    ce = host_reserve(hp.h_out0, &hp.cap_hout0, (size_t)chunk * job.out0_w * sizeof(float), HostPipe::NBUF);
  if (ce == cudaSuccess && stage_out && job.out1)
    ce = host_reserve(hp.h_out1, &hp.cap_hout1, (size_t)chunk * sizeof(float), HostPipe::NBUF);
  std::vector<UnitSpan> spans;  // the chunks of this call, in order (per-chunk guard)
  if (ce == cudaSuccess && flagged != nullptr) {
    // chunk count: the ramp (1/8, 1/4, 1/2) adds at most three chunks
    const size_t n_chunks = (size_t)((job.units + chunk - 1) / chunk) + 4;
    if (hp.d_flag == nullptr) ce = cudaMalloc(reinterpret_cast<void**>(&hp.d_flag), HostPipe::NBUF * sizeof(unsigned int));
    if (ce == cudaSuccess && n_chunks > hp.cap_flags) {
      if (hp.h_flags) cudaFreeHost(hp.h_flags);
      hp.h_flags = nullptr;
      hp.cap_flags = 0;
      ce = cudaMallocHost(reinterpret_cast<void**>(&hp.h_flags), n_chunks * sizeof(unsigned int));
      if (ce == cudaSuccess) hp.cap_flags = n_chunks;
    }
    spans.reserve(n_chunks);
  }
  if (ce != cudaSuccess) return cuda_fail(ce, "host pipeline staging buffers");
  int64_t pend_u0[HostPipe::NBUF] = {}, pend_nu[HostPipe::NBUF] = {};
  // chunk j's staged results -> the caller's arrays, once its D2H copy has completed
  auto drain = [&](int64_t j) {
    const int bj = (int)(j % HostPipe::NBUF);
    cudaError_t e = cudaEventSynchronize(hp.ev_out[bj]);
    if (e != cudaSuccess) return e;
    if (job.out_cols)
      scatter_rows_to_cols(static_cast<const float*>(hp.h_out0[bj]), job.out0_w, job.base + pend_u0[bj], pend_nu[bj],
                           job.out_cols, host_threads);
    else
      parallel_copy(job.out0 + pend_u0[bj] * job.out0_w, hp.h_out0[bj], (size_t)pend_nu[bj] * job.out0_w * sizeof(float),
                    host_threads);
    if (job.out1) memcpy(job.out1 + pend_u0[bj], hp.h_out1[bj], (size_t)pend_nu[bj] * sizeof(float));
    return cudaSuccess;
  };

  int rc = PZB_OK;
  int64_t i = 0;
  // ramp: the first chunks are 1/8, 1/4, 1/2 of the steady-state chunk, so the first kernel starts after a short
  // copy instead of a full-size one (the un-overlapped head of the pipeline)
  int64_t cur = std::max<int64_t>(1, chunk / 8);
  // ... unless the whole call is a couple of waves of CTAs: a kernel launch has a floor of a few tens of microseconds
  // (one tile walks the whole chain), so four small launches in a row cost more than the copy they would hide
  if (job.units * rows_per_unit <= (int64_t)plan->num_sms * 1024 * 2) cur = chunk;
  for (int64_t u0 = 0; u0 < job.units && rc == PZB_OK; ++i) {
    const int b = (int)(i % HostPipe::NBUF);
    const int64_t nu = std::min(cur, job.units - u0);
    const float* src_in = job.in ? job.in + u0 * job.in_w : nullptr;
    const float* src_cond = job.cond ? job.cond + u0 * C : nullptr;
    if (job.in_cols) {
      // the previous copy out of this staging buffer (chunk i - NBUF) must have completed before it is refilled;
      // the cast of chunk i then runs on the host cores while the GPU works on chunks i-1, i-2
      if (i >= HostPipe::NBUF) cudaEventSynchronize(hp.ev_in[b]);
      auto gather = job.cols_f32 ? gather_cast_rows<float> : gather_cast_rows<double>;
      gather(job.in_cols, job.in_w, job.base + u0, nu, static_cast<float*>(hp.h_in[b]), nullptr, nullptr, host_threads);
      src_in = static_cast<const float*>(hp.h_in[b]);
      if (C > 0) {
        gather(job.cond_cols, C, job.base + u0, nu, static_cast<float*>(hp.h_cond[b]), job.cond_mean, job.cond_std,
               host_threads);
        src_cond = static_cast<const float*>(hp.h_cond[b]);
      }
    } else if (stage_in) {
      if (i >= HostPipe::NBUF) cudaEventSynchronize(hp.ev_in[b]);
      parallel_copy(hp.h_in[b], src_in, (size_t)nu * job.in_w * sizeof(float), host_threads);
      src_in = static_cast<const float*>(hp.h_in[b]);
      if (C > 0 && src_cond != nullptr) {
        parallel_copy(hp.h_cond[b], src_cond, (size_t)nu * C * sizeof(float), host_threads);
        src_cond = static_cast<const float*>(hp.h_cond[b]);
      }
    }
    if (i >= HostPipe::NBUF) cudaStreamWaitEvent(hp.s_in, hp.ev_k[b], 0);  // kernel i-NBUF has read d_in[b]
    if (job.in_w > 0 && !job.in_on_device)
      ce = cudaMemcpyAsync(hp.d_in[b], src_in, (size_t)nu * job.in_w * sizeof(float), cudaMemcpyHostToDevice, hp.s_in);
    if (ce == cudaSuccess && C > 0 && !job.in_on_device)
      ce = cudaMemcpyAsync(hp.d_cond[b], src_cond, (size_t)nu * C * sizeof(float), cudaMemcpyHostToDevice, hp.s_in);
    if (ce == cudaSuccess) ce = cudaEventRecord(hp.ev_in[b], hp.s_in);
    if (ce != cudaSuccess) { rc = cuda_fail(ce, "host pipeline H2D"); break; }
    cudaStreamWaitEvent(hp.s_k, hp.ev_in[b], 0);
    if (i >= HostPipe::NBUF) cudaStreamWaitEvent(hp.s_k, hp.ev_out[b], 0);  // D2H i-NBUF has read d_out[b]
    LaunchArgs a{};
    a.mode = job.mode;
    a.X = job.in_on_device ? src_in : static_cast<const float*>(hp.d_in[b]);
    a.cond = C > 0 ? (job.in_on_device ? src_cond : static_cast<const float*>(hp.d_cond[b])) : nullptr;
    a.grid = static_cast<const float*>(hp.d_grid);
    a.G = job.G;
    a.col_idx = job.col_idx;
    a.N = nu * rows_per_unit;
    a.out0 = static_cast<float*>(hp.d_out0[b]);
    a.out1 = job.out1 ? static_cast<float*>(hp.d_out1[b]) : nullptr;
    if (flagged != nullptr) {
      a.ovf_flag = hp.d_flag + b;
      cudaMemsetAsync(a.ovf_flag, 0, sizeof(unsigned int), hp.s_k);
    }
    if (job.n_flows > 0) {
      float* ens = static_cast<float*>(hp.d_ens[b]);
      const size_t per_flow = (size_t)nu * job.out0_w;
      for (int f = 0; f < job.n_flows && rc == PZB_OK; ++f) {
        LaunchArgs af = a;
        af.out0 = ens + f * per_flow;
        rc = launch_chain(job.flows[f], af, hp.s_k);
        // posterior: every flow's NaNs become 0 before the mean, as in the reference (flowEnsemble.py:317-333 -> flow.py:735-737)
        if (rc == PZB_OK && job.mode == MODE_POSTERIOR && job.nan_to_zero)
          rc = pzb_normalize_pdfs(af.out0, nu, job.G, a.grid, 0, 1, hp.s_k);
      }
      if (rc == PZB_OK)
        rc = job.mode == MODE_POSTERIOR
                 ? pzb_ensemble_posterior_mean(ens, job.n_flows, nu, job.G, a.grid, job.normalize, a.out0, hp.s_k)
                 : pzb_ensemble_logmeanexp(ens, job.n_flows, nu, a.out0, hp.s_k);
    } else {
      // a single flow's posterior: reduced and normalised inside the tensor-core kernel where it can be
      const bool fused = job.mode == MODE_POSTERIOR && posterior_fuses(plan, job.G, 0, 0, nu);
      if (fused) {
        a.fuse = 1;
        a.normalize = job.normalize;
        a.nan_to_zero = job.nan_to_zero;
      }
      rc = launch_chain(plan, a, hp.s_k);
      if (rc == PZB_OK && !fused && job.mode == MODE_POSTERIOR && (job.normalize || job.nan_to_zero))
        rc = pzb_normalize_pdfs(a.out0, nu, job.G, a.grid, job.normalize, job.nan_to_zero, hp.s_k);
    }
    if (rc != PZB_OK) break;
    cudaEventRecord(hp.ev_k[b], hp.s_k);
    cudaStreamWaitEvent(hp.s_out, hp.ev_k[b], 0);
    float* dst0 = stage_out ? static_cast<float*>(hp.h_out0[b]) : job.out0 + u0 * job.out0_w;  // (out_cols => staged)
    float* dst1 = stage_out ? static_cast<float*>(hp.h_out1[b]) : (job.out1 ? job.out1 + u0 : nullptr);
    ce = cudaMemcpyAsync(dst0, hp.d_out0[b], (size_t)nu * job.out0_w * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out);
    if (ce == cudaSuccess && job.out1)
      ce = cudaMemcpyAsync(dst1, hp.d_out1[b], (size_t)nu * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out);
    if (ce == cudaSuccess && flagged != nullptr && (size_t)i < hp.cap_flags) {
      ce = cudaMemcpyAsync(hp.h_flags + i, hp.d_flag + b, sizeof(unsigned int), cudaMemcpyDeviceToHost, hp.s_out);
      spans.push_back({u0, nu});
    }
    if (ce == cudaSuccess) ce = cudaEventRecord(hp.ev_out[b], hp.s_out);
    if (ce != cudaSuccess) { rc = cuda_fail(ce, "host pipeline D2H"); break; }
    pend_u0[b] = u0;
    pend_nu[b] = nu;
    // with chunk i queued behind it, move chunk i-1's staged results on (the GPU keeps working on chunk i meanwhile)
    if (stage_out && i >= 1) {
      ce = drain(i - 1);
      if (ce != cudaSuccess) { rc = cuda_fail(ce, "host pipeline staged results"); break; }
    }
    u0 += nu;
    cur = std::min(chunk, cur * 2);
  }
  if (stage_out && rc == PZB_OK && i >= 1) {
    ce = drain(i - 1);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "host pipeline staged results");
  }
  // drain everything, also on the error path: the ring must be idle when the next call starts
  cudaError_t e1 = cudaStreamSynchronize(hp.s_in), e2 = cudaStreamSynchronize(hp.s_k),
              e3 = cudaStreamSynchronize(hp.s_out);
  if (rc != PZB_OK) return rc;
  ce = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
  if (ce != cudaSuccess) return cuda_fail(ce, who);
  if (flagged != nullptr)
    for (size_t k = 0; k < spans.size(); ++k)
      if (hp.h_flags[k] != 0u) flagged->push_back(spans[k]);
  return PZB_OK;
}

int pzb_set_host_threads(int32_t n) {
  if (n < 0) return fail(PZB_ERR_INVALID, "set_host_threads: negative thread count");
  g_host_threads = n > 64 ? 64 : n;
  return PZB_OK;
}

// Content check of a parameter pytree's leaves (the Python mirror's plan cache): per leaf the wrapping sum of its
// 8-byte words plus the sum of its tail bytes, all leaves in one call (the 7-D flow's 42 arrays, 1.2 MB: a few tens of
// microseconds instead of ~130 us of NumPy reductions per Flow call).  No device work.
int pzb_host_digest(const void* const* leaves, const int64_t* nbytes, int32_t n, uint64_t* sums) {
  if (n < 0 || (n > 0 && (!leaves || !nbytes || !sums))) return fail(PZB_ERR_INVALID, "host_digest: NULL argument");
  // one thread: a megabyte of cache-resident weights sums in tens of microseconds, less than waking a thread team costs
  for (int i = 0; i < n; ++i) {
    const unsigned char* p = static_cast<const unsigned char*>(leaves[i]);
    const int64_t nb = nbytes[i], n8 = nb / 8;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t k = 0;
    for (; k + 4 <= n8; k += 4) {
      uint64_t w[4];
      memcpy(w, p + 8 * k, 32);
      a0 += w[0];
      a1 += w[1];
      a2 += w[2];
      a3 += w[3];
    }
    uint64_t acc = (a0 + a1) + (a2 + a3);
    for (; k < n8; ++k) {
      uint64_t w;
      memcpy(&w, p + 8 * k, 8);
      acc += w;
    }
    uint64_t tail = 0;
    for (int64_t k = 8 * n8; k < nb; ++k) tail += p[k];
    sums[2 * i] = acc;
    sums[2 * i + 1] = tail;
  }
  return PZB_OK;
}

int pzb_log_prob_host(const pzb_plan* plan, const float* X, const float* cond, int64_t N, float* out) {
  int rc = check_common(plan, X, cond, N, out, "log_prob_host");
  if (rc) return rc;
  HostJob j;
  j.mode = MODE_LOGPROB;
  j.in = X;
  j.in_w = plan->host.D;
  j.cond = cond;
  j.out0 = out;
  j.out0_w = 1;
  j.units = N;
  return run_host(plan, j, "log_prob_host");
}

static int log_prob_columns(const pzb_plan* plan, const void* const* data_cols, const void* const* cond_cols,
                            bool f32, const float* cond_means, const float* cond_stds, int64_t N, float* out) {
  if (!plan) return fail(PZB_ERR_INVALID, "log_prob_host_columns: plan is NULL");
  if (N < 0) return fail(PZB_ERR_INVALID, "log_prob_host_columns: negative row count");
  const int D = plan->host.D, C = plan->host.C;
  if (N > 0 && (data_cols == nullptr || out == nullptr)) return fail(PZB_ERR_INVALID, "log_prob_host_columns: NULL buffer");
  if (N > 0 && C > 0 && (cond_cols == nullptr || cond_means == nullptr || cond_stds == nullptr))
    return fail(PZB_ERR_INVALID, "log_prob_host_columns: this plan is conditional (C = %d) but conditions are NULL", C);
  for (int j = 0; N > 0 && j < D; ++j)
    if (data_cols[j] == nullptr) return fail(PZB_ERR_INVALID, "log_prob_host_columns: data column %d is NULL", j);
  for (int j = 0; N > 0 && j < C; ++j)
    if (cond_cols[j] == nullptr) return fail(PZB_ERR_INVALID, "log_prob_host_columns: condition column %d is NULL", j);
  HostJob j;
  j.mode = MODE_LOGPROB;
  j.in_cols = data_cols;
  j.cols_f32 = f32;
  j.in_w = D;
  j.cond_cols = C > 0 ? cond_cols : nullptr;
  j.cond_mean = cond_means;
  j.cond_std = cond_stds;
  j.out0 = out;
  j.out0_w = 1;
  j.units = N;
  return run_host(plan, j, "log_prob_host_columns");
}

int pzb_log_prob_host_columns(const pzb_plan* plan, const double* const* data_cols, const double* const* cond_cols,
                              const float* cond_means, const float* cond_stds, int64_t N, float* out) {
  return log_prob_columns(plan, reinterpret_cast<const void* const*>(data_cols),
                          reinterpret_cast<const void* const*>(cond_cols), false, cond_means, cond_stds, N, out);
}

int pzb_log_prob_host_columns_f32(const pzb_plan* plan, const float* const* data_cols, const float* const* cond_cols,
                                  const float* cond_means, const float* cond_stds, int64_t N, float* out) {
  return log_prob_columns(plan, reinterpret_cast<const void* const*>(data_cols),
                          reinterpret_cast<const void* const*>(cond_cols), true, cond_means, cond_stds, N, out);
}

int pzb_ensemble_log_prob_host_columns(const pzb_plan* const* flows, int32_t n_flows, const void* const* data_cols,
                                       int32_t cols_are_f32, const void* const* cond_cols, const float* cond_means,
                                       const float* cond_stds, int64_t N, float* out) {
  if (!flows || n_flows < 1) return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: no flows");
  for (int f = 0; f < n_flows; ++f) {
    if (!flows[f]) return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: flow %d is NULL", f);
    if (flows[f]->device != flows[0]->device || flows[f]->host.D != flows[0]->host.D || flows[f]->host.C != flows[0]->host.C)
      return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: flow %d differs from flow 0 in device or shape", f);
    if (!flows[f]->latent_ready)
      return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: flow %d needs pzb_plan_set_latent first", f);
  }
  const pzb_plan* plan = flows[0];
  if (N < 0) return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: negative row count");
  const int D = plan->host.D, C = plan->host.C;
  if (N > 0 && (data_cols == nullptr || out == nullptr))
    return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: NULL buffer");
  if (N > 0 && C > 0 && (cond_cols == nullptr || cond_means == nullptr || cond_stds == nullptr))
    return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: the flows are conditional (C = %d) but conditions are NULL", C);
  for (int j = 0; N > 0 && j < D; ++j)
    if (data_cols[j] == nullptr) return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: data column %d is NULL", j);
  for (int j = 0; N > 0 && j < C; ++j)
    if (cond_cols[j] == nullptr) return fail(PZB_ERR_INVALID, "ensemble_log_prob_host_columns: condition column %d is NULL", j);
  HostJob j;
  j.mode = MODE_LOGPROB;
  j.in_cols = data_cols;
  j.cols_f32 = cols_are_f32 != 0;
  j.in_w = D;
  j.cond_cols = C > 0 ? cond_cols : nullptr;
  j.cond_mean = cond_means;
  j.cond_std = cond_stds;
  j.out0 = out;
  j.out0_w = 1;
  j.units = N;
  j.flows = flows;
  j.n_flows = n_flows;
  return run_host(plan, j, "ensemble_log_prob_host_columns");
}

int pzb_ensemble_posterior_host(const pzb_plan* const* flows, int32_t n_flows, const float* rest, const float* cond,
                                int64_t Ng, int32_t col_idx, const float* grid, int32_t G, int32_t normalize,
                                int32_t nan_to_zero, float* pdfs) {
  if (!flows || n_flows < 1) return fail(PZB_ERR_INVALID, "ensemble_posterior_host: no flows");
  for (int f = 0; f < n_flows; ++f) {
    if (!flows[f]) return fail(PZB_ERR_INVALID, "ensemble_posterior_host: flow %d is NULL", f);
    if (flows[f]->device != flows[0]->device || flows[f]->host.D != flows[0]->host.D || flows[f]->host.C != flows[0]->host.C)
      return fail(PZB_ERR_INVALID, "ensemble_posterior_host: flow %d differs from flow 0 in device or shape", f);
    if (!flows[f]->latent_ready)
      return fail(PZB_ERR_INVALID, "ensemble_posterior_host: flow %d needs pzb_plan_set_latent first", f);
  }
  const pzb_plan* plan = flows[0];
  int rc = check_common(plan, plan->host.D > 1 ? (const void*)rest : (const void*)grid, cond, Ng, pdfs,
                        "ensemble_posterior_host");
  if (rc) return rc;
  if (G < 1 || grid == nullptr) return fail(PZB_ERR_INVALID, "ensemble_posterior_host: empty grid");
  if (col_idx < 0 || col_idx >= plan->host.D)
    return fail(PZB_ERR_INVALID, "ensemble_posterior_host: column index %d out of range", col_idx);
  HostJob j;
  j.mode = MODE_POSTERIOR;
  j.in = rest;
  j.in_w = plan->host.D - 1;
  j.cond = cond;
  j.out0 = pdfs;
  j.out0_w = G;
  j.grid = grid;
  j.G = G;
  j.col_idx = col_idx;
  j.normalize = normalize;
  j.nan_to_zero = nan_to_zero;
  j.units = Ng;
  j.flows = flows;
  j.n_flows = n_flows;
  return run_host(plan, j, "ensemble_posterior_host");
}

int pzb_forward_host(const pzb_plan* plan, const float* X, const float* cond, int64_t N, float* U, float* log_det) {
  int rc = check_common(plan, X, cond, N, U, "forward_host");
  if (rc) return rc;
  HostJob j;
  j.mode = MODE_FORWARD;
  j.in = X;
  j.in_w = plan->host.D;
  j.cond = cond;
  j.out0 = U;
  j.out0_w = plan->host.D;
  j.out1 = log_det;
  j.units = N;
  return run_host(plan, j, "forward_host");
}

int pzb_inverse_host(const pzb_plan* plan, const float* U, const float* cond, int64_t N, float* X, float* log_det) {
  int rc = check_common(plan, U, cond, N, X, "inverse_host");
  if (rc) return rc;
  HostJob j;
  j.mode = MODE_INVERSE;
  j.in = U;
  j.in_w = plan->host.D;
  j.cond = cond;
  j.out0 = X;
  j.out0_w = plan->host.D;
  j.out1 = log_det;
  j.units = N;
  return run_host(plan, j, "inverse_host");
}

int pzb_inverse_to_host(const pzb_plan* plan, const float* U_dev, const float* cond_dev, int64_t N, float* X,
                        float* log_det, void* after_stream) {
  int rc = check_common(plan, U_dev, cond_dev, N, X, "inverse_to_host");
  if (rc) return rc;
  // the draws were produced on `after_stream`: they must be complete before the pipeline's own streams read them
  if (N > 0) {
    DeviceGuard guard(plan->device);
    cudaError_t ce = cudaStreamSynchronize((cudaStream_t)after_stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "inverse_to_host: producer stream");
  }
  HostJob j;
  j.mode = MODE_INVERSE;
  j.in = U_dev;
  j.in_on_device = true;
  j.in_w = plan->host.D;
  j.cond = cond_dev;
  j.out0 = X;
  j.out0_w = plan->host.D;
  j.out1 = log_det;
  j.units = N;
  return run_host(plan, j, "inverse_to_host");
}

int pzb_inverse_to_host_columns(const pzb_plan* plan, const float* U_dev, const float* cond_dev, int64_t N,
                                float* const* X_cols, void* after_stream) {
  int rc = check_common(plan, U_dev, cond_dev, N, X_cols, "inverse_to_host_columns");
  if (rc) return rc;
  for (int j = 0; N > 0 && j < plan->host.D; ++j)
    if (X_cols[j] == nullptr) return fail(PZB_ERR_INVALID, "inverse_to_host_columns: column %d is NULL", j);
  if (N > 0) {
    DeviceGuard guard(plan->device);
    cudaError_t ce = cudaStreamSynchronize((cudaStream_t)after_stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "inverse_to_host_columns: producer stream");
  }
  HostJob j;
  j.mode = MODE_INVERSE;
  j.in = U_dev;
  j.in_on_device = true;
  j.in_w = plan->host.D;
  j.cond = cond_dev;
  j.out_cols = X_cols;
  j.out0_w = plan->host.D;
  j.units = N;
  return run_host(plan, j, "inverse_to_host_columns");
}

int pzb_posterior_host(const pzb_plan* plan, const float* rest, const float* cond, int64_t Ng, int32_t col_idx,
                       const float* grid, int32_t G, int32_t normalize, int32_t nan_to_zero, float* pdfs) {
  int rc = check_common(plan, plan && plan->host.D > 1 ? (const void*)rest : (const void*)grid, cond, Ng, pdfs,
                        "posterior_host");
  if (rc) return rc;
  if (G < 1 || grid == nullptr) return fail(PZB_ERR_INVALID, "posterior_host: empty grid");
  if (col_idx < 0 || col_idx >= plan->host.D)
    return fail(PZB_ERR_INVALID, "posterior_host: column index %d out of range", col_idx);
  HostJob j;
  j.mode = MODE_POSTERIOR;
  j.in = rest;
  j.in_w = plan->host.D - 1;
  j.cond = cond;
  j.out0 = pdfs;
  j.out0_w = G;
  j.grid = grid;
  j.G = G;
  j.col_idx = col_idx;
  j.normalize = normalize;
  j.nan_to_zero = nan_to_zero;
  j.units = Ng;
  return run_host(plan, j, "posterior_host");
}

}  // extern "C"
