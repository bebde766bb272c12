// Sampling-loss kernels: forward-only scoring, fused forward+backward, fused refinement step.
//
// One kernel template does all three.  Work decomposition (B200: 148 SMs):
//   grid.y  = pose blocks (<= PCL_MAX_POSE_BLOCK poses, their R|t live in shared memory)
//   grid.x  = balanced contiguous ranges of rows (row = 256 consecutive points); the whole grid is a whole number
//             of resident waves (3 CTAs/SM forward, 2 forward+backward), so there is no tail wave
//   thread  = 4 or 5 points held in registers (SoA, coalesced 32-bit loads; lane i <-> point i, so with a
//             Morton-ordered cloud the 32 lanes of a warp gather texels from one small image patch)
//   per (row group, pose): 4-5 evaluations per thread, halving-butterfly warp reduction, a few lanes add into the
//             warp's private fp64 shared-memory row (no atomics on the hot loop)
//   per CTA: one fp64 partial record per pose to global memory; the LAST CTA of a pose block (ticket counter)
//             reduces the records in a fixed two-level order (deterministic) and finishes:
//             loss (score) | loss + 6-DoF gradient | loss + gradient + Adam + plateau + clamp (refine)
//   refinement iterations are launched with programmatic dependent launch: the next iteration's prologue overlaps
//             the previous iteration's tail.
//
// Replaces: utils.py:484-499 (grid scoring), omniloc.py:171-202 / :311-356 (+ autograd backward),
//           omniloc.py:44-58 / :249-269 (optimiser step, scheduler step, clamp).
#include "pcl_common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <mutex>

// ------------------------------------------------------------------------------------------------
// error / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
std::atomic<long long> g_pcl_launches{0};

void pcl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* pcl_last_error(void) { return g_err; }

cudaError_t pcl_pool_alloc(void** p, size_t bytes, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};      // bit d set: device d's pool threshold raised
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && !((configured.load() >> dev) & 1ull)) {
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    unsigned long long thr = ~0ull;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (e != cudaSuccess) return e;
    configured.fetch_or(1ull << dev);
  }
  return cudaMallocAsync(p, bytes ? bytes : 1, st);
}
void pcl_pool_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}
extern "C" int pcl_abi_version(void) { return PCL_ABI_VERSION; }
extern "C" int64_t pcl_launch_count(void) { return (int64_t)g_pcl_launches.load(); }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
// Warp reduction of NS (2 or 8) values with a halving butterfly: at each of the first log2(NS) steps a
// lane keeps half of its values and ships the other half to its xor-partner, so NS values cost
// NS-1 + (5 - log2 NS) shuffles instead of 5·NS.  On return lane L (L % 4 == 0 suffices) holds the full warp
// sum of value index pcl_butterfly_index<NS>(L) in v[0].
template <int NS>
__device__ __forceinline__ void pcl_warp_reduce(float (&v)[NS], const int lane) {
  int n = NS;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (n > 1) {
      const bool up = (lane & o) != 0;
      n >>= 1;
#pragma unroll
      for (int i = 0; i < NS / 2; ++i) {
        if (i < n) {
          const float send = up ? v[i] : v[i + n];
          const float keep = up ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    }
  }
}
template <int NS>
__device__ __forceinline__ int pcl_butterfly_index(const int lane) {
  return NS == 8 ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)) : ((lane >> 4) & 1);
}

// torch.optim.Adam (single-tensor path, betas 0.9/0.999, eps 1e-8) + ReduceLROnPlateau(mode=min,
// threshold=1e-4 rel, cooldown 0, min_lr 0, eps 1e-8) + translation clamp; one thread per candidate.
// fp32 tensors, fp64 python scalars — exactly the split torch has (omniloc.py:33,37,49-58).
__device__ void pcl_refine_update(PclRefineState& st, float* evalp, const float* g, float loss, const PclFinalize& fin) {
  st.last_loss = loss;
  st.step += 1;
  // bias corrections 1 - beta^step are the same for every candidate: computed on the host in fp64 (libm pow, as
  // python's `beta ** step`) and passed with the launch
  const float step_size = (float)(st.lr / fin.bc1);
  const float bc2_sqrt = (float)fin.bc2_sqrt;
  const float w1 = (float)(1.0 - 0.9), b2 = 0.999f, w2 = (float)(1.0 - 0.999);
  float newp[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float gi = g[i];
    const float m = st.m[i] + w1 * (gi - st.m[i]);              // exp_avg.lerp_(grad, 1-beta1)
    const float v = st.v[i] * b2 + (w2 * gi) * gi;              // mul_(beta2).addcmul_(g, g, 1-beta2)
    st.m[i] = m; st.v[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + 1e-8f;
    newp[i] = st.param[i] - step_size * (m / denom);            // addcdiv_(m, denom, value=-step_size)
  }
  // scheduler.step(loss)
  const double cur = (double)loss;
  if (cur < st.best * (1.0 - 1e-4)) { st.best = cur; st.bad = 0; } else { st.bad += 1; }
  if (st.bad > fin.patience) {
    const double new_lr = fmax(st.lr * fin.factor, 0.0);
    if (st.lr - new_lr > 1e-8) st.lr = new_lr;
    st.bad = 0;
  }
  // clamp translation into the quantile box; batch semantics evaluates the pre-clamp copy next
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float c = newp[i];
    if (i < 3) c = fminf(fmaxf(c, __ldg(fin.box + i)), __ldg(fin.box + 3 + i));
    st.param[i] = c;
    evalp[i] = fin.batch_semantics ? newp[i] : c;
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// KK rows of 256 consecutive points (row0 .. row0+KK-1) against every pose of the CTA's pose block.
template <int FMT, bool BWD, int KK, int NS, bool CHECK>
__device__ __forceinline__ void pcl_process_rows(const PclCloudView& C, const PclImage& I, const PclPose* s_pose, const int np,
                                                 double (*s_acc)[PCL_MAX_POSE_BLOCK][NS], const long long row0,
                                                 const int tid, const int lane, const int warp) {
  const long long base = row0 * PCL_THREADS + tid;
  float px[KK], py[KK], pz[KK], cr[KK], cg[KK], cb[KK];
#pragma unroll
  for (int j = 0; j < KK; ++j) {
    const long long i = base + (long long)j * PCL_THREADS;       // arrays are padded: always in bounds
    px[j] = __ldg(C.x + i); py[j] = __ldg(C.y + i); pz[j] = __ldg(C.z + i);
    cr[j] = __ldg(C.r + i); cg[j] = __ldg(C.g + i); cb[j] = __ldg(C.b + i);
  }
  for (int p = 0; p < np; ++p) {
    const PclPose pose = s_pose[p];
    PclAcc acc = {-0.f, -0.f, -0.f, -0.f, -0.f, -0.f, -0.f, -0.f};   // -0 + x == x: first adds fold away
#pragma unroll
    for (int j = 0; j < KK; ++j) {
      const bool valid = CHECK ? ((base + (long long)j * PCL_THREADS) < C.n) : true;   // only the cloud's last row can be ragged
      pcl_eval<FMT, BWD>(pose, I, px[j], py[j], pz[j], cr[j], cg[j], cb[j], valid, acc);
    }
    float v[NS];
    v[0] = acc.se; v[1] = acc.sm;
    if (BWD) { v[2] = acc.ax; v[3] = acc.ay; v[4] = acc.az; v[5] = acc.tx; v[6] = acc.ty; v[7] = acc.tz; }
    pcl_warp_reduce<NS>(v, lane);
    if ((lane & (NS == 8 ? 3 : 15)) == 0) s_acc[warp][p][pcl_butterfly_index<NS>(lane)] += (double)v[0];
  }
}

// Register budgets: forward 3 CTAs/SM x 80 regs, forward+backward 2 x 128 (one more CTA per SM was measured
// slower: the lost ILP costs more than the extra warps hide).  Row groups of 4 and 5 rows (K points per thread in
// registers) tile any range of >= 12 rows exactly, so no CTA falls back to the low-ILP single-row path.
template <int FMT, bool BWD>
__global__ void __launch_bounds__(PCL_THREADS, BWD ? 2 : 3)
pcl_sample_kernel(const PclCloudView C, const PclImage I, const float* poses6, const int P, const int PB,
                  const long long n_rows, double* __restrict__ partial, unsigned int* __restrict__ counters, const PclFinalize fin,
                  const int swap) {
  constexpr int NS = BWD ? PCL_NSUM : 2;
  __shared__ __align__(16) PclPose s_pose[PCL_MAX_POSE_BLOCK];
  __shared__ double s_acc[PCL_WARPS][PCL_MAX_POSE_BLOCK][NS];   // fp64: row-to-row accumulation adds no fp32 error
  __shared__ double2 s_sum[PCL_THREADS];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // block order: pose blocks fastest (swap) keeps the co-resident CTAs on few row ranges
  const unsigned int bp = swap ? blockIdx.x : blockIdx.y, br = swap ? blockIdx.y : blockIdx.x;
  const unsigned int n_ranges = swap ? gridDim.y : gridDim.x;
  const int p0 = bp * PB;
  const int np = min(PB, P - p0);

  for (int i = tid; i < PCL_WARPS * PCL_MAX_POSE_BLOCK * NS; i += PCL_THREADS) (&s_acc[0][0][0])[i] = 0.0;
#if __CUDA_ARCH__ >= 900
  // Programmatic dependent launch: this grid may start while the previous launch on the stream (the previous
  // refinement iteration) is still draining; everything above overlaps with its tail.  The poses written by its
  // finishing CTA are only read after this wait (no-op when the launch did not opt in).
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  if (tid < np) pcl_pose_from_params(poses6 + 6 * (size_t)(p0 + tid), s_pose[tid]);
  __syncthreads();

  // balanced contiguous row range of this CTA (sizes differ by at most one row of 256 points)
  const long long r_begin = n_rows * (long long)br / (long long)n_ranges;
  const long long r_end = n_rows * (long long)(br + 1) / (long long)n_ranges;
  long long r = r_begin;
  const long long r_full = min(r_end, C.n / PCL_THREADS);      // rows below r_full have 256 real points
  {
    // n = 4a + b rows (b < 4): b groups of 5 and a-b groups of 4 when a >= b (always for n >= 12)
    const long long n = r_full - r, a = n >> 2, b = n & 3;
    long long n5 = (a >= b) ? b : 0, n4 = (a >= b) ? a - b : a;
    for (; n5 > 0; --n5, r += 5) pcl_process_rows<FMT, BWD, 5, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
    for (; n4 > 0; --n4, r += 4) pcl_process_rows<FMT, BWD, 4, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
  }
  for (; r < r_full; ++r) pcl_process_rows<FMT, BWD, 1, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
  for (; r < r_end; ++r) pcl_process_rows<FMT, BWD, 1, NS, true>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
  __syncthreads();

  // CTA partial record: partial[blockIdx.x][pose][s]  (NS contiguous doubles per pose)
  const int nout = np * NS;
  for (int i = tid; i < nout; i += PCL_THREADS) {
    const int p = i / NS, s = i - p * NS;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PCL_WARPS; ++w) t += s_acc[w][p][s];
    partial[((size_t)br * (size_t)P + (size_t)(p0 + p)) * NS + s] = t;
  }

  // last-block-done
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&counters[bp], 1u);
    s_last = (ticket == n_ranges - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // Deterministic two-level reduction of the gridDim.x partial records, latency-optimised: an item is one
  // 16-byte pair of sums of one pose; G thread groups stride the records with 128-bit L2 loads (16 in flight per
  // thread), then the groups are summed in fixed order.
  const int nitems = np * (NS / 2);
  const int G = max(1, PCL_THREADS / nitems);
  {
    const int item = tid % nitems, g = tid / nitems;
    double2 t = make_double2(0.0, 0.0);
    if (g < G) {
      const double2* src = reinterpret_cast<const double2*>(partial + (size_t)p0 * NS) + item;
      const size_t stride = (size_t)P * (NS / 2);
#pragma unroll 16
      for (unsigned int bx = g; bx < n_ranges; bx += G) {
        const double2 v = __ldcg(src + (size_t)bx * stride);
        t.x += v.x; t.y += v.y;
      }
    }
    s_sum[tid] = t;
  }
  __syncthreads();
  // second level: one thread per item sums its G group values (same order as before), then the pose threads pick them up
  double2 fin_item = make_double2(0.0, 0.0);
  if (tid < nitems) {
    for (int g = 0; g < G; ++g) { const double2 v = s_sum[g * nitems + tid]; fin_item.x += v.x; fin_item.y += v.y; }
  }
  __syncthreads();
  if (tid < nitems) s_sum[tid] = fin_item;
  __syncthreads();
  if (tid < np) {
    double sums[PCL_NSUM];
#pragma unroll
    for (int h = 0; h < NS / 2; ++h) {
      const double2 t = s_sum[tid * (NS / 2) + h];
      sums[2 * h] = t.x; sums[2 * h + 1] = t.y;
    }
    const int pg = p0 + tid;
    const float* p6 = poses6 + 6 * (size_t)pg;
    float loss, cnt, grad[6];
    pcl_finish_gradient(p6, s_pose[tid], I, sums, &loss, &cnt, BWD ? grad : nullptr);
    if (fin.loss) fin.loss[pg] = loss;
    if (fin.count) fin.count[pg] = cnt;
    if (BWD) {
      if (fin.mode == PCL_FIN_GRAD) {
#pragma unroll
        for (int i = 0; i < 6; ++i) fin.grad[6 * (size_t)pg + i] = grad[i];
      } else if (fin.mode == PCL_FIN_REFINE) {
        pcl_refine_update(fin.state[pg], fin.evalp + 6 * (size_t)pg, grad, loss, fin);
      }
    }
  }
  if (tid == 0) counters[bp] = 0u;     // self-resetting: the next launch needs no memset
}

// ------------------------------------------------------------------------------------------------
// host-side launch
// ------------------------------------------------------------------------------------------------
struct PclLaunchPlan { int PB, gx, gy, NS; long long n_rows; };

// Tuning knobs (A/B experiments and tests): process-global, read ONCE from the environment (PCL_<NAME>) on first use,
// changed afterwards only through pcl_set_option — no getenv on the launch path.
static const char* const g_opt_names[PCL_OPT_COUNT] = {"PERSIST", "PDL", "PB_FWD", "PB_BWD", "WAVES", "SWAP", "GRID_SWAP", "SMALL_TABLE", "RF_NPB", "RF_DEBUG", "RF_RES"};
static const int g_opt_defaults[PCL_OPT_COUNT] = {1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 1};
static std::atomic<int> g_opt[PCL_OPT_COUNT];
static std::once_flag g_opt_once;

static void pcl_opt_init() {
  for (int i = 0; i < PCL_OPT_COUNT; ++i) {
    char name[64];
    snprintf(name, sizeof(name), "PCL_%s", g_opt_names[i]);
    const char* v = getenv(name);
    g_opt[i].store(v ? atoi(v) : g_opt_defaults[i]);
  }
}
int pcl_opt(int id) {
  std::call_once(g_opt_once, pcl_opt_init);
  return g_opt[id].load(std::memory_order_relaxed);
}
extern "C" int pcl_set_option(const char* name, int value) {
  std::call_once(g_opt_once, pcl_opt_init);
  if (!name) { pcl_set_error("null option name"); return PCL_ERR_INVALID; }
  for (int i = 0; i < PCL_OPT_COUNT; ++i) {
    if (strcasecmp(name, g_opt_names[i]) == 0) { g_opt[i].store(value < 0 ? g_opt_defaults[i] : value); return PCL_OK; }
  }
  pcl_set_error("unknown option %s", name);
  return PCL_ERR_INVALID;
}

int pcl_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

// Grid sizing (148 SMs): the whole grid is ONE resident wave — gx*gy = SMs x resident CTAs per SM (3 forward,
// 2 forward+backward, from the register budgets) — unless there are more pose blocks than that; every CTA
// gets an equal share of rows, so there is no tail wave and the last-block reduction reads <= ~450 rows.
static PclLaunchPlan pcl_plan(const pcl_cloud* c, int64_t P, bool bwd) {
  PclLaunchPlan pl;
  pl.NS = bwd ? PCL_NSUM : 2;
  // rows that hold real points (the arrays are padded further, to PCL_TILE_ALIGN): padding rows are never scheduled —
  // they used to land on the last CTA's single-row path and made it the straggler every launch waits for
  pl.n_rows = (c->n + PCL_THREADS - 1) / PCL_THREADS;
  const int resident = pcl_num_sms() * (bwd ? 2 : 3);
  // pose block: as many poses per CTA as possible (point loads amortise over the block) while leaving
  // enough CTAs to fill the machine
  int PB = (int)(P < PCL_MAX_POSE_BLOCK ? P : PCL_MAX_POSE_BLOCK);
  // small refinement batches: two pose blocks (twice the rows per CTA, half the row-count imbalance) measured
  // 8 % faster than one block of all candidates (B=6: 47 vs 51 us per iteration)
  if (bwd && P >= 4 && P <= 16) PB = (int)((P + 1) / 2);
  const int pb_env = pcl_opt(bwd ? PCL_OPT_PB_BWD : PCL_OPT_PB_FWD);
  if (pb_env > 0 && pb_env <= PCL_MAX_POSE_BLOCK) PB = (int)(pb_env < P ? pb_env : P);
  pl.PB = PB;
  pl.gy = (int)((P + PB - 1) / PB);
  // 1..8 full waves: take the (smallest) wave count whose grid fills its slots best (CTAs do equal work).  Eight, not four:
  // with 256 pose blocks (P = 8192 forward+backward, 296 slots) every grid of up to 4 waves leaves 13.5 % of the slots empty,
  // the 7-wave grid 1.2 %
  long long gx = 1;
  double best = -1.0;
  const int w_env = pcl_opt(PCL_OPT_WAVES);
  for (int w = (w_env > 0 ? w_env : 1); w <= (w_env > 0 ? w_env : 8); ++w) {
    long long g = (long long)resident * w / pl.gy;
    if (g < 1) g = 1;
    if (g > pl.n_rows) g = pl.n_rows;
    const double util = (double)(g * pl.gy) / (double)((g * pl.gy + resident - 1) / resident * resident);
    if (util > best + 1e-9) { best = util; gx = g; }
  }
  pl.gx = (int)gx;
  return pl;
}

template <int FMT, bool BWD>
static cudaError_t pcl_launch_fmt(const PclLaunchPlan& pl, const PclCloudView& C, const PclImage& I, const float* poses, int P,
                                  double* partial, unsigned int* counters, const PclFinalize& fin, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  const int swap = (!BWD && pl.gy > 1 && pl.gx <= 65535) ? pcl_opt(PCL_OPT_SWAP) : 0;
  cfg.gridDim = swap ? dim3(pl.gy, pl.gx) : dim3(pl.gx, pl.gy);
  cfg.blockDim = dim3(PCL_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, pcl_sample_kernel<FMT, BWD>, C, I, poses, P, pl.PB, pl.n_rows, partial, counters, fin, swap);
}

template <bool BWD>
static int pcl_launch(const PclLaunchPlan& pl, const pcl_cloud* c, const pcl_image* im, const float* poses, int P,
                      double* partial, unsigned int* counters, const PclFinalize& fin, cudaStream_t st, bool pdl = false) {
  PclCloudView C = {c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
  cudaError_t e;
  // small forward+backward batches (refinement) read the compact companion table when the image has one
  const PclImage& view = (BWD && P <= 16 && im->has_small && pcl_opt(PCL_OPT_SMALL_TABLE)) ? im->view_small : im->view;
  switch (view.fmt) {
    case PCL_FMT_U8Q: e = pcl_launch_fmt<PCL_FMT_U8Q, BWD>(pl, C, view, poses, P, partial, counters, fin, st, pdl); break;
    case PCL_FMT_U8P: e = pcl_launch_fmt<PCL_FMT_U8P, BWD>(pl, C, view, poses, P, partial, counters, fin, st, pdl); break;
    case PCL_FMT_F32: e = pcl_launch_fmt<PCL_FMT_F32, BWD>(pl, C, view, poses, P, partial, counters, fin, st, pdl); break;
    case PCL_FMT_TEX: e = pcl_launch_fmt<PCL_FMT_TEX, BWD>(pl, C, view, poses, P, partial, counters, fin, st, pdl); break;
    case PCL_FMT_F16D: e = pcl_launch_fmt<PCL_FMT_F16D, BWD>(pl, C, view, poses, P, partial, counters, fin, st, pdl); break;
    default: pcl_set_error("unknown image format %d", view.fmt); return PCL_ERR_INVALID;
  }
  g_pcl_launches.fetch_add(1);
  PCL_CUDA(e);
  return PCL_OK;
}

static int pcl_check_inputs(const pcl_cloud* c, const pcl_image* im, const void* poses, int64_t P) {
  if (!c || !im || !poses) { pcl_set_error("null handle or pose pointer"); return PCL_ERR_INVALID; }
  if (P <= 0 || P > 65535ll * PCL_MAX_POSE_BLOCK) { pcl_set_error("pose count %lld out of range (1 .. %lld)", (long long)P, 65535ll * PCL_MAX_POSE_BLOCK); return PCL_ERR_INVALID; }
  return PCL_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: scoring, loss+gradient
// ------------------------------------------------------------------------------------------------
static int pcl_run_once(const pcl_cloud* c, const pcl_image* im, const float* poses, int64_t P, bool bwd,
                        float* loss, float* count, float* grad, cudaStream_t st) {
  const PclLaunchPlan pl = pcl_plan(c, P, bwd);
  double* partial = nullptr;
  unsigned int* counters = nullptr;
  const size_t pbytes = (size_t)pl.gx * pl.NS * (size_t)P * sizeof(double);
  PCL_CUDA(pcl_pool_alloc((void**)&partial, pbytes + (size_t)pl.gy * sizeof(unsigned int), st));
  counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(partial) + pbytes);
  PCL_CUDA(cudaMemsetAsync(counters, 0, (size_t)pl.gy * sizeof(unsigned int), st));
  PclFinalize fin;
  memset(&fin, 0, sizeof(fin));
  fin.mode = bwd ? PCL_FIN_GRAD : PCL_FIN_SCORE;
  fin.loss = loss; fin.count = count; fin.grad = grad;
  int rc = bwd ? pcl_launch<true>(pl, c, im, poses, (int)P, partial, counters, fin, st)
               : pcl_launch<false>(pl, c, im, poses, (int)P, partial, counters, fin, st);
  pcl_pool_free(partial, st);
  return rc;
}

extern "C" int pcl_score(const pcl_cloud* c, const pcl_image* im, const float* poses_p6_dev, int64_t p,
                         float* loss_p_dev, float* count_p_dev, void* stream) {
  int rc = pcl_check_inputs(c, im, poses_p6_dev, p);
  if (rc) return rc;
  if (!loss_p_dev) { pcl_set_error("loss output is null"); return PCL_ERR_INVALID; }
  PclUseGuard guard{c, im, (cudaStream_t)stream};
  return pcl_run_once(c, im, poses_p6_dev, p, false, loss_p_dev, count_p_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int pcl_loss_fwd_bwd(const pcl_cloud* c, const pcl_image* im, const float* poses_b6_dev, int b,
                                float* loss_b_dev, float* count_b_dev, float* grad_b6_dev, void* stream) {
  int rc = pcl_check_inputs(c, im, poses_b6_dev, b);
  if (rc) return rc;
  if (!loss_b_dev || !grad_b6_dev) { pcl_set_error("loss/grad output is null"); return PCL_ERR_INVALID; }
  PclUseGuard guard{c, im, (cudaStream_t)stream};
  return pcl_run_once(c, im, poses_b6_dev, b, true, loss_b_dev, count_b_dev, grad_b6_dev, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// refinement of LARGE batches (B > 16; pcl_refine.cu handles the small ones): one launch per iteration for the
// whole candidate batch, the finishing CTA of each pose block steps Adam / plateau / clamp (PCL_FIN_REFINE)
// ------------------------------------------------------------------------------------------------
int pcl_generic_refine_iters(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, cudaStream_t st) {
  int rc = pcl_check_inputs(c, im, r->evalp, r->B);
  if (rc) return rc;
  const PclLaunchPlan pl = pcl_plan(c, r->B, true);
  const size_t need = (size_t)pl.gx * pl.NS * (size_t)r->B;
  if (need > r->partial_floats) {
    pcl_pool_free(r->partial, st);
    r->partial = nullptr; r->partial_floats = 0;
    PCL_CUDA(pcl_pool_alloc((void**)&r->partial, need * sizeof(double), st));
    r->partial_floats = need;
  }
  PclFinalize fin;
  memset(&fin, 0, sizeof(fin));
  fin.mode = PCL_FIN_REFINE;
  fin.loss = r->loss; fin.state = r->state; fin.evalp = r->evalp;
  fin.box = c->lo_hi_dev;
  fin.factor = r->factor; fin.patience = r->patience; fin.batch_semantics = r->batch_semantics;
  for (int it = 0; it < num_iter; ++it) {
    r->steps_done += 1;
    fin.bc1 = 1.0 - pow(0.9, (double)r->steps_done);
    fin.bc2_sqrt = sqrt(1.0 - pow(0.999, (double)r->steps_done));
    // iterations after the first opt in to programmatic dependent launch (prologue overlaps the previous tail)
    rc = pcl_launch<true>(pl, c, im, r->evalp, r->B, r->partial, r->counters, fin, st, it > 0 && pcl_opt(PCL_OPT_PDL) != 0);
    if (rc) return rc;
  }
  return PCL_OK;
}
