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
// persistent refinement: ALL iterations of a small candidate batch in one cooperative launch
// ------------------------------------------------------------------------------------------------
// The per-iteration launch above pays ~10 us of serial chain per iteration even on an empty cloud (fence ->
// ticket -> last-CTA reduction -> Adam -> grid drain -> dependent launch -> pose set-up; scripts/tail_probe.py),
// a fifth of a C2 iteration.  Here the grid (one resident wave, launched cooperatively) stays on the SMs: per
// iteration every CTA writes its partial record, the CTAs of a pose block meet at ONE barrier, and then EVERY CTA
// reduces the block's records and steps Adam / plateau / clamp for the block's candidates itself — redundantly, in
// the same fixed order, so all replicas are bit-identical and the next iteration's poses are already in the CTA's
// shared memory.  Partial records are double-buffered by iteration parity; one barrier per iteration suffices.
struct PclPersist {
  double* partial;              // [2][n_ranges][P][8]
  unsigned int* barrier;        // [pose blocks] monotonic arrival counters, zero at launch
  const double* bc;             // [num_iter][2] = {1 - 0.9^step, sqrt(1 - 0.999^step)} (host libm, as python's `beta ** step`)
  int num_iter;
};

__device__ __forceinline__ unsigned int pcl_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int FMT>
__global__ void __launch_bounds__(PCL_THREADS, 2)
pcl_refine_persistent_kernel(const PclCloudView C, const PclImage I, const int P, const int PB, const long long n_rows,
                             const PclFinalize fin, const PclPersist ps) {
  constexpr int NS = PCL_NSUM;
  constexpr int MAXB = 16;                                      // candidates per pose block (host guarantees PB <= 16)
  __shared__ __align__(16) PclPose s_pose[PCL_MAX_POSE_BLOCK];
  __shared__ double s_acc[PCL_WARPS][PCL_MAX_POSE_BLOCK][NS];
  __shared__ double2 s_sum[PCL_THREADS];
  __shared__ PclRefineState s_state[MAXB];
  __shared__ float s_evalp[MAXB][6];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int bp = blockIdx.y, br = blockIdx.x, n_ranges = gridDim.x;
  const int p0 = bp * PB;
  const int np = min(PB, P - p0);
  if (tid < np) {
    s_state[tid] = fin.state[p0 + tid];
#pragma unroll
    for (int i = 0; i < 6; ++i) s_evalp[tid][i] = fin.evalp[6 * (size_t)(p0 + tid) + i];
  }
  const long long r_begin = n_rows * (long long)br / (long long)n_ranges;
  const long long r_end = n_rows * (long long)(br + 1) / (long long)n_ranges;
  const long long r_full = min(r_end, C.n / PCL_THREADS);
  const int nitems = np * (NS / 2);
  const int G = max(1, PCL_THREADS / nitems);

  for (int it = 0; it < ps.num_iter; ++it) {
    __syncthreads();                                             // s_evalp of the previous update is visible; s_acc is free
    for (int i = tid; i < PCL_WARPS * np * NS; i += PCL_THREADS) {
      const int w = i / (np * NS), r = i - w * (np * NS);
      s_acc[w][r / NS][r % NS] = 0.0;
    }
    if (tid < np) pcl_pose_from_params(s_evalp[tid], s_pose[tid]);
    __syncthreads();

    long long r = r_begin;
    {
      const long long n = r_full - r, a = n >> 2, b = n & 3;
      long long n5 = (a >= b) ? b : 0, n4 = (a >= b) ? a - b : a;
      for (; n5 > 0; --n5, r += 5) pcl_process_rows<FMT, true, 5, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
      for (; n4 > 0; --n4, r += 4) pcl_process_rows<FMT, true, 4, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
    }
    for (; r < r_full; ++r) pcl_process_rows<FMT, true, 1, NS, false>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
    for (; r < r_end; ++r) pcl_process_rows<FMT, true, 1, NS, true>(C, I, s_pose, np, s_acc, r, tid, lane, warp);
    __syncthreads();

    double* part = ps.partial + (size_t)(it & 1) * (size_t)n_ranges * (size_t)P * NS;
    for (int i = tid; i < np * NS; i += PCL_THREADS) {
      const int p = i / NS, s = i - p * NS;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < PCL_WARPS; ++w) t += s_acc[w][p][s];
      part[((size_t)br * (size_t)P + (size_t)(p0 + p)) * NS + s] = t;
    }

    // barrier of the pose block's CTAs (all co-resident: cooperative launch)
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(ps.barrier + bp, 1u);
      const unsigned int target = (unsigned int)(it + 1) * n_ranges;
      while (pcl_ld_acquire(ps.barrier + bp) < target) { }
    }
    __syncthreads();

    // every CTA: deterministic two-level reduction of the block's records (as in pcl_sample_kernel), then the update
    {
      const int item = tid % nitems, g = tid / nitems;
      double2 t = make_double2(0.0, 0.0);
      if (g < G) {
        const double2* src = reinterpret_cast<const double2*>(part + (size_t)p0 * NS) + item;
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
      float loss, cnt, grad[6];
      pcl_finish_gradient(s_evalp[tid], s_pose[tid], I, sums, &loss, &cnt, grad);
      PclFinalize f = fin;
      f.bc1 = __ldg(ps.bc + 2 * it); f.bc2_sqrt = __ldg(ps.bc + 2 * it + 1);
      pcl_refine_update(s_state[tid], s_evalp[tid], grad, loss, f);
    }
  }
  __syncthreads();
  if (br == 0 && tid < np) {                                     // one replica publishes the end state
    const int pg = p0 + tid;
    fin.state[pg] = s_state[tid];
#pragma unroll
    for (int i = 0; i < 6; ++i) fin.evalp[6 * (size_t)pg + i] = s_evalp[tid][i];
    if (fin.loss) fin.loss[pg] = s_state[tid].last_loss;
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launch
// ------------------------------------------------------------------------------------------------
struct PclLaunchPlan { int PB, gx, gy, NS; long long n_rows; };

static int pcl_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static int pcl_num_sms() {
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
  const int pb_env = pcl_env_int(bwd ? "PCL_PB_BWD" : "PCL_PB_FWD", 0);
  if (pb_env > 0 && pb_env <= PCL_MAX_POSE_BLOCK) PB = (int)(pb_env < P ? pb_env : P);
  pl.PB = PB;
  pl.gy = (int)((P + PB - 1) / PB);
  // 1..4 full waves: take the wave count whose grid fills its slots best (CTAs do equal work)
  long long gx = 1;
  double best = -1.0;
  const int w_env = pcl_env_int("PCL_WAVES", 0);
  for (int w = (w_env > 0 ? w_env : 1); w <= (w_env > 0 ? w_env : 4); ++w) {
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
  const int swap = (!BWD && pl.gy > 1 && pl.gx <= 65535) ? pcl_env_int("PCL_SWAP", 1) : 0;
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
  const PclImage& view = (BWD && P <= 16 && im->has_small && pcl_env_int("PCL_SMALL_TABLE", 1)) ? im->view_small : im->view;
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
  return pcl_run_once(c, im, poses_p6_dev, p, false, loss_p_dev, count_p_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int pcl_loss_fwd_bwd(const pcl_cloud* c, const pcl_image* im, const float* poses_b6_dev, int b,
                                float* loss_b_dev, float* count_b_dev, float* grad_b6_dev, void* stream) {
  int rc = pcl_check_inputs(c, im, poses_b6_dev, b);
  if (rc) return rc;
  if (!loss_b_dev || !grad_b6_dev) { pcl_set_error("loss/grad output is null"); return PCL_ERR_INVALID; }
  return pcl_run_once(c, im, poses_b6_dev, b, true, loss_b_dev, count_b_dev, grad_b6_dev, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// C ABI: fused refinement
// ------------------------------------------------------------------------------------------------
__global__ void pcl_refine_reset_kernel(PclRefineState* st, float* evalp, const float* poses6, int B, double lr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  PclRefineState s;
  for (int i = 0; i < 6; ++i) { s.m[i] = 0.f; s.v[i] = 0.f; s.param[i] = poses6[6 * b + i]; evalp[6 * b + i] = poses6[6 * b + i]; }
  s.last_loss = nanf(""); s.step = 0; s.bad = 0; s.pad = 0; s.lr = lr; s.best = INFINITY;
  st[b] = s;
}

__global__ void pcl_refine_read_kernel(const PclRefineState* st, const float* evalp, int B, int batch, float* pose, float* param,
                                       float* loss, double* lr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < 6; ++i) {
    // sequential semantics returns the clamped parameter; batch semantics the pre-clamp copy (omniloc.py:260,272)
    if (pose) pose[6 * b + i] = batch ? evalp[6 * b + i] : st[b].param[i];
    if (param) param[6 * b + i] = st[b].param[i];
  }
  if (loss) loss[b] = st[b].last_loss;
  if (lr) lr[b] = st[b].lr;
}

extern "C" int pcl_refine_create(int b, double lr, double factor, int patience, int batch_semantics, pcl_refine** out) {
  if (!out || b <= 0 || b > 65536) { pcl_set_error("bad refine batch %d", b); return PCL_ERR_INVALID; }
  pcl_refine* r = (pcl_refine*)calloc(1, sizeof(pcl_refine));
  r->B = b; r->lr0 = lr; r->factor = factor; r->patience = patience; r->batch_semantics = batch_semantics ? 1 : 0;
  *out = r;                          // device storage is allocated by the first pcl_refine_reset, on its stream
  return PCL_OK;
}

extern "C" int pcl_refine_reset(pcl_refine* r, const float* poses_b6_dev, void* stream) {
  if (!r || !poses_b6_dev) { pcl_set_error("null refine handle or poses"); return PCL_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t b = (size_t)r->B;
  const size_t o_eval = (sizeof(PclRefineState) * b + 255) & ~(size_t)255;
  const size_t o_loss = o_eval + ((sizeof(float) * 6 * b + 255) & ~(size_t)255);
  const size_t o_cnt = o_loss + ((sizeof(float) * b + 255) & ~(size_t)255);
  if (!r->block) {
    PCL_CUDA(pcl_pool_alloc((void**)&r->block, o_cnt + sizeof(unsigned int) * b, st));
    r->owner = st;
    r->state = (PclRefineState*)r->block;
    r->evalp = (float*)(r->block + o_eval);
    r->loss = (float*)(r->block + o_loss);
    r->counters = (unsigned int*)(r->block + o_cnt);
  }
  PCL_CUDA(cudaMemsetAsync(r->counters, 0, sizeof(unsigned int) * b, st));
  r->steps_done = 0;
  pcl_refine_reset_kernel<<<(r->B + 127) / 128, 128, 0, st>>>(r->state, r->evalp, poses_b6_dev, r->B, r->lr0);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

template <int FMT>
static cudaError_t pcl_persist_launch_fmt(const PclLaunchPlan& pl, PclCloudView C, PclImage I, int P, PclFinalize fin, PclPersist ps, cudaStream_t st) {
  int PB = pl.PB;
  long long n_rows = pl.n_rows;
  void* args[] = {&C, &I, &P, &PB, &n_rows, &fin, &ps};
  return cudaLaunchCooperativeKernel((const void*)pcl_refine_persistent_kernel<FMT>, dim3(pl.gx, pl.gy), dim3(PCL_THREADS), args, 0, st);
}

// returns PCL_OK, an error, or 1 when the cooperative launch is not possible on this device / configuration
static int pcl_refine_run_persistent(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, const PclLaunchPlan& pl, const PclFinalize& fin,
                                     int num_iter, cudaStream_t st) {
  static int coop = -1;
  if (coop < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) coop = 0;
  }
  if (!coop) return 1;
  const size_t need = 2 * (size_t)pl.gx * pl.NS * (size_t)r->B;
  if (need > r->partial_floats) {
    pcl_pool_free(r->partial, st);
    r->partial = nullptr; r->partial_floats = 0;
    PCL_CUDA(pcl_pool_alloc((void**)&r->partial, need * sizeof(double), st));
    r->partial_floats = need;
  }
  if ((size_t)num_iter * 2 > r->bc_cap) {
    pcl_pool_free(r->bc_dev, st);
    r->bc_dev = nullptr; r->bc_cap = 0;
    PCL_CUDA(pcl_pool_alloc((void**)&r->bc_dev, (size_t)num_iter * 2 * sizeof(double), st));
    r->bc_cap = (size_t)num_iter * 2;
  }
  double* bc = (double*)malloc((size_t)num_iter * 2 * sizeof(double));
  if (!bc) { pcl_set_error("out of host memory"); return PCL_ERR_INVALID; }
  for (int it = 0; it < num_iter; ++it) {
    const double step = (double)(r->steps_done + it + 1);
    bc[2 * it] = 1.0 - pow(0.9, step);
    bc[2 * it + 1] = sqrt(1.0 - pow(0.999, step));
  }
  // pageable source: the call returns once the data is staged, the buffer can be released right away
  cudaError_t e = cudaMemcpyAsync(r->bc_dev, bc, (size_t)num_iter * 2 * sizeof(double), cudaMemcpyHostToDevice, st);
  free(bc);
  PCL_CUDA(e);
  PCL_CUDA(cudaMemsetAsync(r->counters, 0, sizeof(unsigned int) * (size_t)pl.gy, st));
  PclPersist ps = {r->partial, r->counters, r->bc_dev, num_iter};
  PclCloudView C = {c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
  const PclImage& view = (im->has_small && pcl_env_int("PCL_SMALL_TABLE", 1)) ? im->view_small : im->view;
  switch (view.fmt) {
    case PCL_FMT_U8Q: e = pcl_persist_launch_fmt<PCL_FMT_U8Q>(pl, C, view, r->B, fin, ps, st); break;
    case PCL_FMT_U8P: e = pcl_persist_launch_fmt<PCL_FMT_U8P>(pl, C, view, r->B, fin, ps, st); break;
    case PCL_FMT_F32: e = pcl_persist_launch_fmt<PCL_FMT_F32>(pl, C, view, r->B, fin, ps, st); break;
    case PCL_FMT_TEX: e = pcl_persist_launch_fmt<PCL_FMT_TEX>(pl, C, view, r->B, fin, ps, st); break;
    case PCL_FMT_F16D: e = pcl_persist_launch_fmt<PCL_FMT_F16D>(pl, C, view, r->B, fin, ps, st); break;
    default: pcl_set_error("unknown image format %d", view.fmt); return PCL_ERR_INVALID;
  }
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) { (void)cudaGetLastError(); return 1; }
  g_pcl_launches.fetch_add(1);
  PCL_CUDA(e);
  PCL_CUDA(cudaMemsetAsync(r->counters, 0, sizeof(unsigned int) * (size_t)pl.gy, st));     // tickets of the per-iteration path start at zero
  r->steps_done += num_iter;
  return PCL_OK;
}

extern "C" int pcl_refine_run(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, void* stream) {
  if (!r || !r->block) { pcl_set_error("refine handle is null or was never reset"); return PCL_ERR_INVALID; }
  int rc = pcl_check_inputs(c, im, r->evalp, r->B);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const PclLaunchPlan pl = pcl_plan(c, r->B, true);
  const size_t need = (size_t)pl.gx * pl.NS * (size_t)r->B;
  if (need > r->partial_floats) {
    pcl_pool_free(r->partial, st);
    PCL_CUDA(pcl_pool_alloc((void**)&r->partial, need * sizeof(double), st));
    r->partial_floats = need;
  }
  PclFinalize fin;
  memset(&fin, 0, sizeof(fin));
  fin.mode = PCL_FIN_REFINE;
  fin.loss = r->loss; fin.state = r->state; fin.evalp = r->evalp;
  fin.box = c->lo_hi_dev;
  fin.factor = r->factor; fin.patience = r->patience; fin.batch_semantics = r->batch_semantics;
  // small batches whose grid is one resident wave: all iterations in one cooperative launch
  if (num_iter >= 2 && pl.PB <= 16 && (long long)pl.gx * pl.gy <= (long long)pcl_num_sms() * 2 && pcl_env_int("PCL_PERSIST", 1) != 0) {
    rc = pcl_refine_run_persistent(r, c, im, pl, fin, num_iter, st);
    if (rc != 1) return rc;              // 1: the device refused the cooperative launch -> per-iteration launches below
  }
  for (int it = 0; it < num_iter; ++it) {
    r->steps_done += 1;
    fin.bc1 = 1.0 - pow(0.9, (double)r->steps_done);
    fin.bc2_sqrt = sqrt(1.0 - pow(0.999, (double)r->steps_done));
    // iterations after the first opt in to programmatic dependent launch (prologue overlaps the previous tail)
    rc = pcl_launch<true>(pl, c, im, r->evalp, r->B, r->partial, r->counters, fin, st, it > 0 && pcl_env_int("PCL_PDL", 1) != 0);
    if (rc) return rc;
  }
  return PCL_OK;
}

extern "C" int pcl_refine_read(const pcl_refine* r, float* pose_b6_dev, float* param_b6_dev, float* loss_b_dev,
                               double* lr_b_dev, void* stream) {
  if (!r || !r->block) { pcl_set_error("refine handle is null or was never reset"); return PCL_ERR_INVALID; }
  pcl_refine_read_kernel<<<(r->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->state, r->evalp, r->B, r->batch_semantics,
                                                                                pose_b6_dev, param_b6_dev, loss_b_dev, lr_b_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" void pcl_refine_destroy(pcl_refine* r) {
  if (!r) return;
  pcl_pool_free(r->block, r->owner);
  pcl_pool_free(r->partial, r->owner);
  pcl_pool_free(r->bc_dev, r->owner);
  free(r);
}
