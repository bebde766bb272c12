// Fused refinement of small candidate batches (B <= 16): device code shared by the persistent kernel and the
// per-iteration fallback kernel (pcl_refine.cu instantiates both per texel format).
//
// Replaces the optimisation loops of the reference: omniloc.py:44-58 (`omniloc`) and :249-269 (`omniloc_batch`) =
// per iteration { SamplingLoss / BatchSamplingLoss forward (omniloc.py:171-202, :311-356), autograd backward,
// Adam.step, ReduceLROnPlateau.step, clamp of the translation into the quantile box }.
//
// Decomposition (B200, 148 SMs x 2 resident CTAs of 256 threads, 128 registers):
//   * the cloud (or this rank's shard of it) is cut into G = 2·SMs - 1 contiguous point ranges balanced TO THE POINT,
//     one per compute CTA; a CTA keeps its range for the whole run.
//   * the candidates are cut into `nblk` pose blocks of <= 4 candidates.  A *phase* = one pose block over the CTA's
//     range: groups of 4 x 256 points held in registers are evaluated against the block's poses with the 8 sums per
//     pose accumulated in REGISTERS; the sub-1024-point remainder is spread over (pose, point) pairs, thread t taking
//     pose t % np, so that every thread of the grid ends up within one evaluation of the mean.  One warp butterfly
//     + fp64 shared-memory add per pose per phase (or per 32 evaluations on large clouds), one 64-byte record per
//     pose per CTA per phase to global memory (and, when the cloud is sharded over ranks, to every peer over NVLink).
//   * ONE SERVICE CTA (the last CTA of the grid, no points) owns the optimiser: for every (iteration, block) it waits
//     until all records have arrived, reduces them in a fixed two-level fp64 order, steps Adam / plateau / clamp with
//     one thread per (candidate, parameter), and publishes the block's next poses behind a release flag.
//   * split-phase hand-over: a compute CTA that has written its record of block b ARRIVES on b's counter and moves
//     straight on to the next block; it needs b's new poses only one whole round of phases later, and one of its
//     warps prefetches them into the spare pose buffer while the others still compute.  Nobody spins in steady
//     state: the barrier skew, the serial optimiser step, the NVLink latency and the straggler CTA are all hidden
//     behind the other blocks' work.  (B = 1 has nothing to hide behind and pays the chain once per iteration.)
#pragma once
#include "pcl_common.cuh"

#define PCL_RF_MAXB 16          // candidates of a fused run
#define PCL_RF_MAXNPB 4         // candidates per pose block (their 8 sums live in registers)
#define PCL_RF_MAXBLK 8
#define PCL_RF_KK 4             // points per thread per group
#ifndef PCL_RF_THREADS
#define PCL_RF_THREADS 512      // one CTA of 16 warps per SM (128 registers per thread: the whole register file)
#endif
#define PCL_RF_WARPS (PCL_RF_THREADS / 32)
#define PCL_RF_CTAS_PER_SM (512 / PCL_RF_THREADS)
#define PCL_RF_GROUP (PCL_RF_KK * PCL_RF_THREADS)
#define PCL_RF_ROWF (6 * PCL_RF_THREADS)   // floats of a resident row of PCL_RF_THREADS points in shared memory
#define PCL_RF_FLUSH 8          // groups between two flushes of the fp32 register sums into the fp64 shared-memory row
#define PCL_RF_MAXRANKS 8
#define PCL_RF_SMEM_STATIC (20 * 1024)   // upper bound of the persistent kernel's static shared memory (checked by static_assert)

// dynamic shared memory of `res_pts` resident points (whole rows)
inline size_t pcl_rf_res_bytes(int res_pts) { return (size_t)((res_pts + PCL_RF_THREADS - 1) / PCL_RF_THREADS) * PCL_RF_ROWF * sizeof(float); }

struct PclRfParams {
  PclCloudView C;
  PclImage I;
  int B, nblk, npb;             // candidates, pose blocks, candidates per block (the last block may hold fewer)
  long long p_begin, p_end;     // this rank's point range
  int G;                        // CTAs per rank
  int res_pts;                  // points of a CTA's range kept resident in shared memory (whole 2048-point groups, or the whole range); 0 = none
  int rank, nranks;
  unsigned long long* dbg;      // nullable: 4 counters per CTA, see pcl_refine_debug_stats (option RF_DEBUG)
  double* rec[PCL_RF_MAXRANKS];           // record buffers of all ranks (rec[rank] is local): [2][nblk][nranks*G][MAXNPB*8]
  unsigned int* arrive[PCL_RF_MAXRANKS];  // arrival counters of all ranks: [nblk], monotonic
  unsigned int arrive_base[PCL_RF_MAXBLK];   // value of block b's counter when this run starts
  int parity0;                  // record parity of this run's first iteration (records are double-buffered by iteration parity)
  unsigned int ready_base[PCL_RF_MAXBLK];   // tag base of block b: the poses for iteration `it` carry tag ready_base[b] + it
  unsigned long long* posebuf;  // [B][12] local: PclPose of every candidate for its next phase as self-validating 64-bit
                                // words {tag : 32 | float bits : 32} (a 64-bit store is single-copy atomic: no flag, no fence)
  const double* bc;             // [num_iter][2] = {1 - 0.9^step, sqrt(1 - 0.999^step)} (host libm, as python's `beta ** step`)
  int num_iter;
  PclRefineState* state;        // [B] global
  float* evalp;                 // [B][6] global
  float* loss;                  // [B] global (nullable)
  const float* box;             // clamp box {lo(3), hi(3)}
  double factor;
  int patience, batch_semantics;
};

// Flag polls are RELAXED loads (served by L2): an acquire load makes the SM invalidate its whole L1 (CCTL.IVALL in the
// SASS) on every poll, which throws away the texel lines the co-resident compute warps live on.  Everything published
// behind a flag (records, poses) is read with ld.global.cg, i.e. from L2, after the poll has returned the flag value, and
// was made visible there by the writer's release before it raised the flag.
__device__ __forceinline__ unsigned int pcl_ld_poll_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int pcl_ld_poll_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 64-bit pose words {tag : 32 | float bits : 32}: relaxed accesses served by L2; a word validates itself through its tag
__device__ __forceinline__ unsigned long long pcl_ld_word_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void pcl_st_word_gpu(unsigned long long* p, const unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Warp reduction of 8 values with a halving butterfly (7 + 2 shuffles instead of 40).  On return lane L with
// L % 4 == 0 holds the warp sum of value index pcl_rf_bfly_index(L) in v[0].
__device__ __forceinline__ void pcl_rf_warp_reduce8(float (&v)[8], const int lane) {
  int n = 8;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (n > 1) {
      const bool up = (lane & o) != 0;
      n >>= 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
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
__device__ __forceinline__ int pcl_rf_bfly_index(const int lane) {
  return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

__device__ __forceinline__ void pcl_rf_zero(PclAcc& a) {
  a.se = -0.f; a.sm = -0.f; a.ax = -0.f; a.ay = -0.f; a.az = -0.f; a.tx = -0.f; a.ty = -0.f; a.tz = -0.f;   // -0 + x == x
}

// fp32 register sums of ALL poses of the block -> the warp's private fp64 shared-memory rows.  The NPB butterflies are
// issued stage by stage side by side (the shuffles of a stage are independent across poses): the dependent chain of
// one reduction instead of NPB of them in a row.  Poses beyond np hold -0 sums and add nothing.
template <int NPB>
__device__ __forceinline__ void pcl_rf_flush_all(PclAcc (&acc)[NPB], double (*rows)[PCL_NSUM], const int lane) {
  float v[NPB][8];
#pragma unroll
  for (int p = 0; p < NPB; ++p) {
    v[p][0] = acc[p].se; v[p][1] = acc[p].sm; v[p][2] = acc[p].ax; v[p][3] = acc[p].ay;
    v[p][4] = acc[p].az; v[p][5] = acc[p].tx; v[p][6] = acc[p].ty; v[p][7] = acc[p].tz;
  }
  int n = 8;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (n > 1) {
      const bool up = (lane & o) != 0;
      n >>= 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < n) {
#pragma unroll
          for (int p = 0; p < NPB; ++p) {
            const float send = up ? v[p][i] : v[p][i + n];
            const float keep = up ? v[p][i + n] : v[p][i];
            v[p][i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < NPB; ++p) v[p][0] += __shfl_xor_sync(0xffffffffu, v[p][0], o);
    }
  }
  if ((lane & 3) == 0) {
    const int k = pcl_rf_bfly_index(lane);
#pragma unroll
    for (int p = 0; p < NPB; ++p) rows[p][k] += (double)v[p][0];
  }
#pragma unroll
  for (int p = 0; p < NPB; ++p) pcl_rf_zero(acc[p]);
}

// Streamed (non-resident) point loads: read-only path, no L1 allocation, and marked evict-first in L2 — a cloud that is
// larger than L2 must not push the texel table (re-read every iteration at unpredictable places) out of it
// (measured, B = 6, 1024x2048: 10 M points 496 -> 406 us per iteration, 5 M points 271 -> 211 us).
__device__ __forceinline__ unsigned long long pcl_rf_stream_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float pcl_rf_ld_stream(const float* p, const unsigned long long pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

// One phase: the CTA's points [c_begin, c_end) against the np poses of a block.  On return (after the trailing
// __syncthreads) s_acc[w][p][0..7] hold warp w's fp64 sums {Σ m e, Σ m, a(3), τ(3)} of pose p.
struct PclRfNoHook { __device__ __forceinline__ void operator()() const {} };

// `hook` runs once, after the full groups and before the remainder (the persistent kernel prefetches the next phase's
// poses there with one warp).
//
// Resident points: s_pts (nullable) holds the first `res_n` points of the range in rows of 512 points, each row as
// x[512] | y[512] | z[512] | r[512] | g[512] | b[512] (the persistent kernel copies them once per launch): a thread's 24
// values of a group sit at compile-time offsets from ONE base address.  Groups that lie inside read shared memory
// instead of global memory, the remainder does when the whole range is resident.
template <int FMT, int NPB, typename Hook>
__device__ __forceinline__ void pcl_rf_phase(const PclCloudView& C, const PclImage& I, const PclPose* __restrict__ s_pose, const int np,
                                             const long long c_begin, const long long c_end, const float* __restrict__ s_pts, const int res_n,
                                             double (*s_acc)[NPB][PCL_NSUM],
                                             const int tid, const int lane, const int warp, const Hook& hook
#ifdef PCL_RF_TRACE
                                             , long long* trace
#endif
                                             ) {
#ifdef PCL_RF_TRACE
  const long long tr0 = clock64();
#endif
  // each warp owns its rows of s_acc: no CTA-wide synchronisation until the end of the phase
  for (int i = lane; i < NPB * PCL_NSUM; i += 32) (&s_acc[warp][0][0])[i] = 0.0;
  __syncwarp();

  PclAcc acc[NPB];
#pragma unroll
  for (int p = 0; p < NPB; ++p) pcl_rf_zero(acc[p]);

  const long long n_groups = (c_end - c_begin) / PCL_RF_GROUP;
  const unsigned long long pol = pcl_rf_stream_policy();
  long long i0 = c_begin + tid;
  int pending = 0;
  for (long long g = 0; g < n_groups; ++g, i0 += PCL_RF_GROUP) {
    float px[PCL_RF_KK], py[PCL_RF_KK], pz[PCL_RF_KK], cr[PCL_RF_KK], cg[PCL_RF_KK], cb[PCL_RF_KK];
    if ((int)(g + 1) * PCL_RF_GROUP <= res_n) {                  // CTA-uniform: the group is resident
      const float* s = s_pts + (int)g * (PCL_RF_KK * PCL_RF_ROWF) + tid;
#pragma unroll
      for (int j = 0; j < PCL_RF_KK; ++j) {
        const float* sj = s + j * PCL_RF_ROWF;
        px[j] = sj[0]; py[j] = sj[PCL_RF_THREADS]; pz[j] = sj[2 * PCL_RF_THREADS];
        cr[j] = sj[3 * PCL_RF_THREADS]; cg[j] = sj[4 * PCL_RF_THREADS]; cb[j] = sj[5 * PCL_RF_THREADS];
      }
    } else {
#pragma unroll
      for (int j = 0; j < PCL_RF_KK; ++j) {
        const long long i = i0 + (long long)j * PCL_RF_THREADS;
        px[j] = pcl_rf_ld_stream(C.x + i, pol); py[j] = pcl_rf_ld_stream(C.y + i, pol); pz[j] = pcl_rf_ld_stream(C.z + i, pol);
        cr[j] = pcl_rf_ld_stream(C.r + i, pol); cg[j] = pcl_rf_ld_stream(C.g + i, pol); cb[j] = pcl_rf_ld_stream(C.b + i, pol);
      }
    }
#pragma unroll
    for (int p = 0; p < NPB; ++p) {
      if (p < np) {
        const PclPose pose = s_pose[p];
#pragma unroll
        for (int j = 0; j < PCL_RF_KK; ++j) pcl_eval<FMT, true>(pose, I, px[j], py[j], pz[j], cr[j], cg[j], cb[j], true, acc[p]);
      }
    }
    if (++pending == PCL_RF_FLUSH) {
      pending = 0;
      pcl_rf_flush_all<NPB>(acc, s_acc[warp], lane);
    }
  }

#ifdef PCL_RF_TRACE
  const long long tr1 = clock64();
#endif
  hook();

  // remainder (< 1024 points): thread t takes pose t % np and every S-th point from slot t / np, S = 256 / np
  const long long rem0 = c_begin + n_groups * PCL_RF_GROUP;
  const int r = (int)(c_end - rem0);
  const bool rem_res = (long long)res_n >= c_end - c_begin;      // the whole range is resident
  const int rem_off = (int)(rem0 - c_begin);
  if (r > 0) {
    // tid / np and THREADS / np for np <= 4 by multiply-shift (exact for tid <= 65535 / 4)
    const unsigned int inv = np == 1 ? 65536u : np == 2 ? 32768u : np == 3 ? 21846u : 16384u;
    const int S = (int)(((unsigned int)PCL_RF_THREADS * inv) >> 16);
    const int slot = (int)(((unsigned int)tid * inv) >> 16), p_t = tid - slot * np;
    PclAcc ar;
    pcl_rf_zero(ar);
    if (slot < S) {
      const PclPose pose = s_pose[p_t];
      for (int m0 = slot; m0 < r; m0 += PCL_RF_KK * S) {
        float px[PCL_RF_KK], py[PCL_RF_KK], pz[PCL_RF_KK], cr[PCL_RF_KK], cg[PCL_RF_KK], cb[PCL_RF_KK];
        bool ok[PCL_RF_KK];
#pragma unroll
        for (int j = 0; j < PCL_RF_KK; ++j) {
          const int m = m0 + j * S;
          ok[j] = m < r;
          const int mm = ok[j] ? m : slot;                    // masked slots re-read a valid point and are discarded
          if (rem_res) {
            const int idx = rem_off + mm;                     // (the remainder starts on a row boundary: whole groups precede it)
            const float* sj = s_pts + (idx / PCL_RF_THREADS) * PCL_RF_ROWF + (idx % PCL_RF_THREADS);
            px[j] = sj[0]; py[j] = sj[PCL_RF_THREADS]; pz[j] = sj[2 * PCL_RF_THREADS];
            cr[j] = sj[3 * PCL_RF_THREADS]; cg[j] = sj[4 * PCL_RF_THREADS]; cb[j] = sj[5 * PCL_RF_THREADS];
          } else {
            const long long i = rem0 + mm;
            px[j] = __ldg(C.x + i); py[j] = __ldg(C.y + i); pz[j] = __ldg(C.z + i);
            cr[j] = __ldg(C.r + i); cg[j] = __ldg(C.g + i); cb[j] = __ldg(C.b + i);
          }
        }
#pragma unroll
        for (int j = 0; j < PCL_RF_KK; ++j) pcl_eval<FMT, true>(pose, I, px[j], py[j], pz[j], cr[j], cg[j], cb[j], ok[j], ar);
      }
    }
#pragma unroll
    for (int p = 0; p < NPB; ++p) {
      if (p == p_t) {                                            // predicated adds
        acc[p].se += ar.se; acc[p].sm += ar.sm;
        acc[p].ax += ar.ax; acc[p].ay += ar.ay; acc[p].az += ar.az;
        acc[p].tx += ar.tx; acc[p].ty += ar.ty; acc[p].tz += ar.tz;
      }
    }
  }
  pcl_rf_flush_all<NPB>(acc, s_acc[warp], lane);
#ifdef PCL_RF_TRACE
  const long long tr2 = clock64();
#endif
  __syncthreads();
#ifdef PCL_RF_TRACE
  if (trace) { trace[0] += tr1 - tr0; trace[1] += tr2 - tr1; trace[2] += clock64() - tr2; }
#endif
}

// the CTA's record of a phase: rec8[p*8 + s] = Σ_warps s_acc[w][p][s]   (threads < np*8)
template <int NPB>
__device__ __forceinline__ double pcl_rf_cta_sum(double (*s_acc)[NPB][PCL_NSUM], const int tid) {
  const int p = tid >> 3, s = tid & 7;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < PCL_RF_WARPS; ++w) t += s_acc[w][p][s];
  return t;
}

struct PclRfConsts {
  const float* box;
  double factor;
  int patience, batch_semantics;
};

// Every thread of the CTA: deterministic fp64 reduction of the block's n_rec records (slot stride MAXNPB*8 doubles; one warp
// per pair of sums), then one thread per (candidate, parameter): loss + analytic gradient + torch.optim.Adam
// (single-tensor path, betas 0.9/0.999, eps 1e-8; fp32 tensors, fp64 python scalars) + ReduceLROnPlateau(mode=min,
// threshold 1e-4 rel, cooldown 0, min_lr 0, eps 1e-8) + translation clamp, exactly the split torch has
// (omniloc.py:33,37,49-58).  st/evalp/pose are the block's np entries (shared or global memory).
// Contains __syncthreads: call from all threads.  On return evalp/pose hold the next iteration's pose.
__device__ __forceinline__ void pcl_rf_finalize(const double* __restrict__ rec, const int n_rec, const int np, double2* __restrict__ s_sum,
                                                PclRefineState* __restrict__ st, float* __restrict__ evalp, PclPose* __restrict__ pose,
                                                const PclImage& I, const PclRfConsts& k, const double bc1, const double bc2_sqrt, const int tid,
                                                long long* dbg_split = nullptr) {
  const long long f0 = dbg_split ? clock64() : 0;
  const int nitems = np * (PCL_NSUM / 2);                        // 16-byte pairs of sums: at most 16 = one per warp
  // the clamp box of the parameter thread, issued now and consumed at the very end of the optimiser step
  float box_lo = 0.0f, box_hi = 0.0f;
  if (tid < 32 && (tid % 6) < 3) { box_lo = __ldg(k.box + tid % 6); box_hi = __ldg(k.box + 3 + tid % 6); }
  {
    // warp w reduces item w: lane l adds records l, l + 32, ... in order, then a 5-step xor tree (every lane ends with the
    // same bits: each step adds the same two partial sums on both sides) — one L2 round trip, no shared-memory chain
    const int lane = tid & 31, warp = tid >> 5;
    if (warp < nitems) {
      const double2* src = reinterpret_cast<const double2*>(rec) + warp;
      constexpr int stride = PCL_RF_MAXNPB * PCL_NSUM / 2;
      double2 t = make_double2(0.0, 0.0);
      // eight records per lane per round, all eight loads in flight before the first add (one L2 round trip per 256
      // records: the single-GPU case is one round)
      for (int s0 = lane; s0 < n_rec; s0 += 256) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int s = s0 + 32 * u;
          v[u] = s < n_rec ? __ldcg(src + (size_t)s * stride) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { t.x += v[u].x; t.y += v[u].y; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o); }
      if (lane == 0) s_sum[warp] = t;
    }
  }
  __syncthreads();
  if (dbg_split) dbg_split[0] += clock64() - f0;                 // records reduced
  if (tid < 32) {
    const int p = tid / 6, i = tid - 6 * p;
    const bool act = tid < np * 6;
    const unsigned int mask = __ballot_sync(0xffffffffu, act);
    if (act) {
      double sums[PCL_NSUM];
#pragma unroll
      for (int h = 0; h < PCL_NSUM / 2; ++h) { const double2 t = s_sum[p * (PCL_NSUM / 2) + h]; sums[2 * h] = t.x; sums[2 * h + 1] = t.y; }
      float p6[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) p6[j] = evalp[6 * p + j];
      float loss, grad[6];
      pcl_finish_gradient(p6, pose[p], I, sums, &loss, nullptr, grad);
      float gi = grad[0];
#pragma unroll
      for (int j = 1; j < 6; ++j) gi = (i == j) ? grad[j] : gi;
      PclRefineState& S = st[p];
      // Adam: exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2); addcdiv_(m, denom, -step_size)
      const float step_size = (float)(S.lr / bc1);
      const float w1 = (float)(1.0 - 0.9), b2 = 0.999f, w2 = (float)(1.0 - 0.999);
      const float m = S.m[i] + w1 * (gi - S.m[i]);
      const float v = S.v[i] * b2 + (w2 * gi) * gi;
      const float denom = sqrtf(v) / (float)bc2_sqrt + 1e-8f;
      const float newp = S.param[i] - step_size * (m / denom);
      __syncwarp(mask);                                          // every parameter thread has read lr
      S.m[i] = m; S.v[i] = v;
      if (i == 0) {
        S.last_loss = loss;
        S.step += 1;
        const double cur = (double)loss;                         // scheduler.step(loss)
        if (cur < S.best * (1.0 - 1e-4)) { S.best = cur; S.bad = 0; } else { S.bad += 1; }
        if (S.bad > k.patience) {
          const double new_lr = fmax(S.lr * k.factor, 0.0);
          if (S.lr - new_lr > 1e-8) S.lr = new_lr;
          S.bad = 0;
        }
      }
      // clamp the translation into the quantile box; batch semantics evaluates the pre-clamp copy next (omniloc.py:260-269)
      float c = newp;
      if (i < 3) c = fminf(fmaxf(c, box_lo), box_hi);
      S.param[i] = c;
      const float ev = k.batch_semantics ? newp : c;
      evalp[6 * p + i] = ev;
      // the next pose: the three angle threads take cos/sin of their own angle, the first thread assembles R
      float* trig = reinterpret_cast<float*>(s_sum + 32) + 8 * p;  // scratch behind the 16 sums
      if (i >= 3) { trig[2 * (i - 3)] = cosf(ev); trig[2 * (i - 3) + 1] = sinf(ev); }
      __syncwarp(mask);
      if (i == 0) pcl_pose_from_trig(trig[0], trig[1], trig[2], trig[3], trig[4], trig[5], evalp[6 * p], evalp[6 * p + 1], evalp[6 * p + 2], pose[p]);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// persistent kernel: ALL iterations in one cooperative launch; CTA G (the last one) is the service CTA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pcl_rf_spin(const unsigned int* ctr, const unsigned int target, const bool sys, const unsigned int ns) {
  if (sys) { while ((int)(pcl_ld_poll_sys(ctr) - target) < 0) { if (ns) __nanosleep(ns); } }
  else { while ((int)(pcl_ld_poll_gpu(ctr) - target) < 0) { if (ns) __nanosleep(ns); } }
}

template <int FMT, int NPB>
__global__ void __launch_bounds__(PCL_RF_THREADS, PCL_RF_CTAS_PER_SM) pcl_refine_persistent_kernel(const PclRfParams ps) {
  __shared__ __align__(16) PclPose s_pose[2][PCL_RF_MAXB];       // compute: [phase parity][NPB] used; service: [0][B]
  __shared__ double s_acc[2][PCL_RF_WARPS][NPB][PCL_NSUM];
  __shared__ double2 s_sum[PCL_RF_THREADS];
  __shared__ PclRefineState s_state[PCL_RF_MAXB];                // service CTA only
  __shared__ float s_evalp[PCL_RF_MAXB][6];
  __shared__ int s_pref[2];                                      // phase whose poses sit in s_pose[parity]
  extern __shared__ __align__(16) float s_pts[];                 // resident points in rows of 512: x|y|z|r|g|b, 512 floats each

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = ps.G, n_rec = G * ps.nranks;
  const size_t rec_blk = (size_t)n_rec * PCL_RF_MAXNPB * PCL_NSUM;          // doubles per block

  if (cta == G) {
    // ---------------- service CTA: reduce, step the optimiser, publish ----------------
    const PclRfConsts k = {ps.box, ps.factor, ps.patience, ps.batch_semantics};
    if (tid < ps.B) {
      s_state[tid] = ps.state[tid];
#pragma unroll
      for (int i = 0; i < 6; ++i) s_evalp[tid][i] = ps.evalp[6 * tid + i];
      pcl_pose_from_params(s_evalp[tid], s_pose[0][tid]);
    }
    __syncthreads();
    long long sv_wait = 0, sv_work = 0, sv_reduce = 0;
    for (int it = 0; it < ps.num_iter; ++it) {
      for (int b = 0; b < ps.nblk; ++b) {
        const int p0 = b * ps.npb, np = min(ps.npb, ps.B - p0);
        const long long s0 = ps.dbg ? clock64() : 0;
        if (tid == 0) pcl_rf_spin(ps.arrive[ps.rank] + b, ps.arrive_base[b] + (unsigned int)(it + 1) * (unsigned int)n_rec, ps.nranks > 1, 32u);
        __syncthreads();
        const long long s1 = ps.dbg ? clock64() : 0;
        // records are double-buffered by iteration parity: a fast rank's next record must not overwrite the copy a slower
        // rank's service CTA is still reading
        pcl_rf_finalize(ps.rec[ps.rank] + ((size_t)((ps.parity0 + it) & 1) * ps.nblk + b) * rec_blk, n_rec, np, s_sum, s_state + p0, &s_evalp[p0][0], &s_pose[0][p0], ps.I, k,
                        __ldg(ps.bc + 2 * it), __ldg(ps.bc + 2 * it + 1), tid, ps.dbg ? &sv_reduce : nullptr);
        // publish: every float of the block's poses travels with the tag of the iteration it is for, in one 64-bit store
        // (single-copy atomic) — no flag, no release, no second round trip on the reader's side
        if (tid < np * 12) {
          const unsigned long long tag = (unsigned long long)(ps.ready_base[b] + (unsigned int)(it + 1));
          pcl_st_word_gpu(ps.posebuf + (size_t)p0 * 12 + tid, (tag << 32) | (unsigned long long)__float_as_uint(reinterpret_cast<const float*>(&s_pose[0][p0])[tid]));
        }
        if (tid == 0) {
          if (ps.dbg) {
            sv_wait += s1 - s0; sv_work += clock64() - s1;
            if (b == ps.nblk - 1) {                              // end of iteration `it`: wall clock [ns] for the per-iteration timeline
              unsigned long long tns;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
              ps.dbg[4 * (size_t)(G + 1) + it] = tns;
            }
          }
        }
      }
    }
    if (ps.dbg && tid == 0) { ps.dbg[4 * (size_t)G] = (unsigned long long)sv_wait; ps.dbg[4 * (size_t)G + 1] = (unsigned long long)sv_work; ps.dbg[4 * (size_t)G + 2] = (unsigned long long)sv_reduce; ps.dbg[4 * (size_t)G + 3] = 0; }
    if (tid < ps.B) {                                            // the end state
      ps.state[tid] = s_state[tid];
#pragma unroll
      for (int i = 0; i < 6; ++i) ps.evalp[6 * tid + i] = s_evalp[tid][i];
      if (ps.loss) ps.loss[tid] = s_state[tid].last_loss;
    }
    return;
  }

  // ---------------- compute CTAs ----------------
  // this CTA's points: ranges of the rank's shard that differ by at most one point
  const long long n_pts = ps.p_end - ps.p_begin;
  const long long c_begin = ps.p_begin + n_pts * (long long)cta / G;
  const long long c_end = ps.p_begin + n_pts * (long long)(cta + 1) / G;
  const size_t my_slot = ((size_t)ps.rank * G + cta) * PCL_RF_MAXNPB * PCL_NSUM;
  if (tid < 2) s_pref[tid] = -1;
  // resident points: copied once, read by every phase of every iteration (no HBM/L2 traffic and no long-latency load for
  // the point stream after this prologue; the only global loads of an iteration are the texel gathers)
  const int res_n = (int)min((long long)ps.res_pts, c_end - c_begin);
  for (int k = tid; k < res_n; k += PCL_RF_THREADS) {
    const long long i = c_begin + k;
    float* d = s_pts + (k / PCL_RF_THREADS) * PCL_RF_ROWF + (k % PCL_RF_THREADS);
    d[0] = __ldg(ps.C.x + i); d[PCL_RF_THREADS] = __ldg(ps.C.y + i); d[2 * PCL_RF_THREADS] = __ldg(ps.C.z + i);
    d[3 * PCL_RF_THREADS] = __ldg(ps.C.r + i); d[4 * PCL_RF_THREADS] = __ldg(ps.C.g + i); d[5 * PCL_RF_THREADS] = __ldg(ps.C.b + i);
  }
  __syncthreads();

  int ph = 0;
  long long t_busy = 0, t_wait = 0, n_hit = 0;
#ifdef PCL_RF_TRACE
  long long trace[3] = {0, 0, 0};
  __shared__ unsigned long long s_trace[4];
  if (tid < 4) s_trace[tid] = 0ull;
  __syncthreads();
#endif
  for (int it = 0; it < ps.num_iter; ++it) {
    for (int b = 0; b < ps.nblk; ++b, ++ph) {
      const int buf = ph & 1;
      const int p0 = b * ps.npb, np = min(ps.npb, ps.B - p0);
      const long long c0 = ps.dbg ? clock64() : 0;
      if (it == 0) {
        if (tid < np) pcl_pose_from_params(ps.evalp + 6 * (size_t)(p0 + tid), s_pose[buf][tid]);
        __syncthreads();
      } else if (s_pref[buf] == ph) {
        n_hit += 1;
      } else {                                                   // not prefetched (uniform: written before the last barrier)
        if (tid < np * 12) {                                     // every word is polled by the thread that needs it: one L2 round trip
          const unsigned int want = ps.ready_base[b] + (unsigned int)it;
          unsigned long long w;
          do { w = pcl_ld_word_gpu(ps.posebuf + (size_t)p0 * 12 + tid); } while ((int)((unsigned int)(w >> 32) - want) < 0);
          reinterpret_cast<float*>(&s_pose[buf][0])[tid] = __uint_as_float((unsigned int)w);
        }
        __syncthreads();
      }
      // the last warp tries to fetch the NEXT phase's poses while the others are still in this phase
      const int nb = (b + 1 == ps.nblk) ? 0 : b + 1, nit = (b + 1 == ps.nblk) ? it + 1 : it;
      auto hook = [&]() {
        if (warp != PCL_RF_WARPS - 1 || nit == 0 || nit >= ps.num_iter) return;
        const int q0 = nb * ps.npb, nq = min(ps.npb, ps.B - q0);
        const unsigned int want = ps.ready_base[nb] + (unsigned int)nit;
        unsigned long long w0 = 0ull, w1 = 0ull;                   // <= 48 words: two per lane
        bool ok = true;
        if (lane < nq * 12) { w0 = pcl_ld_word_gpu(ps.posebuf + (size_t)q0 * 12 + lane); ok = (int)((unsigned int)(w0 >> 32) - want) >= 0; }
        if (lane + 32 < nq * 12) { w1 = pcl_ld_word_gpu(ps.posebuf + (size_t)q0 * 12 + lane + 32); ok = ok && (int)((unsigned int)(w1 >> 32) - want) >= 0; }
        if (!__all_sync(0xffffffffu, ok)) return;                  // not published yet: the phase start will wait for it
        if (lane < nq * 12) reinterpret_cast<float*>(&s_pose[buf ^ 1][0])[lane] = __uint_as_float((unsigned int)w0);
        if (lane + 32 < nq * 12) reinterpret_cast<float*>(&s_pose[buf ^ 1][0])[lane + 32] = __uint_as_float((unsigned int)w1);
        if (lane == 0) s_pref[buf ^ 1] = ph + 1;
      };
      const long long c1 = ps.dbg ? clock64() : 0;
      pcl_rf_phase<FMT, NPB>(ps.C, ps.I, s_pose[buf], np, c_begin, c_end, s_pts, res_n, s_acc[buf], tid, lane, warp, hook
#ifdef PCL_RF_TRACE
                             , trace
#endif
                             );
      if (ps.dbg) { t_wait += c1 - c0; t_busy += clock64() - c1; }
      if (warp == 0) {
        if (lane < np * PCL_NSUM) {
          const double t = pcl_rf_cta_sum<NPB>(s_acc[buf], lane);
          const size_t off = ((size_t)((ps.parity0 + it) & 1) * ps.nblk + b) * rec_blk + my_slot + lane;
          for (int rk = 0; rk < ps.nranks; ++rk) ps.rec[rk][off] = t;
        }
        __syncwarp();
        // arrive with RELEASE semantics on the reduction itself (cumulative over the __syncwarp above).  A
        // __threadfence() here would be an acq_rel fence, whose acquire half invalidates the SM's whole L1 — the
        // texel and point lines the CTA is about to reuse.
        if (lane < ps.nranks) {
          if (ps.nranks > 1) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(ps.arrive[lane] + b) : "memory");
          else asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ps.arrive[0] + b) : "memory");
        }
      }
    }
  }
#ifdef PCL_RF_TRACE
  if (ps.dbg) {      // per CTA: mean over warps of {cycles in full groups, remainder + flush, waiting at the phase barrier}, and the largest barrier wait
    if (lane == 0) {
      atomicAdd(&s_trace[0], (unsigned long long)trace[0]); atomicAdd(&s_trace[1], (unsigned long long)trace[1]);
      atomicAdd(&s_trace[2], (unsigned long long)trace[2]); atomicMax(&s_trace[3], (unsigned long long)trace[2]);
    }
    __syncthreads();
    if (tid == 0) {
      ps.dbg[4 * (size_t)cta] = s_trace[0] / PCL_RF_WARPS; ps.dbg[4 * (size_t)cta + 1] = s_trace[1] / PCL_RF_WARPS;
      ps.dbg[4 * (size_t)cta + 2] = s_trace[2] / PCL_RF_WARPS; ps.dbg[4 * (size_t)cta + 3] = s_trace[3];
    }
    return;
  }
#endif
  if (ps.dbg && tid == 0) {
    ps.dbg[4 * (size_t)cta] = (unsigned long long)t_busy; ps.dbg[4 * (size_t)cta + 1] = (unsigned long long)t_wait;
    ps.dbg[4 * (size_t)cta + 2] = (unsigned long long)n_hit; ps.dbg[4 * (size_t)cta + 3] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// per-iteration fallback (single rank): grid (G, nblk), the last CTA of a block finishes it.  Same ranges, same
// phase arithmetic, same record order and the same finalize as the persistent kernel: bit-identical trajectories.
// ------------------------------------------------------------------------------------------------
template <int FMT, int NPB>
__global__ void __launch_bounds__(PCL_RF_THREADS, PCL_RF_CTAS_PER_SM) pcl_refine_iter_kernel(const PclRfParams ps, unsigned int* __restrict__ tickets,
                                                                         const double bc1, const double bc2_sqrt) {
  __shared__ __align__(16) PclPose s_pose[PCL_RF_MAXNPB];
  __shared__ double s_acc[PCL_RF_WARPS][NPB][PCL_NSUM];
  __shared__ double2 s_sum[PCL_RF_THREADS];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, b = blockIdx.y, G = ps.G;
  const int p0 = b * ps.npb, np = min(ps.npb, ps.B - p0);
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");             // the previous iteration's poses are complete
#endif
  if (tid < np) pcl_pose_from_params(ps.evalp + 6 * (size_t)(p0 + tid), s_pose[tid]);
  __syncthreads();
  const long long n_pts = ps.p_end - ps.p_begin;
  const long long c_begin = ps.p_begin + n_pts * (long long)cta / G;
  const long long c_end = ps.p_begin + n_pts * (long long)(cta + 1) / G;
  pcl_rf_phase<FMT, NPB>(ps.C, ps.I, s_pose, np, c_begin, c_end, nullptr, 0, s_acc, tid, lane, warp, PclRfNoHook()
#ifdef PCL_RF_TRACE
                         , nullptr
#endif
                         );
  double* rec = ps.rec[0] + (size_t)b * G * PCL_RF_MAXNPB * PCL_NSUM;
  if (tid < np * PCL_NSUM) rec[(size_t)cta * PCL_RF_MAXNPB * PCL_NSUM + tid] = pcl_rf_cta_sum<NPB>(s_acc, tid);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(tickets + b, 1u) == (unsigned int)G - 1u);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const PclRfConsts k = {ps.box, ps.factor, ps.patience, ps.batch_semantics};
  pcl_rf_finalize(rec, G, np, s_sum, ps.state + p0, ps.evalp + 6 * (size_t)p0, s_pose, ps.I, k, bc1, bc2_sqrt, tid);
  if (tid < np && ps.loss) ps.loss[p0 + tid] = ps.state[p0 + tid].last_loss;
  if (tid == 0) tickets[b] = 0u;                                 // self-resetting
}

// launch helpers instantiated per texel format (pcl_refine_fmt*.cu)
template <int FMT> cudaError_t pcl_rf_launch_persistent(const PclRfParams& ps, cudaStream_t st);
template <int FMT> cudaError_t pcl_rf_launch_iter(const PclRfParams& ps, unsigned int* tickets, double bc1, double bc2_sqrt, bool pdl, cudaStream_t st);

#define PCL_RF_INSTANTIATE(FMT)                                                                                                     \
  template <> cudaError_t pcl_rf_launch_persistent<FMT>(const PclRfParams& ps, cudaStream_t st) {                                   \
    PclRfParams p = ps;                                                                                                             \
    void* args[] = {&p};                                                                                                            \
    const void* fn = ps.npb == 1 ? (const void*)pcl_refine_persistent_kernel<FMT, 1>                                                \
                   : ps.npb == 2 ? (const void*)pcl_refine_persistent_kernel<FMT, 2>                                                \
                   : ps.npb == 3 ? (const void*)pcl_refine_persistent_kernel<FMT, 3>                                                \
                                 : (const void*)pcl_refine_persistent_kernel<FMT, 4>;                                               \
    const size_t smem = pcl_rf_res_bytes(ps.res_pts);                                                      \
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                               \
    if (e != cudaSuccess) return e;                                                                                                 \
    return cudaLaunchCooperativeKernel(fn, dim3(ps.G + 1), dim3(PCL_RF_THREADS), args, smem, st);                                   \
  }                                                                                                                                 \
  template <> cudaError_t pcl_rf_launch_iter<FMT>(const PclRfParams& ps, unsigned int* tickets, double bc1, double bc2_sqrt, bool pdl, \
                                                  cudaStream_t st) {                                                                \
    cudaLaunchConfig_t cfg;                                                                                                         \
    memset(&cfg, 0, sizeof(cfg));                                                                                                   \
    cfg.gridDim = dim3(ps.G, ps.nblk);                                                                                              \
    cfg.blockDim = dim3(PCL_RF_THREADS);                                                                                               \
    cfg.stream = st;                                                                                                                \
    cudaLaunchAttribute attr[1];                                                                                                    \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                                \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                                         \
    cfg.attrs = attr;                                                                                                               \
    cfg.numAttrs = pdl ? 1 : 0;                                                                                                     \
    switch (ps.npb) {                                                                                                               \
      case 1: return cudaLaunchKernelEx(&cfg, pcl_refine_iter_kernel<FMT, 1>, ps, tickets, bc1, bc2_sqrt);                          \
      case 2: return cudaLaunchKernelEx(&cfg, pcl_refine_iter_kernel<FMT, 2>, ps, tickets, bc1, bc2_sqrt);                          \
      case 3: return cudaLaunchKernelEx(&cfg, pcl_refine_iter_kernel<FMT, 3>, ps, tickets, bc1, bc2_sqrt);                          \
      default: return cudaLaunchKernelEx(&cfg, pcl_refine_iter_kernel<FMT, 4>, ps, tickets, bc1, bc2_sqrt);                         \
    }                                                                                                                               \
  }
