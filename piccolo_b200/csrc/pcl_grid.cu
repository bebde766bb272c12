// Structured start-grid scoring: the T x R loss table of trim_input_loss (utils.py:484-499) when the rotation list
// contains rotations related by an in-plane turn about the camera z axis, R_j = Rz(delta_j)·R_base.
//
// Such rotations see every point at the same elevation (same panorama row) and at azimuths that differ by the
// constant delta_j, so the rigid transform, rho, theta, the row coordinate and ONE azimuth atan2 are computed once
// per (point, translation, group) and each member rotation only adds its offset, wraps, and samples.  All yaw-only
// lists (generate_rot_points with pitch = roll = 0, utils.py:302-325) are one group; the 24 distinct rotations of the
// reference's 4x4x4 Euler lattice are 6 groups of 4.  Lists without that structure degenerate to R groups of one
// member (the cost of the generic kernel).
//
// Two kernels:
//   pcl_grid_plan_kernel   one warp: rotation matrices in fp64 from the fp32 angles, greedy grouping on the third
//                          row of R (e3ᵀR_j == e3ᵀR_b  <=>  R_j R_bᵀ is a rotation about z), delta = atan2 of R_j R_bᵀ
//   pcl_grid_score_kernel  grid.y = blocks of TB translations (TB·R <= 32 poses per CTA), grid.x = balanced row
//                          ranges, whole resident waves; per CTA the same fp64 partial records + last-block-done
//                          deterministic reduction as the generic kernel (pcl_sampling.cu)
#include "pcl_common.cuh"

#include <stdlib.h>
#include <string.h>

#define PCL_GRID_MAX_ROT 32
#ifndef PCL_GRID_MAX_WAVES
#define PCL_GRID_MAX_WAVES 4
#endif
#define PCL_GRID_TOL 1e-6       // third rows closer than this (fp64 from the fp32 angles) are the same group

struct PclGridPlan {
  int R, NG;
  int g_start[PCL_GRID_MAX_ROT], g_count[PCL_GRID_MAX_ROT];   // member slots of group g
  int slot_rot[PCL_GRID_MAX_ROT];                             // rotation index j of a slot
  float slot_delta[PCL_GRID_MAX_ROT];                         // its azimuth offset against the group's base, (-pi, pi]
  float base_ypr[PCL_GRID_MAX_ROT][3];                        // the group's base rotation (its first member's angles)
};

__global__ void pcl_grid_plan_kernel(const float* __restrict__ rot, const int R, PclGridPlan* __restrict__ plan) {
  __shared__ double s_R[PCL_GRID_MAX_ROT][9];
  __shared__ int s_group[PCL_GRID_MAX_ROT], s_base[PCL_GRID_MAX_ROT];
  const int j = threadIdx.x;
  // the scoring kernel copies the whole plan into shared memory: no uninitialised tail
  for (int i = j; i < (int)(sizeof(PclGridPlan) / 4); i += blockDim.x) reinterpret_cast<int*>(plan)[i] = 0;
  if (j < R) {
    const double y = rot[3 * j], p = rot[3 * j + 1], r = rot[3 * j + 2];
    const double cy = cos(y), sy = sin(y), cp = cos(p), sp = sin(p), cr = cos(r), sr = sin(r);
    // R = Rz(y)·Ry(p)·Rx(r)   (utils.py:425-453)
    s_R[j][0] = cy * cp; s_R[j][1] = cy * sp * sr - sy * cr; s_R[j][2] = cy * sp * cr + sy * sr;
    s_R[j][3] = sy * cp; s_R[j][4] = sy * sp * sr + cy * cr; s_R[j][5] = sy * sp * cr - cy * sr;
    s_R[j][6] = -sp;     s_R[j][7] = cp * sr;                s_R[j][8] = cp * cr;
  }
  __syncthreads();
  __shared__ int s_slot[PCL_GRID_MAX_ROT];
  if (j == 0) {                                                  // the greedy grouping is sequential (integer work only)
    int NG = 0;
    for (int a = 0; a < R; ++a) {
      int g = -1;
      for (int k = 0; k < NG && g < 0; ++k) {
        const int b = s_base[k];
        const double d0 = s_R[a][6] - s_R[b][6], d1 = s_R[a][7] - s_R[b][7], d2 = s_R[a][8] - s_R[b][8];
        if (fabs(d0) < PCL_GRID_TOL && fabs(d1) < PCL_GRID_TOL && fabs(d2) < PCL_GRID_TOL) g = k;
      }
      if (g < 0) { g = NG; s_base[NG++] = a; }
      s_group[a] = g;
    }
    plan->R = R; plan->NG = NG;
    int slot = 0;
    for (int g = 0; g < NG; ++g) {
      const int b = s_base[g];
      plan->g_start[g] = slot;
      plan->base_ypr[g][0] = rot[3 * b]; plan->base_ypr[g][1] = rot[3 * b + 1]; plan->base_ypr[g][2] = rot[3 * b + 2];
      for (int a = 0; a < R; ++a)
        if (s_group[a] == g) s_slot[a] = slot++;
      plan->g_count[g] = slot - plan->g_start[g];
    }
  }
  __syncthreads();
  if (j < R) {                                                   // one thread per rotation: its azimuth offset against the group's base
    const int a = j, b = s_base[s_group[a]];
    // M = R_a R_bᵀ = Rz(delta): M00 = row0(a)·row0(b), M10 = row1(a)·row0(b)
    const double m00 = s_R[a][0] * s_R[b][0] + s_R[a][1] * s_R[b][1] + s_R[a][2] * s_R[b][2];
    const double m10 = s_R[a][3] * s_R[b][0] + s_R[a][4] * s_R[b][1] + s_R[a][5] * s_R[b][2];
    plan->slot_rot[s_slot[a]] = a;
    plan->slot_delta[s_slot[a]] = (a == b) ? 0.0f : (float)atan2(m10, m00);
  }
}

// streamed point loads: read-only path, no L1 allocation, evict-first in L2 (C3: 22.3 -> 22.0 ms for 1 024 poses x 10 M points)
__device__ __forceinline__ float pcl_ld_stream(const float* p, const unsigned long long pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
template <int FMT, int KK, bool CHECK>
__device__ __forceinline__ void pcl_grid_rows(const PclCloudView& C, const PclImage& I, const PclPose* s_pose, const int nt,
                                              const PclGridPlan& s_plan, double (*s_acc)[PCL_MAX_POSE_BLOCK][2],
                                              const long long row0, const int tid, const int lane, const int warp) {
  const long long base = row0 * PCL_THREADS + tid;
  unsigned long long pol;                                        // the point stream is evict-first in L2: the texel table is what should stay
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  float px[KK], py[KK], pz[KK], cr[KK], cg[KK], cb[KK];
#pragma unroll
  for (int j = 0; j < KK; ++j) {
    const long long i = base + (long long)j * PCL_THREADS;
    px[j] = pcl_ld_stream(C.x + i, pol); py[j] = pcl_ld_stream(C.y + i, pol); pz[j] = pcl_ld_stream(C.z + i, pol);
    cr[j] = pcl_ld_stream(C.r + i, pol); cg[j] = pcl_ld_stream(C.g + i, pol); cb[j] = pcl_ld_stream(C.b + i, pol);
  }
  const int NG = s_plan.NG, R = s_plan.R;
  for (int ti = 0; ti < nt; ++ti) {
    for (int g = 0; g < NG; ++g) {
      const PclPose pose = s_pose[ti * NG + g];
      PclGridBase b[KK];
#pragma unroll
      for (int j = 0; j < KK; ++j) pcl_grid_base(pose, I, px[j], py[j], pz[j], b[j]);
      const int s0 = s_plan.g_start[g], s1 = s0 + s_plan.g_count[g];
      // members four at a time: their (se, sm) pairs are reduced by ONE halving butterfly of 8 values (9 shuffles)
      // instead of four xor-trees of 6 shuffles each (round 2: 4.94 -> 4.74 ms on the C2 grid)
      for (int s = s0; s < s1; s += 4) {
        const int nm = min(4, s1 - s);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = -0.0f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (m < nm) {
            const float delta = s_plan.slot_delta[s + m];
#pragma unroll
            for (int j = 0; j < KK; ++j) {
              const bool valid = CHECK ? ((base + (long long)j * PCL_THREADS) < C.n) : true;
              pcl_grid_member<FMT>(I, b[j], delta, cr[j], cg[j], cb[j], valid, v[2 * m], v[2 * m + 1]);
            }
          }
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
                const float send = up ? v[i] : v[i + n];
                const float keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
              }
            }
          } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
          }
        }
        if ((lane & 3) == 0) {
          const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);      // value index this lane ends up with
          if ((k >> 1) < nm) s_acc[warp][ti * R + s + (k >> 1)][k & 1] += (double)v[0];
        }
      }
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(PCL_THREADS, 3)
pcl_grid_score_kernel(const PclCloudView C, const PclImage I, const float* __restrict__ trans, const int T, const int TB,
                      const PclGridPlan* __restrict__ plan, const long long n_rows, double* __restrict__ partial,
                      unsigned int* __restrict__ counters, float* __restrict__ loss, float* __restrict__ count, const int swap) {
  __shared__ PclGridPlan s_plan;
  __shared__ __align__(16) PclPose s_pose[PCL_MAX_POSE_BLOCK];
  __shared__ double s_acc[PCL_WARPS][PCL_MAX_POSE_BLOCK][2];
  __shared__ double2 s_sum[PCL_THREADS];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // block order: translation blocks fastest (swap) keeps the co-resident CTAs on few row ranges = one part of the
  // panorama per rotation, so the live slice of the texel table fits L2
  const unsigned int bt = swap ? blockIdx.x : blockIdx.y, br = swap ? blockIdx.y : blockIdx.x;
  const unsigned int n_ranges = swap ? gridDim.y : gridDim.x;
  const int t0 = bt * TB;
  const int nt = min(TB, T - t0);

  for (int i = tid; i < PCL_WARPS * PCL_MAX_POSE_BLOCK * 2; i += PCL_THREADS) (&s_acc[0][0][0])[i] = 0.0;
  for (int i = tid; i < (int)(sizeof(PclGridPlan) / 4); i += PCL_THREADS)
    reinterpret_cast<int*>(&s_plan)[i] = __ldg(reinterpret_cast<const int*>(plan) + i);
  __syncthreads();
  const int NG = s_plan.NG, R = s_plan.R;
  if (tid < nt * NG) {
    const int ti = tid / NG, g = tid - ti * NG;
    float p6[6];
    p6[0] = __ldg(trans + 3 * (size_t)(t0 + ti)); p6[1] = __ldg(trans + 3 * (size_t)(t0 + ti) + 1); p6[2] = __ldg(trans + 3 * (size_t)(t0 + ti) + 2);
    p6[3] = s_plan.base_ypr[g][0]; p6[4] = s_plan.base_ypr[g][1]; p6[5] = s_plan.base_ypr[g][2];
    pcl_pose_from_params(p6, s_pose[tid]);
  }
  __syncthreads();

  const long long r_begin = n_rows * (long long)br / (long long)n_ranges;
  const long long r_end = n_rows * (long long)(br + 1) / (long long)n_ranges;
  long long r = r_begin;
  const long long r_full = min(r_end, C.n / PCL_THREADS);
  {
    const long long n = r_full - r, a = n >> 2, b = n & 3;
    long long n5 = (a >= b) ? b : 0, n4 = (a >= b) ? a - b : a;
    for (; n5 > 0; --n5, r += 5) pcl_grid_rows<FMT, 5, false>(C, I, s_pose, nt, s_plan, s_acc, r, tid, lane, warp);
    for (; n4 > 0; --n4, r += 4) pcl_grid_rows<FMT, 4, false>(C, I, s_pose, nt, s_plan, s_acc, r, tid, lane, warp);
  }
  for (; r < r_full; ++r) pcl_grid_rows<FMT, 1, false>(C, I, s_pose, nt, s_plan, s_acc, r, tid, lane, warp);
  for (; r < r_end; ++r) pcl_grid_rows<FMT, 1, true>(C, I, s_pose, nt, s_plan, s_acc, r, tid, lane, warp);
  __syncthreads();

  // CTA partial record: partial[blockIdx.x][translation][slot] = {Σ m e, Σ m}
  const int np = nt * R;
  const size_t P = (size_t)T * (size_t)R;
  const size_t p0 = (size_t)t0 * (size_t)R;
  for (int i = tid; i < np; i += PCL_THREADS) {
    double2 t = make_double2(0.0, 0.0);
#pragma unroll
    for (int w = 0; w < PCL_WARPS; ++w) { t.x += s_acc[w][i][0]; t.y += s_acc[w][i][1]; }
    reinterpret_cast<double2*>(partial)[(size_t)br * P + p0 + i] = t;
  }

  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&counters[bt], 1u);
    s_last = (ticket == n_ranges - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // deterministic two-level reduction of the gridDim.x records (fixed order), as in pcl_sample_kernel
  const int G = max(1, PCL_THREADS / np);
  {
    const int item = tid % np, g = tid / np;
    double2 t = make_double2(0.0, 0.0);
    if (g < G) {
      const double2* src = reinterpret_cast<const double2*>(partial) + p0 + item;
#pragma unroll 16
      for (unsigned int bx = g; bx < n_ranges; bx += G) {
        const double2 v = __ldcg(src + (size_t)bx * P);
        t.x += v.x; t.y += v.y;
      }
    }
    s_sum[tid] = t;
  }
  __syncthreads();
  if (tid < np) {
    double2 t = make_double2(0.0, 0.0);
    for (int g = 0; g < G; ++g) { const double2 v = s_sum[g * np + tid]; t.x += v.x; t.y += v.y; }
    const int ti = tid / R, s = tid - ti * R;
    const size_t out = (size_t)(t0 + ti) * (size_t)R + (size_t)s_plan.slot_rot[s];     // pose index i*R + j (utils.py:484-485)
    loss[out] = (float)(t.x / t.y);                   // 0/0 -> NaN, the reference's empty mean
    if (count) count[out] = (float)t.y;
  }
  if (tid == 0) counters[bt] = 0u;
}

// generic fallback for rotation lists longer than a CTA's pose block: expand to (T·R, 6) poses
__global__ void pcl_grid_expand_kernel(const float* __restrict__ trans, const float* __restrict__ rot, const long long T,
                                       const int R, float* __restrict__ poses6) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= T * R) return;
  const long long i = p / R;
  const int j = (int)(p - i * R);
#pragma unroll
  for (int k = 0; k < 3; ++k) { poses6[6 * p + k] = trans[3 * i + k]; poses6[6 * p + 3 + k] = rot[3 * j + k]; }
}

template <int FMT>
static cudaError_t pcl_grid_launch_fmt(dim3 grid, cudaStream_t st, const PclCloudView& C, const PclImage& I, const float* trans, int T,
                                       int TB, const PclGridPlan* plan, long long n_rows, double* partial, unsigned int* counters,
                                       float* loss, float* count, int swap) {
  pcl_grid_score_kernel<FMT><<<grid, PCL_THREADS, 0, st>>>(C, I, trans, T, TB, plan, n_rows, partial, counters, loss, count, swap);
  return cudaGetLastError();
}

extern "C" int pcl_score_grid(const pcl_cloud* c, const pcl_image* im, const float* trans_t3_dev, int64_t t,
                              const float* rot_r3_dev, int r, float* loss_tr_dev, float* count_tr_dev, void* stream) {
  if (!c || !im || !trans_t3_dev || !rot_r3_dev || !loss_tr_dev) { pcl_set_error("null handle or pointer"); return PCL_ERR_INVALID; }
  if (t <= 0 || r <= 0 || t * (int64_t)r > 65535ll * PCL_MAX_POSE_BLOCK) {
    pcl_set_error("grid %lld x %d out of range", (long long)t, r);
    return PCL_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PclUseGuard guard{c, im, st};
  if (r > PCL_GRID_MAX_ROT) {
    float* poses = nullptr;
    const long long P = (long long)t * r;
    PCL_CUDA(pcl_pool_alloc((void**)&poses, (size_t)P * 6 * sizeof(float), st));
    pcl_grid_expand_kernel<<<(unsigned int)((P + 255) / 256), 256, 0, st>>>(trans_t3_dev, rot_r3_dev, t, r, poses);
    g_pcl_launches.fetch_add(1);
    int rc = (cudaGetLastError() == cudaSuccess) ? pcl_score(c, im, poses, P, loss_tr_dev, count_tr_dev, stream) : PCL_ERR_CUDA;
    if (rc == PCL_ERR_CUDA) pcl_set_error("pose expansion launch failed");
    pcl_pool_free(poses, st);
    return rc;
  }
  const int TB = PCL_MAX_POSE_BLOCK / r;                  // translations per CTA: TB·R <= 32 poses
  const int gy = (int)((t + TB - 1) / TB);
  const long long n_rows = (c->n + PCL_THREADS - 1) / PCL_THREADS;      // rows with real points only (see pcl_plan)
  const int resident = pcl_num_sms() * 3;
  long long gx = 1;
  double best = -1.0;
  for (int w = 1; w <= PCL_GRID_MAX_WAVES; ++w) {         // whole resident waves, the best-filled count
    long long g = (long long)resident * w / gy;
    if (g < 1) g = 1;
    if (g > n_rows) g = n_rows;
    const double util = (double)(g * gy) / (double)((g * gy + resident - 1) / resident * resident);
    if (util > best + 1e-9) { best = util; gx = g; }
  }
  const size_t P = (size_t)t * (size_t)r;
  const size_t pbytes = (size_t)gx * P * 2 * sizeof(double);
  const size_t cbytes = ((size_t)gy * sizeof(unsigned int) + 15) & ~(size_t)15;
  char* block = nullptr;
  PCL_CUDA(pcl_pool_alloc((void**)&block, pbytes + cbytes + sizeof(PclGridPlan), st));
  double* partial = reinterpret_cast<double*>(block);
  unsigned int* counters = reinterpret_cast<unsigned int*>(block + pbytes);
  PclGridPlan* plan = reinterpret_cast<PclGridPlan*>(block + pbytes + cbytes);
  cudaError_t e = cudaMemsetAsync(counters, 0, (size_t)gy * sizeof(unsigned int), st);
  if (e == cudaSuccess) {
    pcl_grid_plan_kernel<<<1, PCL_GRID_MAX_ROT, 0, st>>>(rot_r3_dev, r, plan);
    g_pcl_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    PclCloudView C = {c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
    const PclImage& I = im->view;
    const int swap = pcl_opt(PCL_OPT_GRID_SWAP) && gx <= 65535;
    const dim3 grid = swap ? dim3((unsigned int)gy, (unsigned int)gx) : dim3((unsigned int)gx, (unsigned int)gy);
    switch (I.fmt) {
      case PCL_FMT_U8Q: e = pcl_grid_launch_fmt<PCL_FMT_U8Q>(grid, st, C, I, trans_t3_dev, (int)t, TB, plan, n_rows, partial, counters, loss_tr_dev, count_tr_dev, swap); break;
      case PCL_FMT_U8P: e = pcl_grid_launch_fmt<PCL_FMT_U8P>(grid, st, C, I, trans_t3_dev, (int)t, TB, plan, n_rows, partial, counters, loss_tr_dev, count_tr_dev, swap); break;
      case PCL_FMT_F32: e = pcl_grid_launch_fmt<PCL_FMT_F32>(grid, st, C, I, trans_t3_dev, (int)t, TB, plan, n_rows, partial, counters, loss_tr_dev, count_tr_dev, swap); break;
      case PCL_FMT_TEX: e = pcl_grid_launch_fmt<PCL_FMT_TEX>(grid, st, C, I, trans_t3_dev, (int)t, TB, plan, n_rows, partial, counters, loss_tr_dev, count_tr_dev, swap); break;
      case PCL_FMT_F16D: e = pcl_grid_launch_fmt<PCL_FMT_F16D>(grid, st, C, I, trans_t3_dev, (int)t, TB, plan, n_rows, partial, counters, loss_tr_dev, count_tr_dev, swap); break;
      default: e = cudaErrorInvalidValue; break;
    }
    g_pcl_launches.fetch_add(1);
  }
  pcl_pool_free(block, st);
  if (e != cudaSuccess) { pcl_set_error("structured grid scoring failed: %s", cudaGetErrorString(e)); return PCL_ERR_CUDA; }
  return PCL_OK;
}
