// C ABI of the fused refinement (include/piccolo_b200.h: pcl_refine_*): handle, reset / read, and the run entry
// that picks the path:
//   B <= 16  ->  pcl_refine.cuh: ALL iterations in one cooperative launch (split-phase grid barrier), or — when the
//                device refuses the cooperative launch / PERSIST=0 — one launch per iteration with the same arithmetic
//   B  > 16  ->  pcl_sampling.cu: one launch of the generic fwd+bwd kernel per iteration
// and, with a pcl_comm, the point-sharded run over several GPUs (records exchanged by peer stores over NVLink).
//
// Replaces omniloc.py:44-58 (`omniloc`) and :249-269 (`omniloc_batch`).
#include "pcl_refine.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int pcl_generic_refine_iters(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, cudaStream_t st);

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
  if (!r) { pcl_set_error("out of host memory"); return PCL_ERR_INVALID; }
  r->B = b; r->lr0 = lr; r->factor = factor; r->patience = patience; r->batch_semantics = batch_semantics ? 1 : 0;
  *out = r;                          // device storage is allocated by the first pcl_refine_reset, on its stream
  return PCL_OK;
}

static size_t pcl_align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int pcl_refine_reset(pcl_refine* r, const float* poses_b6_dev, void* stream) {
  if (!r || !poses_b6_dev) { pcl_set_error("null refine handle or poses"); return PCL_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t b = (size_t)r->B;
  const size_t o_eval = pcl_align256(sizeof(PclRefineState) * b);
  const size_t o_loss = o_eval + pcl_align256(sizeof(float) * 6 * b);
  const size_t o_cnt = o_loss + pcl_align256(sizeof(float) * b);
  const size_t o_arr = o_cnt + pcl_align256(sizeof(unsigned int) * b);
  const size_t o_ready = o_arr + 256;
  const size_t o_tick = o_ready + 256;
  const size_t o_pose = o_tick + 256;
  const size_t total = o_pose + sizeof(unsigned long long) * 12 * PCL_RF_MAXB;
  if (!r->block) {
    PCL_CUDA(pcl_pool_alloc((void**)&r->block, total, st));
    r->owner = st;
    r->state = (PclRefineState*)r->block;
    r->evalp = (float*)(r->block + o_eval);
    r->loss = (float*)(r->block + o_loss);
    r->counters = (unsigned int*)(r->block + o_cnt);
    r->arrive = (unsigned int*)(r->block + o_arr);
    r->ready = (unsigned int*)(r->block + o_ready);
    r->tickets = (unsigned int*)(r->block + o_tick);
    r->posebuf = (unsigned long long*)(r->block + o_pose);
    PCL_CUDA(cudaMemsetAsync(r->block + o_cnt, 0, total - o_cnt, st));    // tickets are self-resetting, arrivals / ready flags monotonic
    memset(r->arrive_base, 0, sizeof(r->arrive_base));
    memset(r->ready_base, 0, sizeof(r->ready_base));
  }
  r->steps_done = 0;
  pcl_refine_reset_kernel<<<(r->B + 127) / 128, 128, 0, st>>>(r->state, r->evalp, poses_b6_dev, r->B, r->lr0);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

// ------------------------------------------------------------------------------------------------
// fused path
// ------------------------------------------------------------------------------------------------
static void pcl_rf_blocks(int B, int* nblk, int* npb) {
  // two pose blocks at least (the split-phase barrier hides one block's barrier behind the other's work);
  // at most PCL_RF_MAXNPB candidates per block (their sums live in registers)
  int nb = B == 1 ? 1 : (B + PCL_RF_MAXNPB - 1) / PCL_RF_MAXNPB;
  if (B > 1 && nb < 2) nb = 2;
  int pb = (B + nb - 1) / nb;
  const int force = pcl_opt(PCL_OPT_RF_NPB);
  if (force >= 1 && force <= PCL_RF_MAXNPB && (B + force - 1) / force <= PCL_RF_MAXBLK) pb = force;
  *npb = pb;
  *nblk = (B + pb - 1) / pb;
}

static cudaError_t pcl_rf_dispatch_persistent(int fmt, const PclRfParams& ps, cudaStream_t st) {
  switch (fmt) {
    case PCL_FMT_U8Q: return pcl_rf_launch_persistent<PCL_FMT_U8Q>(ps, st);
    case PCL_FMT_U8P: return pcl_rf_launch_persistent<PCL_FMT_U8P>(ps, st);
    case PCL_FMT_F32: return pcl_rf_launch_persistent<PCL_FMT_F32>(ps, st);
    case PCL_FMT_TEX: return pcl_rf_launch_persistent<PCL_FMT_TEX>(ps, st);
    case PCL_FMT_F16D: return pcl_rf_launch_persistent<PCL_FMT_F16D>(ps, st);
    default: return cudaErrorInvalidValue;
  }
}
static cudaError_t pcl_rf_dispatch_iter(int fmt, const PclRfParams& ps, unsigned int* tickets, double bc1, double bc2s, bool pdl, cudaStream_t st) {
  switch (fmt) {
    case PCL_FMT_U8Q: return pcl_rf_launch_iter<PCL_FMT_U8Q>(ps, tickets, bc1, bc2s, pdl, st);
    case PCL_FMT_U8P: return pcl_rf_launch_iter<PCL_FMT_U8P>(ps, tickets, bc1, bc2s, pdl, st);
    case PCL_FMT_F32: return pcl_rf_launch_iter<PCL_FMT_F32>(ps, tickets, bc1, bc2s, pdl, st);
    case PCL_FMT_TEX: return pcl_rf_launch_iter<PCL_FMT_TEX>(ps, tickets, bc1, bc2s, pdl, st);
    case PCL_FMT_F16D: return pcl_rf_launch_iter<PCL_FMT_F16D>(ps, tickets, bc1, bc2s, pdl, st);
    default: return cudaErrorInvalidValue;
  }
}

static int pcl_rf_grow(void** p, size_t* cap, size_t need, size_t elem, cudaStream_t st) {
  if (need <= *cap) return PCL_OK;
  pcl_pool_free(*p, st);
  *p = nullptr; *cap = 0;                                   // a failed allocation leaves no dangling pointer behind
  PCL_CUDA(pcl_pool_alloc(p, need * elem, st));
  *cap = need;
  return PCL_OK;
}

// the small-batch run; comm == nullptr: single GPU
static int pcl_refine_run_fused(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, pcl_comm* comm, cudaStream_t st) {
  const int nranks = comm ? comm->nranks : 1, rank = comm ? comm->rank : 0;
  PclRfParams ps;
  memset(&ps, 0, sizeof(ps));
  ps.C = PclCloudView{c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
  ps.B = r->B;
  pcl_rf_blocks(r->B, &ps.nblk, &ps.npb);
  ps.rank = rank; ps.nranks = nranks;

  // one resident wave of 2 CTAs per SM = G compute CTAs + the service CTA; small shards get one CTA per 256 points
  ps.p_begin = (long long)c->n * rank / nranks;
  ps.p_end = (long long)c->n * (rank + 1) / nranks;
  const long long n_pts = ps.p_end - ps.p_begin;
  long long G = (long long)PCL_RF_CTAS_PER_SM * pcl_num_sms() - 1;
  const long long by_rows = (n_pts + PCL_RF_THREADS - 1) / PCL_RF_THREADS;
  if (G > by_rows) G = by_rows < 1 ? 1 : by_rows;
  if (comm) G = (long long)PCL_RF_CTAS_PER_SM * pcl_num_sms() - 1;                   // every rank must use the same G: the record slots are rank*G + cta
  ps.G = (int)G;
  bool fully_resident = false;
  {
    // points a CTA keeps in shared memory for the whole run: its whole range when that fits beside the kernel's static
    // arrays with >= 60 KB left as L1 (14 rows of 512 points = 1.05 M points per GPU), else whole 2048-point groups of it
    // (measured: 3 M points 125 -> 120 us per iteration, 5 M 206 -> 201, 10 M 399 -> 396; the rest streams evict-first).
    // Option RF_RES: -1/1 = this rule, 0 = off.
    int dev = 0, optin = 0;
    PCL_CUDA(cudaGetDevice(&dev));
    PCL_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // keep >= 60 KB of the 256 KB array as L1 for the texel gather: 28 KB (the next carve-out step) was measured slower
    // than not keeping the points at all
    const long long budget = (long long)(optin < 196 * 1024 ? optin : 196 * 1024) - PCL_RF_SMEM_STATIC;
    const long long cap = budget / (PCL_RF_ROWF * (long long)sizeof(float)) * PCL_RF_THREADS;   // whole rows of 512 points
    const long long per_cta = (n_pts + G - 1) / G;
    long long res = per_cta <= cap ? per_cta : cap / PCL_RF_GROUP * PCL_RF_GROUP;
    if (pcl_opt(PCL_OPT_RF_RES) == 0 || cap <= 0) res = 0;
    fully_resident = res > 0 && res >= per_cta;
    ps.res_pts = (int)res;
  }
  // Texel table.  The poses of a small batch move every iteration, so the gather lands at unpredictable places of the
  // table: it has to stay in L2.  When the cloud streams through L2 as well, only the compact companion table (U8Q,
  // 16 B per footprint) survives next to it; when the points are resident in shared memory the cloud leaves L2 to the
  // table and the fp16-basis table (F16D: no unpacking, 161 instead of 175 instructions per evaluation) is the faster one
  // (C2: 38.7 vs 40.9 us per iteration).  Option SMALL_TABLE: 1 = this rule, 0 = always the main table, 2 = always the companion.
  const PclImage& view = (im->has_small && pcl_opt(PCL_OPT_SMALL_TABLE) && !(fully_resident && pcl_opt(PCL_OPT_SMALL_TABLE) != 2)) ? im->view_small : im->view;
  ps.I = view;
  ps.num_iter = num_iter;
  ps.state = r->state; ps.evalp = r->evalp; ps.loss = r->loss;
  ps.box = c->lo_hi_dev;
  ps.factor = r->factor; ps.patience = r->patience; ps.batch_semantics = r->batch_semantics;

  const size_t rec_blk = (size_t)G * nranks * PCL_RF_MAXNPB * PCL_NSUM;
  const size_t rec_need = 2 * (size_t)ps.nblk * rec_blk;
  if (comm) {
    if (rec_need * sizeof(double) > comm->rec_bytes) { pcl_set_error("pcl_comm window too small for %d ranks", nranks); return PCL_ERR_INVALID; }
    for (int k = 0; k < nranks; ++k) { ps.rec[k] = (double*)(comm->peer[k] + comm->rec_off); ps.arrive[k] = (unsigned int*)(comm->peer[k] + PCL_COMM_OFF_ARRIVE); }
    memcpy(ps.arrive_base, comm->arrive_base, sizeof(ps.arrive_base));
  } else {
    int rc = pcl_rf_grow((void**)&r->rec, &r->rec_doubles, rec_need, sizeof(double), st);
    if (rc) return rc;
    ps.rec[0] = r->rec; ps.arrive[0] = r->arrive;
    memcpy(ps.arrive_base, r->arrive_base, sizeof(ps.arrive_base));
  }
  ps.parity0 = (int)(r->steps_done & 1);
  ps.posebuf = r->posebuf;
  memcpy(ps.ready_base, r->ready_base, sizeof(ps.ready_base));

  if (num_iter >= 1 && (comm || pcl_opt(PCL_OPT_PERSIST) != 0)) {
    static int coop = -1;
    if (coop < 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) coop = 0;
    }
    if (coop) {
      int rc = pcl_rf_grow((void**)&r->bc_dev, &r->bc_cap, (size_t)num_iter * 2, sizeof(double), st);
      if (rc) return rc;
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
      ps.bc = r->bc_dev;
      if (pcl_opt(PCL_OPT_RF_DEBUG)) {
        rc = pcl_rf_grow((void**)&r->dbg, &r->dbg_cap, (size_t)(G + 1) * 4 + (size_t)num_iter, sizeof(unsigned long long), st);   // CTA rows, then one wall-clock stamp per iteration
        if (rc) return rc;
        ps.dbg = r->dbg; r->dbg_ctas = (int)G + 1; r->dbg_iters = num_iter;
      }
      e = pcl_rf_dispatch_persistent(view.fmt, ps, st);
      if (e == cudaSuccess) {
        g_pcl_launches.fetch_add(1);
        const unsigned int adv = (unsigned int)num_iter * (unsigned int)(G * nranks);
        for (int b = 0; b < ps.nblk; ++b) {
          if (comm) comm->arrive_base[b] += adv; else r->arrive_base[b] += adv;
          r->ready_base[b] += (unsigned int)num_iter;
        }
        r->steps_done += num_iter;
        return PCL_OK;
      }
      (void)cudaGetLastError();
      if (comm || (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorNotSupported)) {
        pcl_set_error("persistent refinement launch failed: %s", cudaGetErrorString(e));
        return PCL_ERR_CUDA;
      }
    } else if (comm) {
      pcl_set_error("the sharded refinement needs cooperative launch support");
      return PCL_ERR_CUDA;
    }
  }
  // per-iteration launches (the device refused the cooperative launch, or PERSIST=0): same arithmetic
  for (int it = 0; it < num_iter; ++it) {
    r->steps_done += 1;
    const double bc1 = 1.0 - pow(0.9, (double)r->steps_done), bc2s = sqrt(1.0 - pow(0.999, (double)r->steps_done));
    cudaError_t e = pcl_rf_dispatch_iter(view.fmt, ps, r->tickets, bc1, bc2s, it > 0 && pcl_opt(PCL_OPT_PDL) != 0, st);
    g_pcl_launches.fetch_add(1);
    PCL_CUDA(e);
  }
  return PCL_OK;
}

static int pcl_refine_check(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter) {
  if (!r || !r->block) { pcl_set_error("refine handle is null or was never reset"); return PCL_ERR_INVALID; }
  if (!c || !im) { pcl_set_error("null cloud or image handle"); return PCL_ERR_INVALID; }
  if (num_iter < 0) { pcl_set_error("negative iteration count"); return PCL_ERR_INVALID; }
  return PCL_OK;
}

extern "C" int pcl_refine_run(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, void* stream) {
  int rc = pcl_refine_check(r, c, im, num_iter);
  if (rc) return rc;
  PclUseGuard guard{c, im, (cudaStream_t)stream};
  if (r->B <= PCL_RF_MAXB) return pcl_refine_run_fused(r, c, im, num_iter, nullptr, (cudaStream_t)stream);
  return pcl_generic_refine_iters(r, c, im, num_iter, (cudaStream_t)stream);
}

extern "C" int pcl_refine_run_sharded(pcl_refine* r, const pcl_cloud* c, const pcl_image* im, int num_iter, pcl_comm* comm, void* stream) {
  int rc = pcl_refine_check(r, c, im, num_iter);
  if (rc) return rc;
  PclUseGuard guard{c, im, (cudaStream_t)stream};
  if (!comm || !comm->connected) { pcl_set_error("pcl_comm is null or not connected"); return PCL_ERR_INVALID; }
  if (r->B > PCL_RF_MAXB) { pcl_set_error("the point-sharded refinement handles up to %d candidates", PCL_RF_MAXB); return PCL_ERR_INVALID; }
  if (comm->nranks == 1) return pcl_refine_run_fused(r, c, im, num_iter, nullptr, (cudaStream_t)stream);
  // quiesce: no rank may still be reading records of an earlier run (another refiner / another layout) of this window
  rc = pcl_comm_barrier(comm, stream);
  if (rc) return rc;
  return pcl_refine_run_fused(r, c, im, num_iter, comm, (cudaStream_t)stream);
}

extern "C" int pcl_refine_read(const pcl_refine* r, float* pose_b6_dev, float* param_b6_dev, float* loss_b_dev,
                               double* lr_b_dev, void* stream) {
  if (!r || !r->block) { pcl_set_error("refine handle is null or was never reset"); return PCL_ERR_INVALID; }
  pcl_refine_read_kernel<<<(r->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->state, r->evalp, r->B, r->batch_semantics,
                                                                                pose_b6_dev, param_b6_dev, loss_b_dev, lr_b_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

// Option RF_DEBUG: counters of the last persistent run, 4 per CTA: compute CTAs {cycles of warp 0 in phases, cycles waiting
// for poses, phases whose poses were already there, 0}; the last row is the service CTA {cycles
// waiting for records, cycles reducing + stepping, 0, 0}.  Returns the number of rows (0 when nothing was recorded).  Blocks on `stream`.
extern "C" int pcl_refine_debug_stats(const pcl_refine* r, unsigned long long* out_host, int max_ctas, void* stream) {
  if (!r || !out_host) { pcl_set_error("null refine handle or output"); return PCL_ERR_INVALID; }
  if (!r->dbg || r->dbg_ctas <= 0) return 0;
  const int n = r->dbg_ctas < max_ctas ? r->dbg_ctas : max_ctas;
  PCL_CUDA(cudaMemcpyAsync(out_host, r->dbg, sizeof(unsigned long long) * 4 * (size_t)n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PCL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return n;
}

// Option RF_DEBUG: the service CTA's wall-clock stamp (%globaltimer, ns) at the end of every iteration of the last persistent
// run.  Returns the number of iterations copied (0 when nothing was recorded).  Blocks on `stream`.
extern "C" int pcl_refine_debug_timeline(const pcl_refine* r, unsigned long long* out_ns_host, int max_iters, void* stream) {
  if (!r || !out_ns_host) { pcl_set_error("null refine handle or output"); return PCL_ERR_INVALID; }
  if (!r->dbg || r->dbg_ctas <= 0 || r->dbg_iters <= 0) return 0;
  const int n = r->dbg_iters < max_iters ? r->dbg_iters : max_iters;
  PCL_CUDA(cudaMemcpyAsync(out_ns_host, r->dbg + 4 * (size_t)r->dbg_ctas, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PCL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return n;
}

extern "C" void pcl_refine_destroy(pcl_refine* r) {
  if (!r) return;
  pcl_pool_free(r->block, r->owner);
  pcl_pool_free(r->partial, r->owner);
  pcl_pool_free(r->rec, r->owner);
  pcl_pool_free(r->bc_dev, r->owner);
  pcl_pool_free(r->dbg, r->owner);
  free(r);
}
