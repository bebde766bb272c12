// Shared declarations of the piccolo_b200 CUDA library (host side + handle layouts).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/piccolo_b200.h"
#include "pcl_eval.cuh"

#define PCL_THREADS 256
#define PCL_WARPS (PCL_THREADS / 32)
#define PCL_TILE_ALIGN 2048       // clouds are padded to a multiple of this many points
#define PCL_MAX_POSE_BLOCK 32     // poses a CTA keeps in shared memory
#define PCL_NSUM 8                // {Σme, Σm, a(3), τ(3)}

void pcl_set_error(const char* fmt, ...);
// stream-ordered allocations from the device's default memory pool (release threshold raised once so that freed
// blocks stay cached: creating a cloud / image / refiner per query costs no driver allocation and no device sync)
cudaError_t pcl_pool_alloc(void** p, size_t bytes, cudaStream_t st);
void pcl_pool_free(void* p, cudaStream_t st);
extern std::atomic<long long> g_pcl_launches;
int pcl_num_sms();
// process-global tuning knobs (pcl_set_option / PCL_<NAME> in the environment, read once)
enum { PCL_OPT_PERSIST = 0, PCL_OPT_PDL, PCL_OPT_PB_FWD, PCL_OPT_PB_BWD, PCL_OPT_WAVES, PCL_OPT_SWAP, PCL_OPT_GRID_SWAP, PCL_OPT_SMALL_TABLE,
       PCL_OPT_RF_NPB, PCL_OPT_RF_DEBUG, PCL_OPT_RF_RES, PCL_OPT_COUNT };
int pcl_opt(int id);

#define PCL_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      pcl_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PCL_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define PCL_LAUNCH_CHECK()                      \
  do {                                          \
    g_pcl_launches.fetch_add(1);                \
    PCL_CUDA(cudaGetLastError());               \
  } while (0)

// Handles are built and freed in stream order on their creation stream (`owner`).  Compute entries may run on any other
// stream: each records its last use there (PclUseGuard), and the destroy makes the owner stream wait for that record
// before the storage goes back to the pool.  (One record per handle: with several foreign streams the most recent one
// is waited for; order the others after it yourself.)
struct PclUse {
  cudaEvent_t ev;
  int cross;                    // 1: used on a stream other than the owner since creation
};

struct pcl_cloud {
  cudaStream_t owner;           // stream the storage is ordered on (freed there)
  PclUse use;
  float* block;                 // one allocation: x|y|z|r|g|b, each n_pad floats
  float *x, *y, *z, *r, *g, *b;
  int64_t n, n_pad;
  float* lo_hi_dev;             // clamp box {x,y,z min, x,y,z max} on the device (read by the refine finalize)
  float lo_hi[6];               // host copy, fetched lazily by pcl_cloud_bounds
  int lo_hi_valid;
  int order;
};

struct pcl_image {
  cudaStream_t owner;
  PclUse use;
  void* data;
  size_t bytes;
  PclImage view;                // view.data == data
  // Optional compact second table (U8Q, 16 B per footprint; U8P, 4 B per texel, for panoramas above 1024x2048) used by small refinement batches: their poses move
  // every iteration, and the 32 B/footprint F16D table (67 MB at 1024x2048) does not stay L2-resident beside the
  // cloud under that access pattern (measured 55-57 vs 49 us per iteration), while it wins for scoring.
  void* data_small;
  PclImage view_small;
  int has_small;
};

// per-candidate optimiser state (device)
struct PclRefineState {
  float m[6], v[6];             // Adam moments
  float param[6];               // Adam's parameter (translation clamped to the box)
  float last_loss;
  int step;
  int bad;
  int pad;
  double lr, best;
};

struct pcl_refine {
  cudaStream_t owner;
  char* block;                  // one allocation: state | evalp | loss | counters | arrive | tickets
  int B, patience, batch_semantics;
  long long steps_done;         // Adam step count so far (identical for all candidates)
  double lr0, factor;
  PclRefineState* state;        // [B]
  float* evalp;                 // [B][6] pose evaluated by the next forward
  float* loss;                  // [B]
  // generic path (B > 16): one launch of pcl_sample_kernel per iteration
  double* partial;              // grow-only scratch for per-CTA partial sums
  size_t partial_floats;
  unsigned int* counters;       // last-block-done tickets, [B]
  // fused path (B <= 16): pcl_refine.cuh
  double* rec;                  // grow-only: per-CTA records [2][nblk][G][32]
  size_t rec_doubles;
  unsigned int* arrive;         // [PCL_RF_MAXBLK] monotonic arrival counters of the split-phase barrier
  unsigned int arrive_base[8];  // their current values (host copy; a run advances the counters of ITS pose blocks only)
  unsigned int* ready;          // [PCL_RF_MAXBLK] monotonic "poses published" flags of the service CTA
  unsigned int ready_base[8];
  unsigned long long* posebuf;  // [min(B,16)][12] published poses as {tag : 32 | float bits : 32} words
  unsigned int* tickets;        // [PCL_RF_MAXBLK] last-block-done tickets of the per-iteration fallback
  double* bc_dev;               // grow-only: per-iteration Adam bias corrections of a persistent run, [num_iter][2]
  size_t bc_cap;
  unsigned long long* dbg;      // option RF_DEBUG: per compute CTA cycle counters of the last persistent run
  size_t dbg_cap;
  int dbg_ctas, dbg_iters;
};

// Peer-memory window of one rank (pcl_comm.cu): cudaMalloc'ed, IPC-mapped into every other rank of the box.
// Layout: [0) barrier counter | [64) refinement arrival counters | [128) all-gather counter | [1024) all-gather
// slots [2][nranks][PCL_COMM_AG_MAX] floats | [rec_off) refinement records.
#define PCL_COMM_OFF_BAR 0
#define PCL_COMM_OFF_ARRIVE 64
#define PCL_COMM_OFF_AG 128
#define PCL_COMM_OFF_AGDATA 1024
#define PCL_COMM_AG_MAX 16384
#define PCL_COMM_MAXRANKS 8
struct pcl_comm {
  int rank, nranks, connected, device;
  size_t bytes, rec_off, rec_bytes;
  char* peer[PCL_COMM_MAXRANKS];        // peer[rank] is the local window
  unsigned int bar_epoch, ag_epoch;     // host copies of the monotonic counters (identical call sequences on all ranks)
  unsigned int arrive_base[8];
};

static inline void pcl_note_use(cudaStream_t owner, const PclUse* use_, cudaStream_t st) {
  if (st == owner) return;
  PclUse* u = const_cast<PclUse*>(use_);
  if (!u->ev && cudaEventCreateWithFlags(&u->ev, cudaEventDisableTiming) != cudaSuccess) { u->ev = nullptr; return; }
  if (cudaEventRecord(u->ev, st) == cudaSuccess) u->cross = 1;
}
// before freeing a handle's storage on its owner stream
static inline void pcl_use_release(cudaStream_t owner, PclUse* u) {
  if (u->ev) {
    if (u->cross) cudaStreamWaitEvent(owner, u->ev, 0);
    cudaEventDestroy(u->ev);
    u->ev = nullptr;
  }
}
// declared at the top of a compute entry: records the use when the entry returns (after everything is enqueued)
struct PclUseGuard {
  const pcl_cloud* c;
  const pcl_image* im;
  cudaStream_t st;
  ~PclUseGuard() {
    if (c) pcl_note_use(c->owner, &c->use, st);
    if (im) pcl_note_use(im->owner, &im->use, st);
  }
};

struct PclCloudView {
  const float *x, *y, *z, *r, *g, *b;
  long long n;
};

// what the last CTA of a pose block does with the reduced sums
enum { PCL_FIN_SCORE = 0, PCL_FIN_GRAD = 1, PCL_FIN_REFINE = 2 };

struct PclFinalize {
  int mode;
  float* loss;                  // [P]
  float* count;                 // [P] nullable
  float* grad;                  // [P][6]           (GRAD)
  PclRefineState* state;        // [P]              (REFINE)
  float* evalp;                 // [P][6] in/out    (REFINE)
  const float* box;             // clamp box {lo(3), hi(3)} on the device   (REFINE)
  double factor;
  double bc1, bc2_sqrt;         // Adam bias corrections of THIS iteration: 1-0.9^step, sqrt(1-0.999^step)
  int patience;
  int batch_semantics;
};
