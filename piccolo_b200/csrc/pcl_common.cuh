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

struct pcl_cloud {
  cudaStream_t owner;           // stream the storage is ordered on (freed there)
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
  char* block;                  // one allocation: state | evalp | loss | counters
  int B, patience, batch_semantics;
  long long steps_done;         // Adam step count so far (identical for all candidates)
  double lr0, factor;
  PclRefineState* state;        // [B]
  float* evalp;                 // [B][6] pose evaluated by the next forward
  double* partial;              // grow-only scratch for per-CTA partial sums
  size_t partial_floats;
  unsigned int* counters;       // last-block-done tickets
  float* loss;                  // [B]
  double* bc_dev;               // grow-only: per-iteration Adam bias corrections of a persistent run, [num_iter][2]
  size_t bc_cap;
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
