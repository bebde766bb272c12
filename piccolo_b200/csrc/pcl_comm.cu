// Peer-memory communicator for the ranks of ONE box (one process per GPU): every rank owns a cudaMalloc'ed window
// that all other ranks map through CUDA IPC, so kernels exchange data with plain stores and system-scope atomics
// over NVLink / NVSwitch — no NCCL, no host round-trip, everything stream-ordered.
//
// Used by the point-sharded refinement (pcl_refine.cuh: per-CTA records and arrival counters written straight into
// every peer's window) and by the tiny all-gathers of the sharded pose search (per-pose losses after scoring,
// (loss, pose) rows before the arg-min) when the caller binds the C ABI without torch.distributed.
// The reference has no counterpart: it is single-device (localize.py:124).
#include "pcl_common.cuh"

#include <stdlib.h>
#include <string.h>

struct PclCommView {
  char* peer[PCL_COMM_MAXRANKS];
  int rank, nranks;
};

__device__ __forceinline__ unsigned int pcl_comm_ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one CTA; on completion every rank's earlier work on its stream has finished
__global__ void pcl_comm_barrier_kernel(const PclCommView v, const unsigned int target) {
  if ((int)threadIdx.x < v.nranks) {
    __threadfence_system();
    atomicAdd_system(reinterpret_cast<unsigned int*>(v.peer[threadIdx.x] + PCL_COMM_OFF_BAR), 1u);
  }
  if (threadIdx.x == 0) {
    const unsigned int* ctr = reinterpret_cast<const unsigned int*>(v.peer[v.rank] + PCL_COMM_OFF_BAR);
    while ((int)(pcl_comm_ld_acquire_sys(ctr) - target) < 0) { }
  }
}

// dst[k*n + i] = src of rank k.  Slots are double-buffered by call parity: a rank can only be two calls ahead of
// another after that one has finished reading the older slot.
__global__ void pcl_comm_allgather_kernel(const PclCommView v, const float* __restrict__ src, const int n, float* __restrict__ dst,
                                          const unsigned int target, const int parity) {
  const size_t slot = ((size_t)parity * v.nranks + v.rank) * PCL_COMM_AG_MAX;
  for (int k = 0; k < v.nranks; ++k) {
    float* out = reinterpret_cast<float*>(v.peer[k] + PCL_COMM_OFF_AGDATA) + slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = src[i];
  }
  __syncthreads();
  if ((int)threadIdx.x < v.nranks) {
    __threadfence_system();
    atomicAdd_system(reinterpret_cast<unsigned int*>(v.peer[threadIdx.x] + PCL_COMM_OFF_AG), 1u);
  }
  if (threadIdx.x == 0) {
    const unsigned int* ctr = reinterpret_cast<const unsigned int*>(v.peer[v.rank] + PCL_COMM_OFF_AG);
    while ((int)(pcl_comm_ld_acquire_sys(ctr) - target) < 0) { }
  }
  __syncthreads();
  const float* in = reinterpret_cast<const float*>(v.peer[v.rank] + PCL_COMM_OFF_AGDATA) + (size_t)parity * v.nranks * PCL_COMM_AG_MAX;
  for (int k = 0; k < v.nranks; ++k)
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[(size_t)k * n + i] = __ldcg(in + (size_t)k * PCL_COMM_AG_MAX + i);
}

static PclCommView pcl_comm_view(const pcl_comm* c) {
  PclCommView v;
  memset(&v, 0, sizeof(v));
  for (int k = 0; k < c->nranks; ++k) v.peer[k] = c->peer[k];
  v.rank = c->rank; v.nranks = c->nranks;
  return v;
}

extern "C" int pcl_comm_create(int rank, int nranks, size_t window_bytes, pcl_comm** out) {
  if (!out || nranks < 1 || nranks > PCL_COMM_MAXRANKS || rank < 0 || rank >= nranks) { pcl_set_error("bad communicator shape: rank %d of %d (max %d)", rank, nranks, PCL_COMM_MAXRANKS); return PCL_ERR_INVALID; }
  pcl_comm* c = (pcl_comm*)calloc(1, sizeof(pcl_comm));
  if (!c) { pcl_set_error("out of host memory"); return PCL_ERR_INVALID; }
  c->rank = rank; c->nranks = nranks;
  c->rec_off = (PCL_COMM_OFF_AGDATA + (size_t)2 * PCL_COMM_MAXRANKS * PCL_COMM_AG_MAX * sizeof(float) + 4095) & ~(size_t)4095;
  c->bytes = window_bytes ? window_bytes : ((size_t)16 << 20);
  if (c->bytes < c->rec_off + ((size_t)1 << 20)) { free(c); pcl_set_error("communicator window too small"); return PCL_ERR_INVALID; }
  c->rec_bytes = c->bytes - c->rec_off;
  void* p = nullptr;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&p, c->bytes);              // plain cudaMalloc: pool memory cannot be exported over IPC
  if (e == cudaSuccess) e = cudaMemset(p, 0, c->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { if (p) cudaFree(p); free(c); pcl_set_error("communicator window allocation failed: %s", cudaGetErrorString(e)); return PCL_ERR_CUDA; }
  c->peer[rank] = (char*)p;
  c->connected = nranks == 1;
  *out = c;
  return PCL_OK;
}

extern "C" int pcl_comm_handle(const pcl_comm* c, void* handle64) {
  if (!c || !handle64) { pcl_set_error("null communicator or handle buffer"); return PCL_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == PCL_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  PCL_CUDA(cudaIpcGetMemHandle(&h, c->peer[c->rank]));
  memcpy(handle64, &h, sizeof(h));
  return PCL_OK;
}

// handles: nranks x 64 bytes in rank order (this rank's own entry is ignored).  Call after every rank has created its
// window (exchange the handles through any host channel: torch.distributed, MPI, a file).
extern "C" int pcl_comm_connect(pcl_comm* c, const void* handles) {
  if (!c || !handles) { pcl_set_error("null communicator or handles"); return PCL_ERR_INVALID; }
  if (c->connected) return PCL_OK;
  for (int k = 0; k < c->nranks; ++k) {
    if (k == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)k * PCL_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    PCL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[k] = (char*)p;
  }
  c->connected = 1;
  return PCL_OK;
}

// alternative to the IPC exchange: peer windows mapped by the caller (e.g. torch symmetric memory); windows[rank] replaces
// the local window, which must be zero-filled and at least `window_bytes` of pcl_comm_create large
extern "C" int pcl_comm_connect_ptrs(pcl_comm* c, void* const* windows) {
  if (!c || !windows) { pcl_set_error("null communicator or window list"); return PCL_ERR_INVALID; }
  if (c->connected && c->nranks > 1) { pcl_set_error("communicator is already connected"); return PCL_ERR_INVALID; }
  if (windows[c->rank] && windows[c->rank] != c->peer[c->rank]) { cudaFree(c->peer[c->rank]); c->connected = 2; }   // 2: windows are not ours
  for (int k = 0; k < c->nranks; ++k) {
    if (!windows[k]) { pcl_set_error("window %d is null", k); return PCL_ERR_INVALID; }
    c->peer[k] = (char*)windows[k];
  }
  if (c->connected != 2) c->connected = 3;                                                                          // 3: own window, foreign mappings
  return PCL_OK;
}

extern "C" int pcl_comm_rank(const pcl_comm* c) { return c ? c->rank : -1; }
extern "C" int pcl_comm_size(const pcl_comm* c) { return c ? c->nranks : 0; }

extern "C" int pcl_comm_barrier(pcl_comm* c, void* stream) {
  if (!c || !c->connected) { pcl_set_error("communicator is null or not connected"); return PCL_ERR_INVALID; }
  if (c->nranks == 1) return PCL_OK;
  c->bar_epoch += (unsigned int)c->nranks;
  pcl_comm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pcl_comm_view(c), c->bar_epoch);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" int pcl_comm_allgather_f32(pcl_comm* c, const float* src_dev, int n, float* dst_dev, void* stream) {
  if (!c || !c->connected || !src_dev || !dst_dev) { pcl_set_error("communicator not connected or null buffer"); return PCL_ERR_INVALID; }
  if (n < 0 || n > PCL_COMM_AG_MAX) { pcl_set_error("all-gather of %d floats per rank exceeds the slot size %d", n, PCL_COMM_AG_MAX); return PCL_ERR_INVALID; }
  if (n == 0) return PCL_OK;
  if (c->nranks == 1) {
    PCL_CUDA(cudaMemcpyAsync(dst_dev, src_dev, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PCL_OK;
  }
  const int parity = (int)((c->ag_epoch / (unsigned int)c->nranks) & 1u);
  c->ag_epoch += (unsigned int)c->nranks;
  pcl_comm_allgather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pcl_comm_view(c), src_dev, n, dst_dev, c->ag_epoch, parity);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" void pcl_comm_destroy(pcl_comm* c) {
  if (!c) return;
  if (c->connected == 1 && c->nranks > 1) {
    for (int k = 0; k < c->nranks; ++k)
      if (k != c->rank && c->peer[k]) cudaIpcCloseMemHandle(c->peer[k]);
  }
  if (c->connected != 2 && c->peer[c->rank]) cudaFree(c->peer[c->rank]);
  free(c);
}
