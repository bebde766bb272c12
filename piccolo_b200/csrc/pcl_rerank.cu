// Colour-histogram re-rank of the scored candidates (SURVEY §8f "next" #1).
// Replaces trim_input_hist_secondary (utils.py:510-588) = K x { make_pano (utils.py:134-205: sort by distance,
// nine index_put_ calls), 8 block histograms (color_utils.py:68-119), histogram intersection (:122-144) }.
//
// One depth-tested splat for ALL candidates at once: every (candidate, point) issues up to nine 64-bit
// atomicMax of  key = [write rank:4 | ~distance bits:32 | point index:28]  into a per-candidate key image, which
// realises the painter's order the reference intends (centre write beats the eight neighbour writes in call
// order, nearest point wins inside a call) without sorting the cloud per candidate.  A second kernel builds the
// 8x8x8 colour histogram of every (candidate, block) in shared memory and intersects it with the query's.
#include "pcl_common.cuh"

#include <string.h>

#define PCL_IDX_BITS 28
#define PCL_IDX_MASK ((1ull << PCL_IDX_BITS) - 1ull)

__global__ void pcl_rr_pose_kernel(const float* __restrict__ poses6, int K, PclPose* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) pcl_pose_from_params(poses6 + 6 * k, out[k]);
}

// rows [y_lo, y_hi) of the key image are kept (the middle row blocks, the only ones that are compared)
__global__ void pcl_rr_splat_kernel(const PclCloudView C, const PclPose* __restrict__ poses, const int H, const int W,
                                    const int y_lo, const int y_hi, unsigned long long* __restrict__ keys) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= C.n) return;
  const PclPose P = poses[blockIdx.y];
  const float dx = C.x[i] - P.tx, dy = C.y[i] - P.ty, dz = C.z[i] - P.tz;
  const float qx = P.r00 * dx + P.r01 * dy + P.r02 * dz;
  const float qy = P.r10 * dx + P.r11 * dy + P.r12 * dz;
  const float qz = P.r20 * dx + P.r21 * dy + P.r22 * dz;
  // cloud2idx, fp32 op for op (utils.py:44-59) with the kernels' minimax atan2 (1e-7 rad: moves a point across a
  // pixel-truncation boundary with probability ~1e-5), then make_pano's pixel truncation (utils.py:159-165).
  // Rows first: only the middle row blocks are ever compared, points landing elsewhere stop here.
  const float theta = pcl_atan2_pos(sqrtf(qx * qx + qy * qy), qz + 1e-6f);
  const float v = 2.0f * (theta / 3.14159265358979323846f) - 1.0f;
  const int y = (int)(((v + 1.0f) / 2.0f) * (float)(H - 1));
  if (y + 1 < y_lo || y - 1 >= y_hi) return;
  const float phi = pcl_atan2(qy, qx + 1e-6f) + 3.14159265358979323846f;
  const float u = 2.0f * (1.0f - phi / 6.28318530717958647692f) - 1.0f;
  const int x = (int)(((u + 1.0f) / 2.0f) * (float)(W - 1));
  const float dist = sqrtf(qx * qx + qy * qy + qz * qz);
  const unsigned long long base = ((unsigned long long)(~__float_as_uint(dist)) << PCL_IDX_BITS) | ((unsigned long long)i & PCL_IDX_MASK);
  unsigned long long* img = keys + (size_t)blockIdx.y * (size_t)(y_hi - y_lo) * (size_t)W;
  const int yp = min(y + 1, H - 1), ym = max(y - 1, 0), xp = min(x + 1, W - 1), xm = max(x - 1, 0);
  // write ranks = call order of utils.py:190-198 (later call overwrites earlier): idx8,7,6,5,4,3,2,1, centre
  const int ys[9] = {y, y, ym, ym, ym, yp, yp, yp, y};
  const int xs[9] = {xm, xp, xm, x, xp, xm, x, xp, x};
#pragma unroll
  for (int r = 8; r >= 0; --r) {              // centre first: it wins most pixels, later (weaker) keys are filtered by the read
    if (ys[r] >= y_lo && ys[r] < y_hi) {
      unsigned long long* cell = img + (size_t)(ys[r] - y_lo) * (size_t)W + (size_t)xs[r];
      const unsigned long long key = ((unsigned long long)(r + 1) << 60) | base;
      if (__ldcg(cell) < key) atomicMax(cell, key);      // a stale read only costs a redundant atomic, never a wrong result
    }
  }
}

__device__ __forceinline__ int pcl_rr_bin(float r, float g, float b) {
  // histogram(): value.long() // ceil(255/8)  with value = colour*255 in fp32 (color_utils.py:84-97)
  const int br = (int)(long long)(r * 255.0f) / 32, bg = (int)(long long)(g * 255.0f) / 32, bb = (int)(long long)(b * 255.0f) / 32;
  return br + 8 * bg + 64 * bb;
}

// query side, once per image: raw 512-bin histogram counts per compared block (blockIdx.y slices the rows)
__global__ void pcl_rr_img_hist_kernel(const float* __restrict__ img, const int H, const int W, const int nsh, const int nsw,
                                       unsigned int* __restrict__ img_hist /*[nblk][512]*/, unsigned int* __restrict__ n_gt /*[nblk]*/) {
  __shared__ unsigned int hist[512];
  __shared__ unsigned int total;
  const int blk = blockIdx.x, bh = H / nsh, bw = W / nsw;
  const int h = 1 + blk / nsw, w = blk % nsw;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  for (int p = blockIdx.y * blockDim.x + threadIdx.x; p < bh * bw; p += gridDim.y * blockDim.x) {
    const int y = h * bh + p / bw, x = w * bw + p % bw;
    const float* px = img + ((size_t)y * W + x) * 3;
    const float r = px[0], g = px[1], b = px[2];
    if (!(r * 255.0f == 0.0f && g * 255.0f == 0.0f && b * 255.0f == 0.0f)) { atomicAdd(&hist[pcl_rr_bin(r, g, b)], 1u); atomicAdd(&total, 1u); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) if (hist[i]) atomicAdd(&img_hist[blk * 512 + i], hist[i]);
  if (threadIdx.x == 0 && total) atomicAdd(&n_gt[blk], total);
}

// candidate side: one CTA per (block, candidate)
__global__ void pcl_rr_cand_hist_kernel(const PclCloudView C, const float* __restrict__ img, const unsigned long long* __restrict__ keys,
                                        const int H, const int W, const int nsh, const int nsw, const int y_lo, const int y_hi,
                                        const unsigned int* __restrict__ img_hist, const unsigned int* __restrict__ n_gt,
                                        float* __restrict__ inter /*[K][nblk]*/, int* __restrict__ n_tgt) {
  __shared__ unsigned int hist[512];
  __shared__ unsigned int total;
  __shared__ float red[32];
  const int blk = blockIdx.x, cand = blockIdx.y, nblk = gridDim.x, bh = H / nsh, bw = W / nsw;
  const int h = 1 + blk / nsw, w = blk % nsw;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const unsigned long long* kimg = keys + (size_t)cand * (size_t)(y_hi - y_lo) * (size_t)W;
  for (int p = threadIdx.x; p < bh * bw; p += blockDim.x) {
    const int y = h * bh + p / bw, x = w * bw + p % bw;
    const unsigned long long key = kimg[(size_t)(y - y_lo) * W + x];
    if (key == 0ull) continue;
    const float* px = img + ((size_t)y * W + x) * 3;
    if (px[0] * 255.0f == 0.0f && px[1] * 255.0f == 0.0f && px[2] * 255.0f == 0.0f) continue;     // img_mask
    const long long i = (long long)(key & PCL_IDX_MASK);
    const float r = C.r[i], g = C.g[i], b = C.b[i];
    if (r * 255.0f == 0.0f && g * 255.0f == 0.0f && b * 255.0f == 0.0f) continue;                  // proj_mask
    atomicAdd(&hist[pcl_rr_bin(r, g, b)], 1u);
    atomicAdd(&total, 1u);
  }
  __syncthreads();
  float s = 0.0f;
  const float tot = (float)total, gtot = (float)n_gt[blk];      // hist / hist.sum() on both sides (color_utils.py:103)
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s += fminf((float)img_hist[blk * 512 + i] / gtot, (float)hist[i] / tot);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    inter[cand * nblk + blk] = t;
    n_tgt[cand * nblk + blk] = (int)total;
  }
}

// The reference's loop, quirks included (utils.py:547-579): the split table persists across candidates, an empty
// block writes 0 and BREAKS the inner (column) loop (cells to its right keep the previous candidate's value),
// NaN -> 0, mean over ALL nsh*nsw cells.  One thread per table cell walks the candidates in order; the per-
// candidate sum over cells is a block reduction.  Launch with 64 threads.
__global__ void pcl_rr_final_kernel(const float* __restrict__ inter, const int* __restrict__ n_tgt, const unsigned int* __restrict__ n_gt,
                                    const int K, const int nsh, const int nsw, float* __restrict__ out) {
  __shared__ float red[2];
  const int cell = threadIdx.x, ncell = nsh * nsw, nblk = (nsh - 2) * nsw;
  const int h = cell / nsw, w = cell - h * nsw;
  const bool compared = cell < ncell && h >= 1 && h < nsh - 1;
  float cur = 0.0f;
  for (int c = 0; c < K; ++c) {
    if (compared) {
      // first empty column of this row for this candidate (the `break`)
      int first_empty = nsw;
      for (int ww = 0; ww <= w; ++ww) {
        const int blk = (h - 1) * nsw + ww;
        if (n_tgt[c * nblk + blk] == 0 || n_gt[blk] == 0u) { first_empty = ww; break; }
      }
      if (w < first_empty) cur = inter[c * nblk + (h - 1) * nsw + w];
      else if (w == first_empty) cur = 0.0f;
      if (isnan(cur)) cur = 0.0f;
    }
    float s = (cell < ncell) ? cur : 0.0f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) out[c] = (red[0] + red[1]) / (float)ncell;
    __syncthreads();
  }
}

extern "C" int pcl_hist_rerank(const pcl_cloud* c, const float* img_hw3_dev, int h, int w, const float* poses_k6_dev, int k,
                               int num_split_h, int num_split_w, float* hist_intersect_k_dev, void* stream) {
  if (!c || !img_hw3_dev || !poses_k6_dev || !hist_intersect_k_dev || k <= 0 || h < 4 || w < 4) { pcl_set_error("bad re-rank arguments"); return PCL_ERR_INVALID; }
  if (num_split_h < 3 || num_split_w < 1 || num_split_h * num_split_w > 64) { pcl_set_error("num_split_h must be >= 3 and num_split_h*num_split_w <= 64"); return PCL_ERR_INVALID; }
  if (c->n > (long long)PCL_IDX_MASK) { pcl_set_error("re-rank supports up to 2^28 points"); return PCL_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int bh = h / num_split_h, nblk = (num_split_h - 2) * num_split_w;
  const int y_lo = bh, y_hi = (num_split_h - 1) * bh;
  const size_t key_bytes = (size_t)k * (size_t)(y_hi - y_lo) * (size_t)w * sizeof(unsigned long long);
  const size_t off_pose = (key_bytes + 255) & ~(size_t)255;
  const size_t off_ih = off_pose + (((size_t)k * sizeof(PclPose) + 255) & ~(size_t)255);
  const size_t off_ngt = off_ih + (size_t)nblk * 512 * sizeof(unsigned int);
  const size_t off_inter = off_ngt + (((size_t)nblk * sizeof(int) + 255) & ~(size_t)255);
  const size_t off_ntgt = off_inter + (((size_t)k * nblk * sizeof(float) + 255) & ~(size_t)255);
  const size_t total = off_ntgt + (size_t)k * nblk * sizeof(int);
  char* buf;
  PCL_CUDA(pcl_pool_alloc((void**)&buf, total, st));
  PCL_CUDA(cudaMemsetAsync(buf, 0, key_bytes, st));
  PCL_CUDA(cudaMemsetAsync(buf + off_ih, 0, off_inter - off_ih, st));
  unsigned long long* keys = (unsigned long long*)buf;
  PclPose* poses = (PclPose*)(buf + off_pose);
  unsigned int* img_hist = (unsigned int*)(buf + off_ih);
  unsigned int* n_gt = (unsigned int*)(buf + off_ngt);
  float* inter = (float*)(buf + off_inter);
  int* n_tgt = (int*)(buf + off_ntgt);
  PclCloudView C = {c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
  pcl_rr_pose_kernel<<<(k + 63) / 64, 64, 0, st>>>(poses_k6_dev, k, poses);
  PCL_LAUNCH_CHECK();
  pcl_rr_splat_kernel<<<dim3((unsigned int)((c->n + 255) / 256), k), 256, 0, st>>>(C, poses, h, w, y_lo, y_hi, keys);
  PCL_LAUNCH_CHECK();
  pcl_rr_img_hist_kernel<<<dim3(nblk, 32), 256, 0, st>>>(img_hw3_dev, h, w, num_split_h, num_split_w, img_hist, n_gt);
  PCL_LAUNCH_CHECK();
  pcl_rr_cand_hist_kernel<<<dim3(nblk, k), 512, 0, st>>>(C, img_hw3_dev, keys, h, w, num_split_h, num_split_w, y_lo, y_hi, img_hist, n_gt, inter, n_tgt);
  PCL_LAUNCH_CHECK();
  pcl_rr_final_kernel<<<1, 64, 0, st>>>(inter, n_tgt, n_gt, k, num_split_h, num_split_w, hist_intersect_k_dev);
  PCL_LAUNCH_CHECK();
  pcl_pool_free(buf, st);
  return PCL_OK;
}
