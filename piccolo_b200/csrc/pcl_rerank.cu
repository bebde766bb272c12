// Colour-histogram re-rank of the scored candidates (SURVEY §8f "next" #1).
// Replaces trim_input_hist_secondary (utils.py:510-588) = K x { make_pano (utils.py:134-205: sort by distance,
// nine index_put_ calls), 8 block histograms (color_utils.py:68-119), histogram intersection (:122-144) }.
//
// make_pano paints every point nine times — its centre pixel and the eight neighbours, one index_put_ call per offset,
// far points first — so the colour of a pixel is decided by (call order, then distance): the LAST call that reaches
// the pixel wins, and inside a call the nearest point.  All points that reach pixel P in call r have their centre in
// the same source pixel P - offset_r (plus the clamped copies on the first/last column), hence:
//   1. splat: ONE 64-bit atomicMax per (candidate, point) of key = [present | ~distance bits:32 | lit:1 | colour bin:9]
//      into the centre pixel of a per-candidate key image -> the nearest point of every source pixel, carrying the only
//      two things the histogram needs from it;
//   2. gather: per output pixel read the 3x3 neighbourhood of keys (nine independent loads) and take the key of the
//      last call that reached the pixel — no atomics, no second gather of colours — then count its bin in shared memory.
// (Round 1 issued the nine read+atomicMax per point: 450 M cell visits for 50 candidates x 1 M points, latency-bound.)
#include "pcl_common.cuh"

#include <string.h>

// key of a splatted point: [present:1 | ~distance bits:32 | colour is lit:1 | 8x8x8 colour bin:9].  The nearest point of a
// pixel has the largest key; what the histogram needs from the winner — its bin and whether it is lit — rides in the low bits,
// so the gather never has to fetch the point's colour again.
#define PCL_RR_LOW_BITS 10
#define PCL_RR_PRESENT (1ull << (32 + PCL_RR_LOW_BITS))

__global__ void pcl_rr_pose_kernel(const float* __restrict__ poses6, int K, PclPose* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) pcl_pose_from_params(poses6 + 6 * k, out[k]);
}

__device__ __forceinline__ int pcl_rr_bin(float r, float g, float b) {
  // histogram(): value.long() // ceil(255/8)  with value = colour*255 in fp32 (color_utils.py:84-97)
  const int br = (int)(long long)(r * 255.0f) / 32, bg = (int)(long long)(g * 255.0f) / 32, bb = (int)(long long)(b * 255.0f) / 32;
  return br + 8 * bg + 64 * bb;
}

// rows [ky_lo, ky_hi) of the key image are kept: the compared row blocks plus one source row above and below.
// One thread per POINT, looping over the candidates (poses in shared memory): the point, its colour bin and its lit flag
// are loaded / formed once instead of once per candidate (round 2: 210 -> 128 instructions per point·candidate).
#define PCL_RR_POSE_CHUNK 64
__global__ void __launch_bounds__(256) pcl_rr_splat_kernel(const PclCloudView C, const PclPose* __restrict__ poses, const int K, const int H, const int W,
                                                           const int ky_lo, const int ky_hi, unsigned long long* __restrict__ keys) {
  __shared__ PclPose s_pose[PCL_RR_POSE_CHUNK];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool live = i < C.n;
  float px = 0.f, py = 0.f, pz = 0.f;
  unsigned long long low = 0ull;
  if (live) {
    px = C.x[i]; py = C.y[i]; pz = C.z[i];
    const float r = C.r[i], g = C.g[i], b = C.b[i];
    const unsigned int lit = !(r * 255.0f == 0.0f && g * 255.0f == 0.0f && b * 255.0f == 0.0f);       // proj_mask (utils.py:554)
    low = (unsigned long long)((lit << 9) | (unsigned int)pcl_rr_bin(r, g, b));
  }
  const size_t cand_stride = (size_t)(ky_hi - ky_lo) * (size_t)W;
  for (int k0 = 0; k0 < K; k0 += PCL_RR_POSE_CHUNK) {
    const int kn = min(PCL_RR_POSE_CHUNK, K - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < kn * 12; j += blockDim.x) reinterpret_cast<float*>(s_pose)[j] = reinterpret_cast<const float*>(poses + k0)[j];
    __syncthreads();
    if (!live) continue;
#pragma unroll 2
    for (int k = 0; k < kn; ++k) {
      const PclPose P = s_pose[k];
      const float dx = px - P.tx, dy = py - P.ty, dz = pz - P.tz;
      const float qx = P.r00 * dx + P.r01 * dy + P.r02 * dz;
      const float qy = P.r10 * dx + P.r11 * dy + P.r12 * dz;
      const float qz = P.r20 * dx + P.r21 * dy + P.r22 * dz;
      // cloud2idx, fp32 op for op (utils.py:44-59) with the kernels' minimax atan2 (1e-7 rad: moves a point across a
      // pixel-truncation boundary with probability ~1e-5), then make_pano's pixel truncation (utils.py:159-165).
      // Rows first: points whose centre row is outside the kept band stop here.
      const float theta = pcl_atan2_pos(sqrtf(qx * qx + qy * qy), qz + 1e-6f);
      const float v = 2.0f * (theta / 3.14159265358979323846f) - 1.0f;
      const int y = (int)(((v + 1.0f) / 2.0f) * (float)(H - 1));
      if (y < ky_lo || y >= ky_hi) continue;
      const float phi = pcl_atan2(qy, qx + 1e-6f) + 3.14159265358979323846f;
      const float u = 2.0f * (1.0f - phi / 6.28318530717958647692f) - 1.0f;
      const int x = (int)(((u + 1.0f) / 2.0f) * (float)(W - 1));
      const float dist = sqrtf(qx * qx + qy * qy + qz * qz);
      const unsigned long long key = PCL_RR_PRESENT | ((unsigned long long)(~__float_as_uint(dist)) << PCL_RR_LOW_BITS) | low;
      unsigned long long* cell = keys + (size_t)(k0 + k) * cand_stride + (size_t)(y - ky_lo) * (size_t)W + (size_t)x;
      if (__ldcg(cell) < key) atomicMax(cell, key);      // a stale read only costs a redundant atomic, never a wrong result
    }
  }
}

// The point make_pano leaves in pixel (y, x): offsets in REVERSE call order (utils.py:190-198: idx8, 7, ..., 1, centre;
// the last call wins), the first source pixel that holds a key decides.  Source of call (dy, dx) for pixel (y, x) is
// (y - dy, x - dx); on the first / last column (row) the clamped writes of the pixel itself land there too.
__device__ __forceinline__ unsigned long long pcl_rr_winner(const unsigned long long* __restrict__ kimg, const int H, const int W,
                                                            const int ky_lo, const int y, const int x) {
  // (dy, dx) of the calls, last call first: centre, idx1 (+1,+1), idx2 (+1,0), idx3 (+1,-1), idx4 (-1,+1), idx5 (-1,0),
  // idx6 (-1,-1), idx7 (0,+1), idx8 (0,-1)
  const int dys[9] = {0, 1, 1, 1, -1, -1, -1, 0, 0};
  const int dxs[9] = {0, 1, 0, -1, 1, 0, -1, 1, -1};
  if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
    // interior pixel: exactly one source per call.  All nine loads are issued before the first is examined (neighbouring
    // pixels share them through L1): nine independent requests instead of a chain of up to nine dependent ones.
    unsigned long long k[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) k[r] = kimg[(size_t)(y - dys[r] - ky_lo) * W + (x - dxs[r])];
    unsigned long long best = 0ull;
#pragma unroll
    for (int r = 8; r >= 0; --r) best = k[r] ? k[r] : best;      // the earliest r (= the last call) that holds a key wins
    return best;
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int dy = dys[r], dx = dxs[r];
    unsigned long long best = 0ull;
    // candidate source rows / columns: the regular one, and the pixel's own when the write was clamped onto it
    const int sy0 = y - dy, sx0 = x - dx;
    const bool y_reg = sy0 >= 0 && sy0 < H, x_reg = sx0 >= 0 && sx0 < W;
    const bool y_clamp = (dy == 1 && y == H - 1) || (dy == -1 && y == 0);
    const bool x_clamp = (dx == 1 && x == W - 1) || (dx == -1 && x == 0);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (a == 0 ? !y_reg : !y_clamp) continue;
      const int sy = a == 0 ? sy0 : y;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (b == 0 ? !x_reg : !x_clamp) continue;
        const int sx = b == 0 ? sx0 : x;
        const unsigned long long k = kimg[(size_t)(sy - ky_lo) * W + sx];
        best = k > best ? k : best;
      }
    }
    if (best) return best;
  }
  return 0ull;
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

// candidate side, pass 1: raw 512-bin histogram counts of the rendered pixels per (candidate, compared block).  A CTA
// takes a strip of PCL_RR_STRIP rows of one block (K x nblk x bh/STRIP CTAs: no tail wave, no long per-thread loops),
// counts in shared memory and adds its non-empty bins to the block's histogram in global memory.
#define PCL_RR_STRIP 32
__global__ void __launch_bounds__(256) pcl_rr_cand_hist_kernel(const float* __restrict__ img, const unsigned long long* __restrict__ keys,
                                                               const int H, const int W, const int nsh, const int nsw, const int ky_lo, const int ky_hi,
                                                               unsigned int* __restrict__ cand_hist /*[K][nblk][512]*/) {
  __shared__ unsigned int hist[512];
  const int nblk = (nsh - 2) * nsw, bh = H / nsh, bw = W / nsw;
  const int strips = (bh + PCL_RR_STRIP - 1) / PCL_RR_STRIP;
  const int blk = blockIdx.x / strips, strip = blockIdx.x - blk * strips, cand = blockIdx.y;
  const int h = 1 + blk / nsw, w = blk % nsw;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const unsigned long long* kimg = keys + (size_t)cand * (size_t)(ky_hi - ky_lo) * (size_t)W;
  const int y0 = h * bh + strip * PCL_RR_STRIP, y1 = min(y0 + PCL_RR_STRIP, (h + 1) * bh);
  for (int y = y0; y < y1; ++y) {
    const bool row_interior = y > 0 && y < H - 1;
    // the three key rows around y (64-bit words as uint2: .y holds the presence bit, .x the lit flag and the colour bin)
    const uint2* kr = reinterpret_cast<const uint2*>(kimg + (size_t)(y - ky_lo) * W);
    const float* irow = img + (size_t)y * W * 3;
    for (int xo = threadIdx.x; xo < bw; xo += blockDim.x) {
      const int x = w * bw + xo;
      const float i0 = irow[3 * x], i1 = irow[3 * x + 1], i2 = irow[3 * x + 2];   // independent of the keys: in flight together with them
      unsigned int low;
      if (row_interior && x > 0 && x < W - 1) {
        // interior pixel: exactly one source per call.  Nine independent loads at constant offsets from one address;
        // source of call (dy, dx) is (y - dy, x - dx); priority = reverse call order: centre, idx1 (+1,+1), idx2 (+1,0),
        // idx3 (+1,-1), idx4 (-1,+1), idx5 (-1,0), idx6 (-1,-1), idx7 (0,+1), idx8 (0,-1).  The first source that holds a
        // key decides; only its low word (lit flag + bin) is needed.
        const uint2* c = kr + x;
        const uint2 k0 = c[0], k1 = c[-W - 1], k2 = c[-W], k3 = c[-W + 1], k4 = c[W - 1], k5 = c[W], k6 = c[W + 1], k7 = c[-1], k8 = c[1];
        low = k8.y ? k8.x : 0u;
        low = k7.y ? k7.x : low;
        low = k6.y ? k6.x : low;
        low = k5.y ? k5.x : low;
        low = k4.y ? k4.x : low;
        low = k3.y ? k3.x : low;
        low = k2.y ? k2.x : low;
        low = k1.y ? k1.x : low;
        low = k0.y ? k0.x : low;
      } else {
        low = (unsigned int)pcl_rr_winner(kimg, H, W, ky_lo, y, x);        // 0 when no source holds a key
      }
      const bool img_lit = !(i0 * 255.0f == 0.0f && i1 * 255.0f == 0.0f && i2 * 255.0f == 0.0f);       // img_mask
      if (img_lit && ((low >> 9) & 1u)) atomicAdd(&hist[low & 511u], 1u);
    }
  }
  __syncthreads();
  unsigned int* out = cand_hist + ((size_t)cand * nblk + blk) * 512;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) if (hist[i]) atomicAdd(out + i, hist[i]);
}

// pass 2: one warp per (candidate, block): lit-pixel count = sum of the histogram, intersection with the query's
// normalised histogram (hist / hist.sum() on both sides, color_utils.py:103)
__global__ void pcl_rr_intersect_kernel(const unsigned int* __restrict__ cand_hist, const unsigned int* __restrict__ img_hist,
                                        const unsigned int* __restrict__ n_gt, const int nblk, const int K,
                                        float* __restrict__ rows /*[K][2*nblk]: intersection per block, then lit-pixel count per block*/) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= K * nblk) return;
  const int cand = wid / nblk, blk = wid - cand * nblk;
  const unsigned int* hc = cand_hist + (size_t)wid * 512;
  unsigned int cnt[16], total = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) { cnt[j] = hc[lane + 32 * j]; total += cnt[j]; }
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  const float tot = (float)total, gtot = (float)n_gt[blk];
  // summation order of the round-1 kernel (512 threads: thread i owned bin i, warp sums, then the 16 warp sums in order)
  float ws[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = fminf((float)img_hist[blk * 512 + lane + 32 * j] / gtot, (float)cnt[j] / tot);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    ws[j] = s;
  }
  if (lane == 0) {
    float t = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; ++j) t += ws[j];
    rows[(size_t)cand * 2 * nblk + blk] = t;
    rows[(size_t)cand * 2 * nblk + nblk + blk] = tot;            // exact: a block has < 2^24 pixels
  }
}

// The reference's loop, quirks included (utils.py:547-579): the split table persists across candidates, an empty
// block writes 0 and BREAKS the inner (column) loop (cells to its right keep the previous candidate's value),
// NaN -> 0, mean over ALL nsh*nsw cells.  One thread per table cell walks the candidates in order; the per-
// candidate sum over cells is a warp reduction.  Launch with 32 threads.
__global__ void pcl_rr_final_kernel(const float* __restrict__ rows, const float* __restrict__ n_gt,
                                    const int K, const int nsh, const int nsw, float* __restrict__ out) {
  // ONE warp, two table cells per lane (nsh*nsw <= 64): the walk over the candidates is sequential by construction, so it
  // runs without block barriers; the rows sit in shared memory because every read is on the critical path
  extern __shared__ float s_rows[];                              // [K][2*nblk] then n_gt[nblk]
  const int lane = threadIdx.x, ncell = nsh * nsw, nblk = (nsh - 2) * nsw;
  for (int i = lane; i < K * 2 * nblk; i += 32) s_rows[i] = rows[i];
  float* s_ngt = s_rows + (size_t)K * 2 * nblk;
  for (int i = lane; i < nblk; i += 32) s_ngt[i] = n_gt[i];
  __syncwarp();
  float cur[2] = {0.0f, 0.0f};
  for (int c = 0; c < K; ++c) {
    const float* row = s_rows + (size_t)c * 2 * nblk;
    float s = 0.0f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int cell = lane + 32 * half;
      const int h = cell / nsw, w = cell - h * nsw;
      if (cell < ncell && h >= 1 && h < nsh - 1) {
        // first empty column of this row for this candidate (the `break`)
        int first_empty = nsw;
        for (int ww = 0; ww <= w; ++ww) {
          const int blk = (h - 1) * nsw + ww;
          if (row[nblk + blk] == 0.0f || s_ngt[blk] == 0.0f) { first_empty = ww; break; }
        }
        if (w < first_empty) cur[half] = row[(h - 1) * nsw + w];
        else if (w == first_empty) cur[half] = 0.0f;
        if (isnan(cur[half])) cur[half] = 0.0f;
      }
    }
    // the sum over the cells in the order of the 64-thread version (two warp sums of 32 cells, then their sum)
    float s0 = (lane < ncell) ? cur[0] : 0.0f, s1 = (lane + 32 < ncell) ? cur[1] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    s = s0 + s1;
    if (lane == 0) out[c] = s / (float)ncell;
  }
}

__global__ void pcl_rr_ngt_kernel(const unsigned int* __restrict__ n_gt, int nblk, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nblk) out[i] = (float)n_gt[i];
}

static int pcl_rr_check_split(int h, int w, int nsh, int nsw) {
  if (h < 4 || w < 4 || nsh < 1 || nsw < 1 || nsh * nsw > 64 || h / nsh < 1 || w / nsw < 1) {
    pcl_set_error("bad re-rank geometry: %d x %d panorama, %d x %d blocks (at most 64 blocks)", h, w, nsh, nsw);
    return PCL_ERR_INVALID;
  }
  return PCL_OK;
}

// Stage 1, independent per candidate (shards over ranks): rows_k_dev[k][2*nblk] = per compared block the histogram
// intersection, then the number of lit rendered pixels; ngt_dev[nblk] = lit pixels of the query per block.
extern "C" int pcl_hist_rerank_blocks(const pcl_cloud* c, const float* img_hw3_dev, int h, int w, const float* poses_k6_dev, int k,
                                      int num_split_h, int num_split_w, float* rows_k_dev, float* ngt_dev, void* stream) {
  if (!c || !img_hw3_dev || !poses_k6_dev || !rows_k_dev || !ngt_dev || k <= 0) { pcl_set_error("bad re-rank arguments"); return PCL_ERR_INVALID; }
  int rc = pcl_rr_check_split(h, w, num_split_h, num_split_w);
  if (rc) return rc;
  if (num_split_h < 3) return PCL_OK;                     // no compared row blocks (utils.py:548: range(1, nsh-1) is empty): nothing to compute
  cudaStream_t st = (cudaStream_t)stream;
  PclUseGuard guard{c, nullptr, st};
  const int bh = h / num_split_h, nblk = (num_split_h - 2) * num_split_w;
  const int y_lo = bh, y_hi = (num_split_h - 1) * bh;
  const int ky_lo = y_lo - 1, ky_hi = y_hi + 1 < h ? y_hi + 1 : h;       // one source row above and below the compared band
  const size_t key_bytes = (size_t)k * (size_t)(ky_hi - ky_lo) * (size_t)w * sizeof(unsigned long long);
  const size_t off_pose = (key_bytes + 255) & ~(size_t)255;
  const size_t off_ih = off_pose + (((size_t)k * sizeof(PclPose) + 255) & ~(size_t)255);
  const size_t off_ngt = off_ih + (size_t)nblk * 512 * sizeof(unsigned int);
  const size_t off_ch = off_ngt + (((size_t)nblk * sizeof(int) + 255) & ~(size_t)255);
  const size_t total = off_ch + (size_t)k * nblk * 512 * sizeof(unsigned int);
  char* buf;
  PCL_CUDA(pcl_pool_alloc((void**)&buf, total, st));
  cudaError_t e = cudaMemsetAsync(buf, 0, key_bytes, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(buf + off_ih, 0, total - off_ih, st);
  if (e != cudaSuccess) { pcl_pool_free(buf, st); pcl_set_error("re-rank scratch clear failed: %s", cudaGetErrorString(e)); return PCL_ERR_CUDA; }
  unsigned long long* keys = (unsigned long long*)buf;
  PclPose* poses = (PclPose*)(buf + off_pose);
  unsigned int* img_hist = (unsigned int*)(buf + off_ih);
  unsigned int* n_gt = (unsigned int*)(buf + off_ngt);
  unsigned int* cand_hist = (unsigned int*)(buf + off_ch);
  PclCloudView C = {c->x, c->y, c->z, c->r, c->g, c->b, (long long)c->n};
  pcl_rr_pose_kernel<<<(k + 63) / 64, 64, 0, st>>>(poses_k6_dev, k, poses);
  pcl_rr_splat_kernel<<<(unsigned int)((c->n + 255) / 256), 256, 0, st>>>(C, poses, k, h, w, ky_lo, ky_hi, keys);
  pcl_rr_img_hist_kernel<<<dim3(nblk, 32), 256, 0, st>>>(img_hw3_dev, h, w, num_split_h, num_split_w, img_hist, n_gt);
  const int strips = (bh + PCL_RR_STRIP - 1) / PCL_RR_STRIP;
  pcl_rr_cand_hist_kernel<<<dim3(nblk * strips, k), 256, 0, st>>>(img_hw3_dev, keys, h, w, num_split_h, num_split_w, ky_lo, ky_hi, cand_hist);
  pcl_rr_intersect_kernel<<<(k * nblk * 32 + 255) / 256, 256, 0, st>>>(cand_hist, img_hist, n_gt, nblk, k, rows_k_dev);
  pcl_rr_ngt_kernel<<<1, 64, 0, st>>>(n_gt, nblk, ngt_dev);
  g_pcl_launches.fetch_add(6);
  e = cudaGetLastError();
  pcl_pool_free(buf, st);
  if (e != cudaSuccess) { pcl_set_error("re-rank launch failed: %s", cudaGetErrorString(e)); return PCL_ERR_CUDA; }
  return PCL_OK;
}

// Stage 2, over ALL K candidates in their original order (the reference's table persists from one candidate to the next).
extern "C" int pcl_hist_rerank_finish(const float* rows_k_dev, const float* ngt_dev, int k, int num_split_h, int num_split_w,
                                      float* hist_intersect_k_dev, void* stream) {
  if (!rows_k_dev || !ngt_dev || !hist_intersect_k_dev || k <= 0 || num_split_h < 1 || num_split_w < 1 || num_split_h * num_split_w > 64) {
    pcl_set_error("bad re-rank arguments");
    return PCL_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (num_split_h < 3) {                                  // reference: every hist_intersect stays 0
    PCL_CUDA(cudaMemsetAsync(hist_intersect_k_dev, 0, sizeof(float) * (size_t)k, st));
    return PCL_OK;
  }
  const int nblk_f = (num_split_h - 2) * num_split_w;
  const size_t smem_f = ((size_t)k * 2 * nblk_f + nblk_f) * sizeof(float);
  if (smem_f > 200 * 1024) { pcl_set_error("re-rank of %d candidates x %d blocks exceeds the finishing kernel's shared memory", k, nblk_f); return PCL_ERR_INVALID; }
  if (smem_f > 48 * 1024) PCL_CUDA(cudaFuncSetAttribute(pcl_rr_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
  pcl_rr_final_kernel<<<1, 32, smem_f, st>>>(rows_k_dev, ngt_dev, k, num_split_h, num_split_w, hist_intersect_k_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" int pcl_hist_rerank(const pcl_cloud* c, const float* img_hw3_dev, int h, int w, const float* poses_k6_dev, int k,
                               int num_split_h, int num_split_w, float* hist_intersect_k_dev, void* stream) {
  if (!hist_intersect_k_dev || k <= 0) { pcl_set_error("bad re-rank arguments"); return PCL_ERR_INVALID; }
  int rc = pcl_rr_check_split(h, w, num_split_h, num_split_w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = num_split_h >= 3 ? (num_split_h - 2) * num_split_w : 0;
  float* rows = nullptr;
  PCL_CUDA(pcl_pool_alloc((void**)&rows, sizeof(float) * ((size_t)k * 2 * nblk + 64), st));
  float* ngt = rows + (size_t)k * 2 * nblk;
  rc = pcl_hist_rerank_blocks(c, img_hw3_dev, h, w, poses_k6_dev, k, num_split_h, num_split_w, rows, ngt, stream);
  if (rc == PCL_OK) rc = pcl_hist_rerank_finish(rows, ngt, k, num_split_h, num_split_w, hist_intersect_k_dev, stream);
  pcl_pool_free(rows, st);
  return rc;
}
