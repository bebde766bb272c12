// Per-query colour matching of the panorama to the cloud (`color_match`, color_utils.py:146-234; the
// `match_color` switch of configs/omniscenes.ini, called at localize.py:402-404) — SURVEY §8f "next" #4.
//
// The reference matches, per channel, the sin(latitude)-weighted cumulative histogram of the lit panorama pixels to
// the cumulative histogram of the cloud colours and rewrites every lit pixel through the resulting level -> value
// map.  Both inputs are uint8/255 data, so everything the map depends on is three 256-bin histograms per side:
//   pcl_color_stats   one pass over the panorama (weighted level histogram, level and value presence) and one over
//                     the cloud (value counts); flags inputs that are not exactly k/255
//   (host)            cumulative sums + the reference's interpolation on <= 256 entries per channel
//                     (piccolo_b200/color_utils.py, shared with the CPU restatement that is pinned to the reference)
//   pcl_color_apply   one pass: lit pixels are rewritten through the 3 x 256 look-up table
#include "pcl_common.cuh"

// value id of a uint8/255 float, or -1 when the float is not exactly (float)k / 255.0f
__device__ __forceinline__ int pcl_color_value_id(float v) {
  const int k = __float2int_rn(v * 255.0f);
  return (k >= 0 && k <= 255 && v == (float)k / 255.0f) ? k : -1;
}

// lit mask of color_utils.py:222-223: (img * 255).long().sum(-1) > 0  (truncated levels)
__device__ __forceinline__ bool pcl_color_lit(float r, float g, float b) {
  return ((long long)(r * 255.0f) + (long long)(g * 255.0f) + (long long)(b * 255.0f)) > 0;
}

// One CTA per row at a time (grid-stride over rows): integer shared-memory histograms of the row, then thread i folds
// bins i, i+256, i+512 into register accumulators with the row's weight — no floating-point atomics in shared memory.
__global__ void __launch_bounds__(256)
pcl_color_image_stats_kernel(const float* __restrict__ img, const int H, const int W, const float* __restrict__ row_weight,
                             double* __restrict__ whist /*[3][256]*/, unsigned int* __restrict__ level_cnt /*[3][256]*/,
                             unsigned int* __restrict__ value_cnt /*[3][256]*/, int* __restrict__ inexact) {
  __shared__ unsigned int s_l[768], s_v[768];
  double acc_w[3] = {0.0, 0.0, 0.0};
  unsigned int acc_l[3] = {0u, 0u, 0u}, acc_v[3] = {0u, 0u, 0u};
  bool bad = false;
  for (int row = blockIdx.x; row < H; row += gridDim.x) {
    for (int i = threadIdx.x; i < 768; i += 256) { s_l[i] = 0u; s_v[i] = 0u; }
    __syncthreads();
    const float* line = img + (size_t)row * W * 3;
    for (int x = threadIdx.x; x < W; x += 256) {
      const float r = line[3 * x], g = line[3 * x + 1], b = line[3 * x + 2];
      const int kr = pcl_color_value_id(r), kg = pcl_color_value_id(g), kb = pcl_color_value_id(b);
      if ((kr | kg | kb) < 0) { bad = true; continue; }
      if (!pcl_color_lit(r, g, b)) continue;
      // bincount((source * 255).int(), weight): TRUNCATED level; unique(source): exact value id
      atomicAdd(&s_l[(int)(r * 255.0f)], 1u); atomicAdd(&s_l[256 + (int)(g * 255.0f)], 1u); atomicAdd(&s_l[512 + (int)(b * 255.0f)], 1u);
      atomicAdd(&s_v[kr], 1u); atomicAdd(&s_v[256 + kg], 1u); atomicAdd(&s_v[512 + kb], 1u);
    }
    __syncthreads();
    const double wgt = (double)__ldg(row_weight + row);     // sin_weight = sin(row / H * pi) (color_utils.py:216-217), from the caller
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const unsigned int l = s_l[threadIdx.x + 256 * j];
      acc_w[j] += wgt * (double)l; acc_l[j] += l; acc_v[j] += s_v[threadIdx.x + 256 * j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (acc_l[j]) { atomicAdd(whist + i, acc_w[j]); atomicAdd(level_cnt + i, acc_l[j]); }
    if (acc_v[j]) atomicAdd(value_cnt + i, acc_v[j]);
  }
  if (bad) atomicOr(inexact, 1);
}

__global__ void pcl_color_cloud_stats_kernel(const float* __restrict__ rgb, const long long n, unsigned long long* __restrict__ counts /*[3][256]*/,
                                             int* __restrict__ inexact) {
  __shared__ unsigned int s_c[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) (&s_c[0][0])[i] = 0u;
  __syncthreads();
  bool bad = false;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int k = pcl_color_value_id(rgb[3 * p + c]);
      if (k < 0) bad = true; else atomicAdd(&s_c[c][k], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 768; i += blockDim.x) if ((&s_c[0][0])[i]) atomicAdd(counts + i, (unsigned long long)(&s_c[0][0])[i]);
  if (bad) atomicOr(inexact, 2);
}

__global__ void pcl_color_apply_kernel(const float* __restrict__ img, const long long npix, const float* __restrict__ lut /*[3][256]*/,
                                       float* __restrict__ out) {
  __shared__ float s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const float r = img[3 * p], g = img[3 * p + 1], b = img[3 * p + 2];
    const bool lit = pcl_color_lit(r, g, b);
    out[3 * p] = lit ? s_lut[__float2int_rn(r * 255.0f)] : r;
    out[3 * p + 1] = lit ? s_lut[256 + __float2int_rn(g * 255.0f)] : g;
    out[3 * p + 2] = lit ? s_lut[512 + __float2int_rn(b * 255.0f)] : b;
  }
}

static int pcl_color_blocks(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

// stats_dev layout (caller-allocated, 3 x 256 each): double whist | uint32 level_cnt | uint32 value_cnt | uint64 cloud_cnt | int32 flags[2]
extern "C" int pcl_color_stats(const float* img_hw3_dev, int h, int w, const float* row_weight_h_dev, const float* rgb_n3_dev, int64_t n,
                               void* stats_dev, void* stream) {
  if (!img_hw3_dev || !row_weight_h_dev || !rgb_n3_dev || !stats_dev || h < 1 || w < 1 || n < 1) { pcl_set_error("bad colour-statistics arguments"); return PCL_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)stats_dev;
  PCL_CUDA(cudaMemsetAsync(base, 0, PCL_COLOR_STATS_BYTES, st));
  double* whist = (double*)base;
  unsigned int* level_cnt = (unsigned int*)(base + 768 * 8);
  unsigned int* value_cnt = level_cnt + 768;
  unsigned long long* cloud_cnt = (unsigned long long*)(base + 768 * 16);
  int* flags = (int*)(base + 768 * 24);
  pcl_color_image_stats_kernel<<<(h < 148 * 4 ? h : 148 * 4), 256, 0, st>>>(img_hw3_dev, h, w, row_weight_h_dev, whist, level_cnt, value_cnt, flags);
  PCL_LAUNCH_CHECK();
  pcl_color_cloud_stats_kernel<<<pcl_color_blocks(n), 256, 0, st>>>(rgb_n3_dev, n, cloud_cnt, flags);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" int pcl_color_apply(const float* img_hw3_dev, int h, int w, const float* lut_3x256_dev, float* out_hw3_dev, void* stream) {
  if (!img_hw3_dev || !lut_3x256_dev || !out_hw3_dev || h < 1 || w < 1) { pcl_set_error("bad colour-apply arguments"); return PCL_ERR_INVALID; }
  pcl_color_apply_kernel<<<pcl_color_blocks((long long)h * w), 256, 0, (cudaStream_t)stream>>>(img_hw3_dev, (long long)h * w, lut_3x256_dev, out_hw3_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

// ------------------------------------------------------------------------------------------------
// color_mod (`sharpen_color`, color_utils.py:7-65; localize.py:173-179, :405-410): joint histogram equalisation of
// the luma of panorama and cloud in 8-bit YCrCb.  The reference goes through cv2.cvtColor on uint8 data; its
// fixed-point conversion (yuv_shift 14, coefficients 4899/9617/1868, 11682/9241, 22987/-11698/-5636/29049, delta
// 128) is restated here in integers — checked against cv2 for all 2^24 triples in both directions
// (tests/test_host_logic.py).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pcl_sat8(int v) { return min(max(v, 0), 255); }

struct PclYcc { int y, cr, cb; };

// (colour * 255.).astype(uint8) -> cv2.COLOR_RGB2YCR_CB
__device__ __forceinline__ PclYcc pcl_rgb_to_ycc(float rf, float gf, float bf) {
  const int r = (int)(unsigned char)(rf * 255.0f), g = (int)(unsigned char)(gf * 255.0f), b = (int)(unsigned char)(bf * 255.0f);
  PclYcc o;
  o.y = pcl_sat8((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
  o.cr = pcl_sat8(((r - o.y) * 11682 + (128 << 14) + (1 << 13)) >> 14);
  o.cb = pcl_sat8(((b - o.y) * 9241 + (128 << 14) + (1 << 13)) >> 14);
  return o;
}

// luma bin: ((u8 / 255.) * (num_bins - 1)).long()   (color_utils.py:38-39)
__device__ __forceinline__ int pcl_luma_bin(int y, float scale) { return (int)(((float)y / 255.0f) * scale); }

__global__ void pcl_color_mod_stats_kernel(const float* __restrict__ img, const long long npix, const float* __restrict__ rgb, const long long n,
                                           const int num_bins, unsigned long long* __restrict__ hist /*[2][num_bins]: image, cloud*/) {
  extern __shared__ unsigned int s_h[];       // [2][num_bins]
  for (int i = threadIdx.x; i < 2 * num_bins; i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  const float scale = (float)(num_bins - 1);
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long p = t0; p < npix; p += stride) {
    const float r = img[3 * p], g = img[3 * p + 1], b = img[3 * p + 2];
    if (!pcl_color_lit(r, g, b)) continue;
    atomicAdd(&s_h[pcl_luma_bin(pcl_rgb_to_ycc(r, g, b).y, scale)], 1u);
  }
  for (long long p = t0; p < n; p += stride)
    atomicAdd(&s_h[num_bins + pcl_luma_bin(pcl_rgb_to_ycc(rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]).y, scale)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * num_bins; i += blockDim.x) if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

// equalised luma -> (ycc * 255.).astype(uint8) -> cv2.COLOR_YCR_CB2RGB -> / 255.
__device__ __forceinline__ void pcl_color_mod_one(float r, float g, float b, const float* __restrict__ cdf, float scale, float* out) {
  const PclYcc c = pcl_rgb_to_ycc(r, g, b);
  const int y = (int)(unsigned char)(__ldg(cdf + pcl_luma_bin(c.y, scale)) * 255.0f);
  const int cr = (int)(unsigned char)(((float)c.cr / 255.0f) * 255.0f) - 128, cb = (int)(unsigned char)(((float)c.cb / 255.0f) * 255.0f) - 128;
  out[0] = (float)pcl_sat8(y + ((cr * 22987 + (1 << 13)) >> 14)) / 255.0f;
  out[1] = (float)pcl_sat8(y + ((cb * -5636 + cr * -11698 + (1 << 13)) >> 14)) / 255.0f;
  out[2] = (float)pcl_sat8(y + ((cb * 29049 + (1 << 13)) >> 14)) / 255.0f;
}

__global__ void pcl_color_mod_apply_kernel(const float* __restrict__ img, const long long npix, const float* __restrict__ rgb, const long long n,
                                           const int num_bins, const float* __restrict__ cdf, float* __restrict__ out_img, float* __restrict__ out_rgb) {
  const float scale = (float)(num_bins - 1);
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long p = t0; p < npix; p += stride) {
    const float r = img[3 * p], g = img[3 * p + 1], b = img[3 * p + 2];
    if (pcl_color_lit(r, g, b)) pcl_color_mod_one(r, g, b, cdf, scale, out_img + 3 * p);
    else { out_img[3 * p] = r; out_img[3 * p + 1] = g; out_img[3 * p + 2] = b; }
  }
  for (long long p = t0; p < n; p += stride) pcl_color_mod_one(rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2], cdf, scale, out_rgb + 3 * p);
}

extern "C" int pcl_color_mod_stats(const float* img_hw3_dev, int h, int w, const float* rgb_n3_dev, int64_t n, int num_bins,
                                   unsigned long long* hist_2xbins_dev, void* stream) {
  if (!img_hw3_dev || !rgb_n3_dev || !hist_2xbins_dev || h < 1 || w < 1 || n < 1 || num_bins < 2 || num_bins > 4096) {
    pcl_set_error("bad colour-equalisation arguments (num_bins must be 2..4096)");
    return PCL_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PCL_CUDA(cudaMemsetAsync(hist_2xbins_dev, 0, (size_t)2 * num_bins * sizeof(unsigned long long), st));
  const long long m = (long long)h * w > n ? (long long)h * w : n;
  pcl_color_mod_stats_kernel<<<pcl_color_blocks(m), 256, (size_t)2 * num_bins * sizeof(unsigned int), st>>>(img_hw3_dev, (long long)h * w, rgb_n3_dev, n,
                                                                                                           num_bins, hist_2xbins_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}

extern "C" int pcl_color_mod_apply(const float* img_hw3_dev, int h, int w, const float* rgb_n3_dev, int64_t n, int num_bins,
                                   const float* cdf_bins_dev, float* out_img_hw3_dev, float* out_rgb_n3_dev, void* stream) {
  if (!img_hw3_dev || !rgb_n3_dev || !cdf_bins_dev || !out_img_hw3_dev || !out_rgb_n3_dev || h < 1 || w < 1 || n < 1 || num_bins < 2) {
    pcl_set_error("bad colour-equalisation arguments");
    return PCL_ERR_INVALID;
  }
  const long long m = (long long)h * w > n ? (long long)h * w : n;
  pcl_color_mod_apply_kernel<<<pcl_color_blocks(m), 256, 0, (cudaStream_t)stream>>>(img_hw3_dev, (long long)h * w, rgb_n3_dev, n, num_bins, cdf_bins_dev,
                                                                                    out_img_hw3_dev, out_rgb_n3_dev);
  PCL_LAUNCH_CHECK();
  return PCL_OK;
}
