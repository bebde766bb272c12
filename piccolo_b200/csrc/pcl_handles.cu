// Handles: coloured point cloud (SoA, optional Morton order, clamp box), panorama texel tables, top-k.
// One-off set-up work per room / per query (localize.py:159-164, :169-170; utils.py:208-229, :501-502);
// the radix sorts use CUB (library code, not on the per-iteration path).
#include "pcl_common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// cloud
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int pcl_float_order(float f) {        // monotone float -> uint
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pcl_float_unorder(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void pcl_bbox_kernel(const float* __restrict__ xyz, long long n, unsigned int* mm /*[6]: min3,max3*/) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { const float v = xyz[3 * i + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[k], pcl_float_order(lo[k])); atomicMax(&mm[3 + k], pcl_float_order(hi[k])); }
  }
}

__device__ __forceinline__ unsigned int pcl_spread10(unsigned int v) {   // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// 30-bit Morton code (10 bits per axis: 1024^3 cells — finer than the point spacing needed for lane locality);
// 32-bit keys halve the radix-sort passes of the one-off packing step
__global__ void pcl_morton_kernel(const float* __restrict__ xyz, long long n, const unsigned int* mm,
                                  unsigned int* keys, unsigned int* idx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int code = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = pcl_float_unorder(mm[k]), hi = pcl_float_unorder(mm[3 + k]);
    const float ext = fmaxf(hi - lo, 1e-20f);
    float t = (xyz[3 * i + k] - lo) / ext;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    code |= pcl_spread10((unsigned int)(t * 1023.0f)) << k;
  }
  keys[i] = code;
  idx[i] = (unsigned int)i;
}

__global__ void pcl_gather_kernel(const float* __restrict__ xyz, const float* __restrict__ rgb, const unsigned int* __restrict__ perm,
                                  long long n, float* x, float* y, float* z, float* r, float* g, float* b) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long s = perm ? (long long)perm[i] : i;
  x[i] = xyz[3 * s]; y[i] = xyz[3 * s + 1]; z[i] = xyz[3 * s + 2];
  r[i] = rgb[3 * s]; g[i] = rgb[3 * s + 1]; b[i] = rgb[3 * s + 2];
}

static int pcl_cloud_build(pcl_cloud* c, const float* xyz, const float* rgb, int64_t n, double q, int order, cudaStream_t st);

extern "C" int pcl_cloud_create(const float* xyz, const float* rgb, int64_t n, double q, int order, void* stream, pcl_cloud** out) {
  if (!xyz || !rgb || !out || n <= 0 || n > 0x7fffff00ll) { pcl_set_error("bad cloud arguments (n=%lld)", (long long)n); return PCL_ERR_INVALID; }
  if (!(q >= 0.0 && q <= 1.0)) { pcl_set_error("quantile %g outside [0,1]", q); return PCL_ERR_INVALID; }
  pcl_cloud* c = (pcl_cloud*)calloc(1, sizeof(pcl_cloud));
  if (!c) { pcl_set_error("out of host memory"); return PCL_ERR_INVALID; }
  const int rc = pcl_cloud_build(c, xyz, rgb, n, q, order, (cudaStream_t)stream);
  if (rc != PCL_OK) { pcl_cloud_destroy(c); return rc; }      // nothing leaks on a failed build
  *out = c;
  return PCL_OK;
}

static int pcl_cloud_build(pcl_cloud* c, const float* xyz, const float* rgb, int64_t n, double q, int order, cudaStream_t st) {
  c->n = n;
  c->n_pad = (n + PCL_TILE_ALIGN - 1) / PCL_TILE_ALIGN * PCL_TILE_ALIGN;
  c->order = order;
  c->owner = st;
  PCL_CUDA(pcl_pool_alloc((void**)&c->block, sizeof(float) * (6 * (size_t)c->n_pad + 8), st));
  PCL_CUDA(cudaMemsetAsync(c->block, 0, sizeof(float) * (6 * (size_t)c->n_pad + 8), st));
  c->lo_hi_dev = c->block + 6 * (size_t)c->n_pad;
  c->x = c->block; c->y = c->x + c->n_pad; c->z = c->y + c->n_pad;
  c->r = c->z + c->n_pad; c->g = c->r + c->n_pad; c->b = c->g + c->n_pad;

  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  unsigned int* perm = nullptr;
  struct Scratch {                       // freed on every exit path, in stream order
    void* p[3] = {nullptr, nullptr, nullptr};
    cudaStream_t st;
    ~Scratch() { for (void* q : p) pcl_pool_free(q, st); }
  } sc;
  sc.st = st;
  void*& scratch = sc.p[0];
  if (order == PCL_CLOUD_MORTON) {
    unsigned int* mm; unsigned int *k_in, *k_out; unsigned int *v_in, *v_out;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned int*)nullptr, (unsigned int*)nullptr,
                                    (unsigned int*)nullptr, (unsigned int*)nullptr, (int)n, 0, 30, st);
    const size_t nb = (size_t)n;
    const size_t total = 64 + nb * 4 * 2 + nb * 4 * 2 + tmp_bytes + 256;
    PCL_CUDA(pcl_pool_alloc(&scratch, total, st));
    char* pch = (char*)scratch;
    mm = (unsigned int*)pch; pch += 64;
    k_in = (unsigned int*)pch; pch += nb * 4;
    k_out = (unsigned int*)pch; pch += nb * 4;
    v_in = (unsigned int*)pch; pch += nb * 4;
    v_out = (unsigned int*)pch; pch += nb * 4;
    pch = (char*)(((uintptr_t)pch + 255) & ~(uintptr_t)255);
    const unsigned int init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    PCL_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    pcl_bbox_kernel<<<blocks < 1184 ? blocks : 1184, threads, 0, st>>>(xyz, n, mm);
    PCL_LAUNCH_CHECK();
    pcl_morton_kernel<<<blocks, threads, 0, st>>>(xyz, n, mm, k_in, v_in);
    PCL_LAUNCH_CHECK();
    PCL_CUDA(cub::DeviceRadixSort::SortPairs(pch, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, 30, st));
    perm = v_out;
  }
  pcl_gather_kernel<<<blocks, threads, 0, st>>>(xyz, rgb, perm, n, c->x, c->y, c->z, c->r, c->g, c->b);
  PCL_LAUNCH_CHECK();

  // clamp box: order statistics int(N q), int(N (1-q)) per axis (utils.py:222-227)
  {
    long long i_lo = (long long)((double)n * q), i_hi = (long long)((double)n * (1.0 - q));
    if (i_lo > n - 1) i_lo = n - 1;
    if (i_hi > n - 1) i_hi = n - 1;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (float*)nullptr, (float*)nullptr, (int)n, 0, 32, st);
    void*& tmp = sc.p[1];
    PCL_CUDA(pcl_pool_alloc(&tmp, tmp_bytes + 256, st));
    PCL_CUDA(pcl_pool_alloc(&sc.p[2], sizeof(float) * (size_t)n, st));
    float* sorted = (float*)sc.p[2];
    const float* axes[3] = {c->x, c->y, c->z};
    for (int k = 0; k < 3; ++k) {
      PCL_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, axes[k], sorted, (int)n, 0, 32, st));
      PCL_CUDA(cudaMemcpyAsync(c->lo_hi_dev + k, sorted + i_lo, sizeof(float), cudaMemcpyDeviceToDevice, st));
      PCL_CUDA(cudaMemcpyAsync(c->lo_hi_dev + 3 + k, sorted + i_hi, sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  }
  return PCL_OK;
}

extern "C" int64_t pcl_cloud_size(const pcl_cloud* c) { return c ? c->n : 0; }

extern "C" int pcl_cloud_bounds(const pcl_cloud* c, float* lo_hi) {
  if (!c || !lo_hi) { pcl_set_error("null cloud or output"); return PCL_ERR_INVALID; }
  if (!c->lo_hi_valid) {            // first call: one blocking read of the 6 floats
    pcl_cloud* cc = const_cast<pcl_cloud*>(c);
    PCL_CUDA(cudaMemcpyAsync(cc->lo_hi, c->lo_hi_dev, sizeof(float) * 6, cudaMemcpyDeviceToHost, c->owner));
    PCL_CUDA(cudaStreamSynchronize(c->owner));
    cc->lo_hi_valid = 1;
  }
  memcpy(lo_hi, c->lo_hi, sizeof(float) * 6);
  return PCL_OK;
}

extern "C" void pcl_cloud_destroy(pcl_cloud* c) {
  if (!c) return;
  pcl_use_release(c->owner, &c->use);
  if (c->block) pcl_pool_free(c->block, c->owner);
  free(c);
}

// ------------------------------------------------------------------------------------------------
// image
// ------------------------------------------------------------------------------------------------
__global__ void pcl_u8_exact_kernel(const float* __restrict__ img, long long n, int* not_exact) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = img[i];
  const float b = rintf(v * 255.0f);
  const bool ok = (b >= 0.0f) && (b <= 255.0f) && ((b / 255.0f == v) || ((float)((double)b / 255.0) == v));
  if (!ok) *not_exact = 1;
}

__device__ __forceinline__ unsigned int pcl_pack_texel(const float* __restrict__ img, int H, int W, int y, int x) {
  if (x < 0 || y < 0 || x >= W || y >= H) return 0u;
  const float* p = img + ((size_t)y * W + x) * 3;
  const unsigned int r = (unsigned int)rintf(p[0] * 255.0f), g = (unsigned int)rintf(p[1] * 255.0f), b = (unsigned int)rintf(p[2] * 255.0f);
  return r | (g << 8) | (b << 16);
}

// quad table: entry (y0+1, x0+1), y0 in [-1,H-1], x0 in [-1,W-1] = {nw, ne, sw, se}
__global__ void pcl_build_u8q_kernel(const float* __restrict__ img, int H, int W, uint4* tab) {
  const int xe = blockIdx.x * blockDim.x + threadIdx.x, ye = blockIdx.y;
  if (xe > W) return;
  const int x0 = xe - 1, y0 = ye - 1;
  uint4 e;
  e.x = pcl_pack_texel(img, H, W, y0, x0); e.y = pcl_pack_texel(img, H, W, y0, x0 + 1);
  e.z = pcl_pack_texel(img, H, W, y0 + 1, x0); e.w = pcl_pack_texel(img, H, W, y0 + 1, x0 + 1);
  tab[(size_t)ye * (W + 1) + xe] = e;
}

// fp16 basis table: entry (y0+1, x0+1) = 16 halves {nw, ne-nw, sw-nw, (se-sw)-(ne-nw)} x {R,G,B} + 4 pad
__global__ void pcl_build_f16d_kernel(const float* __restrict__ img, int H, int W, uint4* tab) {
  const int xe = blockIdx.x * blockDim.x + threadIdx.x, ye = blockIdx.y;
  if (xe > W) return;
  const int x0 = xe - 1, y0 = ye - 1;
  const unsigned int t[4] = {pcl_pack_texel(img, H, W, y0, x0), pcl_pack_texel(img, H, W, y0, x0 + 1),
                             pcl_pack_texel(img, H, W, y0 + 1, x0), pcl_pack_texel(img, H, W, y0 + 1, x0 + 1)};
  unsigned int w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < 3; ++c) {
    const int nw = (t[0] >> (8 * c)) & 0xff, ne = (t[1] >> (8 * c)) & 0xff, sw = (t[2] >> (8 * c)) & 0xff, se = (t[3] >> (8 * c)) & 0xff;
    const unsigned int h0 = __half_as_ushort(__int2half_rn(nw)), h1 = __half_as_ushort(__int2half_rn(ne - nw));
    const unsigned int h2 = __half_as_ushort(__int2half_rn(sw - nw)), h3 = __half_as_ushort(__int2half_rn((se - sw) - (ne - nw)));
    w[2 * c] = h0 | (h1 << 16);
    w[2 * c + 1] = h2 | (h3 << 16);
  }
  uint4* e = tab + 2 * ((size_t)ye * (W + 1) + xe);
  e[0] = make_uint4(w[0], w[1], w[2], w[3]);
  e[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// plain tables with a one-texel zero border: entry (y+1, x+1), y in [-1,H], x in [-1,W]
__global__ void pcl_build_u8p_kernel(const float* __restrict__ img, int H, int W, unsigned int* tab) {
  const int xe = blockIdx.x * blockDim.x + threadIdx.x, ye = blockIdx.y;
  if (xe > W + 1) return;
  tab[(size_t)ye * (W + 2) + xe] = pcl_pack_texel(img, H, W, ye - 1, xe - 1);
}

__global__ void pcl_build_f32_kernel(const float* __restrict__ img, int H, int W, float4* tab) {
  const int xe = blockIdx.x * blockDim.x + threadIdx.x, ye = blockIdx.y;
  if (xe > W + 1) return;
  const int x = xe - 1, y = ye - 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= 0 && y >= 0 && x < W && y < H) { const float* p = img + ((size_t)y * W + x) * 3; v = make_float4(p[0], p[1], p[2], 0.f); }
  tab[(size_t)ye * (W + 2) + xe] = v;
}

static int pcl_image_build(pcl_image* im, const float* img, int h, int w, int format, cudaStream_t st);

extern "C" int pcl_image_create(const float* img, int h, int w, int format, void* stream, pcl_image** out) {
  if (!img || !out || h < 2 || w < 2 || h > 32768 || w > 65536) { pcl_set_error("bad image arguments (%d x %d)", h, w); return PCL_ERR_INVALID; }
  pcl_image* im = (pcl_image*)calloc(1, sizeof(pcl_image));
  if (!im) { pcl_set_error("out of host memory"); return PCL_ERR_INVALID; }
  im->owner = (cudaStream_t)stream;
  const int rc = pcl_image_build(im, img, h, w, format, (cudaStream_t)stream);
  if (rc != PCL_OK) { pcl_image_destroy(im); return rc; }      // nothing leaks on a failed build
  *out = im;
  return PCL_OK;
}

static int pcl_image_build(pcl_image* im, const float* img, int h, int w, int format, cudaStream_t st) {
  int fmt = format;
  if (fmt == PCL_IMAGE_AUTO || fmt == PCL_IMAGE_U8Q || fmt == PCL_IMAGE_U8P || fmt == PCL_IMAGE_TEX || fmt == PCL_IMAGE_F16D) {
    int* flag; int host_flag = 0;
    PCL_CUDA(pcl_pool_alloc((void**)&flag, sizeof(int), st));
    PCL_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    const long long nv = (long long)h * w * 3;
    pcl_u8_exact_kernel<<<(unsigned int)((nv + 255) / 256), 256, 0, st>>>(img, nv, flag);
    PCL_LAUNCH_CHECK();
    PCL_CUDA(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PCL_CUDA(cudaStreamSynchronize(st));
    pcl_pool_free(flag, st);
    if (host_flag) {
      if (fmt != PCL_IMAGE_AUTO) { pcl_set_error("image is not exactly uint8/255: a u8 texel table would change the result"); return PCL_ERR_FORMAT; }
      fmt = PCL_IMAGE_F32;
    } else if (fmt == PCL_IMAGE_AUTO) {
      // Measured on B200 (profiles/r1_texel_format_sizes.md): the 32-byte fp16 basis table is the fastest scoring
      // format up to 2048x4096 (268 MB: it no longer fits the 126 MB L2, but dense clouds share sectors between
      // lanes and the kernels' block order keeps the live slice resident); beyond that the 16-byte quad table wins,
      // and the 4 B/texel texture only remains for panoramas whose quad table would not be reasonable to allocate.
      const size_t entries = (size_t)(h + 1) * (w + 1);
      fmt = (entries * 32 <= (size_t)384 << 20) ? PCL_IMAGE_F16D : (entries * 16 <= (size_t)4096 << 20) ? PCL_IMAGE_U8Q : PCL_IMAGE_TEX;
    }
  } else if (fmt != PCL_IMAGE_F32 && fmt != PCL_IMAGE_TEX) {
    pcl_set_error("unknown image format %d", format);
    return PCL_ERR_INVALID;
  }
  pcl_image_set_geometry(im->view, h, w, (fmt == PCL_IMAGE_U8Q || fmt == PCL_IMAGE_F16D) ? w + 1 : w + 2);
  im->view.fmt = fmt;
  dim3 block(128), grid((w + 2 + 127) / 128, 1);
  if (fmt == PCL_IMAGE_U8Q) {
    im->bytes = (size_t)(h + 1) * (w + 1) * 16; im->view.tex_scale = 1.0f / 255.0f;
    PCL_CUDA(pcl_pool_alloc(&im->data, im->bytes, st));
    grid.y = h + 1;
    pcl_build_u8q_kernel<<<grid, block, 0, st>>>(img, h, w, (uint4*)im->data);
  } else if (fmt == PCL_IMAGE_F16D) {
    im->bytes = (size_t)(h + 1) * (w + 1) * 32; im->view.tex_scale = 1.0f / 255.0f;
    PCL_CUDA(pcl_pool_alloc(&im->data, im->bytes, st));
    grid.y = h + 1;
    if (format == PCL_IMAGE_AUTO) {      // compact companion table for small refinement batches (see pcl_common.cuh)
      // quad entries (16 B) while that table is small (1024x2048: 33.5 MB); plain texels (4 B) for larger panoramas,
      // where a sparse cloud's moving poses miss L2 and the smaller table wins (2048x4096, 1 M points: 52 vs 59 us)
      const bool quad = (size_t)(h + 1) * (w + 1) * 16 <= ((size_t)48 << 20);
      pcl_image_set_geometry(im->view_small, h, w, quad ? w + 1 : w + 2);
      im->view_small.fmt = quad ? PCL_IMAGE_U8Q : PCL_IMAGE_U8P; im->view_small.tex_scale = 1.0f / 255.0f;
      PCL_CUDA(pcl_pool_alloc(&im->data_small, quad ? (size_t)(h + 1) * (w + 1) * 16 : (size_t)(h + 2) * (w + 2) * 4, st));
      if (quad) {
        pcl_build_u8q_kernel<<<grid, block, 0, st>>>(img, h, w, (uint4*)im->data_small);
      } else {
        pcl_build_u8p_kernel<<<dim3(grid.x, h + 2), block, 0, st>>>(img, h, w, (unsigned int*)im->data_small);
      }
      PCL_LAUNCH_CHECK();
      im->view_small.data = im->data_small;
      im->has_small = 1;
    }
    pcl_build_f16d_kernel<<<grid, block, 0, st>>>(img, h, w, (uint4*)im->data);
  } else if (fmt == PCL_IMAGE_U8P) {
    im->bytes = (size_t)(h + 2) * (w + 2) * 4; im->view.tex_scale = 1.0f / 255.0f;
    PCL_CUDA(pcl_pool_alloc(&im->data, im->bytes, st));
    grid.y = h + 2;
    pcl_build_u8p_kernel<<<grid, block, 0, st>>>(img, h, w, (unsigned int*)im->data);
  } else if (fmt == PCL_IMAGE_TEX) {
    // RGBA8 texels in a block-linear cudaArray behind a texture object (point filter, border = 0,
    // unnormalised coordinates, normalised-float reads, gather enabled)
    im->bytes = (size_t)h * w * 4; im->view.tex_scale = 1.0f;
    unsigned int* staging;
    PCL_CUDA(pcl_pool_alloc((void**)&staging, (size_t)(h + 2) * (w + 2) * 4, st));
    grid.y = h + 2;
    pcl_build_u8p_kernel<<<grid, block, 0, st>>>(img, h, w, staging);
    PCL_LAUNCH_CHECK();
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr;
    PCL_CUDA(cudaMallocArray(&arr, &desc, w, h, cudaArrayTextureGather));
    PCL_CUDA(cudaMemcpy2DToArrayAsync(arr, 0, 0, staging + (w + 2) + 1, (size_t)(w + 2) * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    PCL_CUDA(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    PCL_CUDA(cudaStreamSynchronize(st));
    pcl_pool_free(staging, st);
    im->view.tex = (unsigned long long)tex;
    im->data = (void*)arr;
  } else {
    im->bytes = (size_t)(h + 2) * (w + 2) * 16; im->view.tex_scale = 1.0f;
    PCL_CUDA(pcl_pool_alloc(&im->data, im->bytes, st));
    grid.y = h + 2;
    pcl_build_f32_kernel<<<grid, block, 0, st>>>(img, h, w, (float4*)im->data);
  }
  if (fmt != PCL_IMAGE_TEX) PCL_LAUNCH_CHECK();
  // no synchronisation here: the table builds are ordered on `st` like every consumer of the handle (the only host wait
  // of this function is the one-word "is it exact uint8/255 data" answer above, which picks the texel format)
  im->view.data = im->data;
  return PCL_OK;
}

extern "C" int pcl_image_format(const pcl_image* im) { return im ? im->view.fmt : PCL_ERR_INVALID; }

extern "C" void pcl_image_destroy(pcl_image* im) {
  if (!im) return;
  pcl_use_release(im->owner, &im->use);
  if (im->view.fmt == PCL_FMT_TEX) {
    if (im->view.tex) cudaDestroyTextureObject((cudaTextureObject_t)im->view.tex);
    if (im->data) cudaFreeArray((cudaArray_t)im->data);
  } else if (im->data) {
    pcl_pool_free(im->data, im->owner);
  }
  if (im->has_small) pcl_pool_free(im->data_small, im->owner);
  free(im);
}

// ------------------------------------------------------------------------------------------------
// top-k (ascending, ties -> lower index, NaN last): stable radix sort of canonicalised keys
// ------------------------------------------------------------------------------------------------
__global__ void pcl_topk_keys_kernel(const float* __restrict__ loss, long long n, float* keys, long long* idx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = loss[i];
  keys[i] = isnan(v) ? __uint_as_float(0x7fc00000u) : (v == 0.0f ? 0.0f : v);
  idx[i] = i;
}

// Small inputs (the start grids of the shipped protocols: 1 320 - 4 096 poses; the 50 re-ranked candidates): ONE CTA sorts
// 64-bit keys [order-preserving float bits | index] in shared memory with a bitonic network — the index in the low bits
// IS the "ties -> lower index" rule, the canonical NaN maps above +inf.  One launch instead of the ~10 of the radix sort.
#define PCL_TOPK_SMALL 4096
__global__ void __launch_bounds__(1024) pcl_topk_small_kernel(const float* __restrict__ loss, const int n, const int npow2, const int k,
                                                              long long* __restrict__ idx_k) {
  extern __shared__ unsigned long long s_key[];
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const float v = loss[i];
      unsigned int u = isnan(v) ? 0x7fc00000u : (v == 0.0f ? 0u : __float_as_uint(v));
      u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;                // ascending unsigned order == ascending float order
      key = ((unsigned long long)u << 32) | (unsigned int)i;
    }
    s_key[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = s_key[lo], b = s_key[hi];
        if ((a > b) == up) { s_key[lo] = b; s_key[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) idx_k[i] = (long long)(s_key[i] & 0xffffffffull);
}

extern "C" int pcl_topk(const float* loss, int64_t p, int k, int64_t* idx_k, void* stream) {
  if (!loss || !idx_k || p <= 0 || k <= 0 || p > 0x7fffffffll) { pcl_set_error("bad top-k arguments"); return PCL_ERR_INVALID; }
  if (k > p) k = (int)p;
  cudaStream_t st = (cudaStream_t)stream;
  if (p <= PCL_TOPK_SMALL) {
    int npow2 = 2;
    while (npow2 < (int)p) npow2 <<= 1;
    pcl_topk_small_kernel<<<1, 1024, (size_t)npow2 * sizeof(unsigned long long), st>>>(loss, (int)p, npow2, k, (long long*)idx_k);
    PCL_LAUNCH_CHECK();
    return PCL_OK;
  }
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (float*)nullptr, (float*)nullptr, (long long*)nullptr, (long long*)nullptr, (int)p, 0, 32, st);
  const size_t np = (size_t)p;
  const size_t off_tmp = (np * 4 * 2 + np * 8 * 2 + 255) & ~(size_t)255;
  char* buf;
  PCL_CUDA(pcl_pool_alloc((void**)&buf, off_tmp + tmp_bytes + 256, st));
  float* k_in = (float*)buf; float* k_out = k_in + np;
  long long* v_in = (long long*)(buf + np * 8); long long* v_out = v_in + np;
  pcl_topk_keys_kernel<<<(unsigned int)((p + 255) / 256), 256, 0, st>>>(loss, p, k_in, v_in);
  PCL_LAUNCH_CHECK();
  PCL_CUDA(cub::DeviceRadixSort::SortPairs(buf + off_tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)p, 0, 32, st));
  PCL_CUDA(cudaMemcpyAsync(idx_k, v_out, sizeof(long long) * (size_t)k, cudaMemcpyDeviceToDevice, st));
  pcl_pool_free(buf, st);
  return PCL_OK;
}
