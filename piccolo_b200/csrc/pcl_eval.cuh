// Per-(pose, point) evaluation of the PICCOLO sampling loss and its analytic camera-frame gradient.
//
// One call = one pose·point evaluation:  rigid transform (omniloc.py:190-191) -> equirectangular
// projection (utils.py:44-59) -> clip ±0.99 + bilinear sample, align_corners=False, zeros padding
// (utils.py:96-98) -> zero mask + L2 colour residual (omniloc.py:198-200) and, when BWD, the
// gradient w.r.t. the camera-frame point q, accumulated as force a = Σ g_q and
// torque τ = Σ q × g_q  (dL/dangle = ω_angle · τ, dL/dt = -Rᵀ a; see pcl_finish_gradient).
//
// The header is `__host__ __device__` clean: tests/ compile it for the host (tests/emul) to check
// the arithmetic against the oracle without a GPU.  The product only ever runs it on the device.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#if defined(__CUDACC__)
#define PCL_HD __host__ __device__ __forceinline__
#else
#define PCL_HD inline
#endif

#define PCL_PI_F 3.14159265358979323846f
#define PCL_HALF_PI_F 1.57079632679489661923f

// image formats
enum { PCL_FMT_AUTO = 0, PCL_FMT_U8Q = 1, PCL_FMT_F32 = 2, PCL_FMT_U8P = 3, PCL_FMT_TEX = 4, PCL_FMT_F16D = 5 };

struct PclPose {          // R row-major, then t   (48 bytes, 16-byte aligned for LDS.128)
  float r00, r01, r02, r10;
  float r11, r12, r20, r21;
  float r22, tx, ty, tz;
};

struct PclImage {
  const void* data;       // format-dependent texel table (see pcl_image.cu)
  int H, W;
  int pitch;              // entries per row of the padded table
  int fmt;
  float kx, cx;           // ix = cx - phi*kx      (phi = atan2(qy, qx+1e-6) in (-pi, pi])
  float ky, cy;           // iy = theta*ky + cy    (theta = atan2(rho, qz+1e-6) in [0, pi])
  float kxy;              // -kx / ky  (backward: the common factor ky is applied after the reduction)
  float ix_lo, ix_hi, iy_lo, iy_hi;   // the ±0.99 clip expressed in pixel coordinates
  float tex_scale;        // multiplies a blended texel to bring it to [0,1] (1/255 for u8 formats)
  unsigned int idx_bias;  // see pcl_image_set_geometry
  unsigned long long tex; // cudaTextureObject_t (PCL_FMT_TEX only)
};

#define PCL_MAGIC_F 12582912.0f      // 1.5 * 2^23: (x + MAGIC) rounded DOWN leaves MAGIC_I + floor(x) in the mantissa
#define PCL_MAGIC_I 0x4B400000u

// Geometry constants of an H×W panorama (host side; fp32 chain of grid_sampler_unnormalize for the
// clip bounds so that clipped points land on exactly the reference's pixel coordinate).
inline void pcl_image_set_geometry(PclImage& I, int H, int W, int pitch) {
  I.H = H; I.W = W; I.pitch = pitch;
  I.kx = (float)((double)W / (2.0 * 3.14159265358979323846));
  I.cx = 0.5f * (float)W - 0.5f;                     // ix = W/2 - 0.5 - phi*W/(2 pi)
  I.ky = (float)((double)H / 3.14159265358979323846);
  I.cy = -0.5f;                                      // iy = theta*H/pi - 0.5
  I.kxy = -I.kx / I.ky;
  const float c = 0.99f;
  I.ix_hi = ((c + 1.0f) * (float)W - 1.0f) / 2.0f;
  I.ix_lo = ((-c + 1.0f) * (float)W - 1.0f) / 2.0f;
  I.iy_hi = ((c + 1.0f) * (float)H - 1.0f) / 2.0f;
  I.iy_lo = ((-c + 1.0f) * (float)H - 1.0f) / 2.0f;
  // footprint (x0,y0) lives at table entry (y0+1)*pitch + (x0+1); the kernel has the integers
  // MAGIC_I + x0 and MAGIC_I + y0 (mantissa of the magic add), so entry = yi*pitch + xi - bias (mod 2^32)
  I.idx_bias = PCL_MAGIC_I * (unsigned int)pitch + PCL_MAGIC_I - (unsigned int)pitch - 1u;
}

struct PclAcc {           // per-pose running sums
  float se, sm;           // Σ m·e , Σ m
  float ax, ay, az;       // Σ g_q            (BWD only)
  float tx, ty, tz;       // Σ q × g_q        (BWD only)
};

// ---------------------------------------------------------------------------------------------
// fast math primitives (device: single MUFU op; host: libm, for the emulation tests)
// ---------------------------------------------------------------------------------------------
PCL_HD float pcl_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / x;
#endif
}
PCL_HD float pcl_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / sqrtf(x);
#endif
}

// MAGIC + floor(x) for |x| < 2^22: one FADD with round-toward-minus-infinity
PCL_HD float pcl_floor_magic(float x) {
#if defined(__CUDA_ARCH__)
  return __fadd_rd(x, PCL_MAGIC_F);
#else
  return PCL_MAGIC_F + floorf(x);
#endif
}
PCL_HD float pcl_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return sqrtf(x);
#endif
}

// atan(a) for a in [0,1]:  a + a^3 q(a^2), q minimax of degree 6 (max abs error 1.1e-7 in fp32).
PCL_HD float pcl_atan01(float a) {
  const float s = a * a;
  float q = -0.004355400800704956f;
  q = fmaf(q, s, 0.02304011955857277f);
  q = fmaf(q, s, -0.057773567736148834f);
  q = fmaf(q, s, 0.09794232994318008f);
  q = fmaf(q, s, -0.13976581394672394f);
  q = fmaf(q, s, 0.19962704181671143f);
  q = fmaf(q, s, -0.3333165943622589f);
  return fmaf(a * s, q, a);
}

// atan2(y, x) in (-pi, pi]
PCL_HD float pcl_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(fmaxf(ax, ay), 1e-37f);
  const float mn = fminf(ax, ay);
  float r = pcl_atan01(mn * pcl_rcp(mx));
  r = (ay > ax) ? (PCL_HALF_PI_F - r) : r;
  r = (x < 0.0f) ? (PCL_PI_F - r) : r;
  return copysignf(r, y);
}

// atan2(y, x) for y >= 0, result in [0, pi]
PCL_HD float pcl_atan2_pos(float y, float x) {
  const float ax = fabsf(x);
  const float mx = fmaxf(fmaxf(ax, y), 1e-37f);
  const float mn = fminf(ax, y);
  float r = pcl_atan01(mn * pcl_rcp(mx));
  r = (y > ax) ? (PCL_HALF_PI_F - r) : r;
  r = (x < 0.0f) ? (PCL_PI_F - r) : r;
  return r;
}

// ---------------------------------------------------------------------------------------------
// texel fetch: returns the four taps of the footprint whose north-west texel is (x0, y0) with
// x0 in [-1, W-1], y0 in [-1, H-1]; out-of-image taps read the zero border baked into the table.
// Values are in "table units" (0..255 for the u8 formats, 0..1 for f32).
// ---------------------------------------------------------------------------------------------
// Taps come back BIASED: u8 formats return 2^23 + byte (one PRMT each, exact); differences of biased taps
// are exact byte differences, only the north-west tap is un-biased (one FADD per channel).
struct PclTaps { float nw[3], ne[3], sw[3], se[3]; };

template <int FMT> struct PclTapBias { static constexpr float value = 8388608.0f; };
template <> struct PclTapBias<PCL_FMT_F32> { static constexpr float value = 0.0f; };
template <> struct PclTapBias<PCL_FMT_TEX> { static constexpr float value = 0.0f; };

PCL_HD float pcl_u8_biased(uint32_t w, int byte) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + (uint32_t)byte));   // 2^23 + byte
#else
  return 8388608.0f + (float)((w >> (8 * byte)) & 0xffu);
#endif
}

#if defined(__CUDACC__)
typedef uint4 pcl_u4;
typedef float4 pcl_f4;
#else
struct pcl_u4 { uint32_t x, y, z, w; };
struct pcl_f4 { float x, y, z, w; };
#endif
#if defined(__CUDA_ARCH__)
#define PCL_LDG128(p) __ldg(reinterpret_cast<const uint4*>(p))
#define PCL_LDG32(p) __ldg(reinterpret_cast<const uint32_t*>(p))
#define PCL_LDGF4(p) __ldg(reinterpret_cast<const float4*>(p))
#else
#define PCL_LDG128(p) (*reinterpret_cast<const pcl_u4*>(p))
#define PCL_LDG32(p) (*reinterpret_cast<const uint32_t*>(p))
#define PCL_LDGF4(p) (*reinterpret_cast<const pcl_f4*>(p))
#endif

// idx = table entry of the footprint's north-west texel (32-bit; tables have < 2^32 entries)
template <int FMT>
PCL_HD void pcl_fetch(const PclImage& I, unsigned int idx, float x0f, float y0f, PclTaps& t) {
  if (FMT == PCL_FMT_TEX) {
#if defined(__CUDA_ARCH__)
    // RGBA8 block-linear cudaArray read through the texture unit: three 2x2 gathers (one per channel)
    // return the four taps already converted to float (byte/255); no address math, no unpacking.
    // Gather at the footprint centre (x0+1, y0+1) selects texels x0..x0+1, y0..y0+1; border mode gives 0.
    const float gx = x0f + 1.0f, gy = y0f + 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 g4 = tex2Dgather<float4>((cudaTextureObject_t)I.tex, gx, gy, c);   // .w nw  .z ne  .x sw  .y se
      t.nw[c] = g4.w; t.ne[c] = g4.z; t.sw[c] = g4.x; t.se[c] = g4.y;
    }
#endif
  } else if (FMT == PCL_FMT_U8Q) {
    // one 16-byte entry per footprint: {nw, ne, sw, se} as RGBA8
    const pcl_u4 e = PCL_LDG128(reinterpret_cast<const pcl_u4*>(I.data) + idx);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      t.nw[c] = pcl_u8_biased(e.x, c);
      t.ne[c] = pcl_u8_biased(e.y, c);
      t.sw[c] = pcl_u8_biased(e.z, c);
      t.se[c] = pcl_u8_biased(e.w, c);
    }
  } else if (FMT == PCL_FMT_U8P) {
    // plain RGBA8 texels with a 1-texel zero border: 4 B per texel, four 32-bit loads
    const uint32_t* p = reinterpret_cast<const uint32_t*>(I.data) + idx;
    const uint32_t a = PCL_LDG32(p), b = PCL_LDG32(p + 1), c2 = PCL_LDG32(p + I.pitch), d = PCL_LDG32(p + I.pitch + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      t.nw[c] = pcl_u8_biased(a, c);
      t.ne[c] = pcl_u8_biased(b, c);
      t.sw[c] = pcl_u8_biased(c2, c);
      t.se[c] = pcl_u8_biased(d, c);
    }
  } else {
    // fp32 RGBA texels with a 1-texel zero border: 16 B per texel, four 128-bit loads
    const pcl_f4* p = reinterpret_cast<const pcl_f4*>(I.data) + idx;
    const pcl_f4 a = PCL_LDGF4(p), b = PCL_LDGF4(p + 1), c2 = PCL_LDGF4(p + I.pitch), d = PCL_LDGF4(p + I.pitch + 1);
    t.nw[0] = a.x; t.nw[1] = a.y; t.nw[2] = a.z;
    t.ne[0] = b.x; t.ne[1] = b.y; t.ne[2] = b.z;
    t.sw[0] = c2.x; t.sw[1] = c2.y; t.sw[2] = c2.z;
    t.se[0] = d.x; t.se[1] = d.y; t.se[2] = d.z;
  }
}

// Bilinear basis of a footprint: s(fx, fy) = nw + fx·dxt + fy·(dy0 + fx·ddx)
//   dxt = ne - nw, dy0 = sw - nw, ddx = (se - sw) - (ne - nw);   ds/dfx = dxt + fy·ddx, ds/dfy = dy0 + fx·ddx
struct PclBasis { float nw[3], dxt[3], dy0[3], ddx[3]; };

// fp16 <-> fp32 for the F16D table (values are integers in [-510, 510]: exact in fp16)
PCL_HD float pcl_half_lo(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu)));
#else
  const uint32_t h = w & 0xffffu, sgn = (h & 0x8000u) << 16, ex = (h >> 10) & 0x1fu, man = h & 0x3ffu;
  if (ex == 0 && man == 0) { float f; uint32_t b = sgn; memcpy(&f, &b, 4); return f; }
  uint32_t b = sgn | ((ex + 112u) << 23) | (man << 13); float f; memcpy(&f, &b, 4); return f;   // normals only
#endif
}
PCL_HD float pcl_half_hi(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return __half2float(__ushort_as_half((unsigned short)(w >> 16)));
#else
  return pcl_half_lo(w >> 16);
#endif
}

// Raw footprint data as it comes back from memory: fetching (pcl_fetch_raw) is split from unpacking
// (pcl_basis_from_raw) so that kernels can issue the loads of several evaluations before using any of them.
template <int FMT> struct PclRaw { PclTaps t; };                       // U8P, F32, TEX: the four taps, converted
template <> struct PclRaw<PCL_FMT_U8Q> { pcl_u4 e; };                  // {nw, ne, sw, se} RGBA8
template <> struct PclRaw<PCL_FMT_F16D> { pcl_u4 lo, hi; };            // 16 halves: the basis itself

template <int FMT>
PCL_HD void pcl_fetch_raw(const PclImage& I, unsigned int idx, float x0f, float y0f, PclRaw<FMT>& r) {
  if constexpr (FMT == PCL_FMT_F16D) {
    // 32-byte entry per footprint: for each channel the four basis values as fp16 (exact small integers):
    // one 256-bit load (one 32-byte sector), one conversion per value, no differences to form
    const pcl_u4* e = reinterpret_cast<const pcl_u4*>(I.data) + 2 * (size_t)idx;
#if defined(__CUDA_ARCH__)
    // sm_100 256-bit load (LDG.E.ENL2.256): the whole 32-byte entry in ONE request per lane
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w) : "l"(e));
#else
    r.lo = PCL_LDG128(e); r.hi = PCL_LDG128(e + 1);
#endif
  } else if constexpr (FMT == PCL_FMT_U8Q) {
    r.e = PCL_LDG128(reinterpret_cast<const pcl_u4*>(I.data) + idx);   // one 16-byte entry per footprint
  } else {
    pcl_fetch<FMT>(I, idx, x0f, y0f, r.t);
  }
}

template <int FMT>
PCL_HD void pcl_basis_from_raw(const PclRaw<FMT>& r, PclBasis& b) {
  if constexpr (FMT == PCL_FMT_F16D) {
    b.nw[0] = pcl_half_lo(r.lo.x); b.dxt[0] = pcl_half_hi(r.lo.x); b.dy0[0] = pcl_half_lo(r.lo.y); b.ddx[0] = pcl_half_hi(r.lo.y);
    b.nw[1] = pcl_half_lo(r.lo.z); b.dxt[1] = pcl_half_hi(r.lo.z); b.dy0[1] = pcl_half_lo(r.lo.w); b.ddx[1] = pcl_half_hi(r.lo.w);
    b.nw[2] = pcl_half_lo(r.hi.x); b.dxt[2] = pcl_half_hi(r.hi.x); b.dy0[2] = pcl_half_lo(r.hi.y); b.ddx[2] = pcl_half_hi(r.hi.y);
  } else {
    PclTaps t;
    if constexpr (FMT == PCL_FMT_U8Q) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t.nw[c] = pcl_u8_biased(r.e.x, c);
        t.ne[c] = pcl_u8_biased(r.e.y, c);
        t.sw[c] = pcl_u8_biased(r.e.z, c);
        t.se[c] = pcl_u8_biased(r.e.w, c);
      }
    } else {
      t = r.t;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      b.nw[c] = t.nw[c] - PclTapBias<FMT>::value;          // the only un-biasing
      b.dxt[c] = t.ne[c] - t.nw[c];                        // exact differences of biased taps
      b.dy0[c] = t.sw[c] - t.nw[c];
      b.ddx[c] = (t.se[c] - t.sw[c]) - b.dxt[c];
    }
  }
}

template <int FMT>
PCL_HD void pcl_fetch_basis(const PclImage& I, unsigned int idx, float x0f, float y0f, PclBasis& b) {
  PclRaw<FMT> r;
  pcl_fetch_raw<FMT>(I, idx, x0f, y0f, r);
  pcl_basis_from_raw<FMT>(r, b);
}

// ---------------------------------------------------------------------------------------------
// one pose·point evaluation, in two halves: A = rigid transform, projection, footprint address;
// B = blend, mask, residual and (BWD) the camera-frame gradient.  The fetch sits between the two: the fused refinement
// (pcl_refine.cuh) issues it as an asynchronous copy into shared memory right after A and runs B one pipeline stage
// later, so the L2 latency of the texel gather never stalls a warp.
// ---------------------------------------------------------------------------------------------
struct PclMid {
  float qx, qy, qz, rinv;       // camera-frame point, 1/rho (BWD)
  float fx, fy;                 // fractional pixel coordinates
  bool px_pass, py_pass;        // the clip passed the coordinate through (gradient flows)
};
struct PclAddr { unsigned int idx; float x0f, y0f; };   // footprint: table entry + integer pixel coordinates

template <bool BWD>
PCL_HD void pcl_eval_a(const PclPose& P, const PclImage& I, float px, float py, float pz, PclMid& M, PclAddr& A) {
  // q = R (p - t)
  const float dx = px - P.tx, dy = py - P.ty, dz = pz - P.tz;
  const float qx = fmaf(P.r02, dz, fmaf(P.r01, dy, P.r00 * dx));
  const float qy = fmaf(P.r12, dz, fmaf(P.r11, dy, P.r10 * dx));
  const float qz = fmaf(P.r22, dz, fmaf(P.r21, dy, P.r20 * dx));
  const float xp = qx + 1e-6f, zp = qz + 1e-6f;
  const float rho2 = fmaf(qx, qx, qy * qy);
  float rinv = 0.0f, rho;
  if (BWD) { rinv = pcl_rsqrt(fmaxf(rho2, 1e-37f)); rho = rho2 * rinv; } else { rho = pcl_sqrt(rho2); }

  const float phi = pcl_atan2(qy, xp);
  const float theta = pcl_atan2_pos(rho, zp);
  // pixel coordinates, clipped; floor by a round-down magic add (mantissa = MAGIC_I + floor)
  const float ix_raw = fmaf(-phi, I.kx, I.cx);
  const float iy_raw = fmaf(theta, I.ky, I.cy);
  const float ix = fminf(fmaxf(ix_raw, I.ix_lo), I.ix_hi);
  const float iy = fminf(fmaxf(iy_raw, I.iy_lo), I.iy_hi);
  const float tx = pcl_floor_magic(ix), ty = pcl_floor_magic(iy);
  M.fx = ix - (tx - PCL_MAGIC_F);                               // exact fractional parts
  M.fy = iy - (ty - PCL_MAGIC_F);
#if defined(__CUDA_ARCH__)
  const unsigned int xi = __float_as_uint(tx), yi = __float_as_uint(ty);
#else
  unsigned int xi, yi; { float a = tx, b = ty; memcpy(&xi, &a, 4); memcpy(&yi, &b, 4); }
#endif
  A.idx = yi * (unsigned int)I.pitch + xi - I.idx_bias;
  A.x0f = tx - PCL_MAGIC_F; A.y0f = ty - PCL_MAGIC_F;
  M.qx = qx; M.qy = qy; M.qz = qz; M.rinv = rinv;
  M.px_pass = (ix_raw == ix);                                   // clip passes the gradient inclusively at the bound
  M.py_pass = (iy_raw == iy);
}

template <int FMT, bool BWD>
PCL_HD void pcl_eval_b(const PclImage& I, const PclMid& M, const PclRaw<FMT>& raw, float cr, float cg, float cb, bool valid, PclAcc& acc) {   // valid: false for padding points
  const float fx = M.fx, fy = M.fy;
  PclBasis b;
  pcl_basis_from_raw<FMT>(raw, b);

  float d[3], dsdx[3], dsdy[3], ssum = -0.0f;   // -0.0f + x == x exactly: the first add folds away
  bool all_zero = true;
  const float col[3] = {cr, cg, cb};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float top = fmaf(fx, b.dxt[c], b.nw[c]);
    const float dyv = fmaf(fx, b.ddx[c], b.dy0[c]);        // bottom - top
    const float s = fmaf(fy, dyv, top);
    ssum += s;
    if (FMT == PCL_FMT_F32) all_zero = all_zero && (s == 0.0f);
    d[c] = fmaf(s, I.tex_scale, -col[c]);
    if (BWD) {
      dsdx[c] = fmaf(fy, b.ddx[c], b.dxt[c]);
      dsdy[c] = dyv;
    }
  }
  // zero mask (omniloc.py:198: all three channels == 0).  The u8 tables hold texels >= 0, where Σ == 0 <=> all
  // three are 0 (one compare); fp32 tables take arbitrary floats, negative ones included: channel-wise test
  const bool m = valid && (FMT == PCL_FMT_F32 ? !all_zero : (ssum > 0.0f));
  const float e2 = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
  float einv = 0.0f, e;
  if (BWD) { einv = pcl_rsqrt(fmaxf(e2, 1e-37f)); e = e2 * einv; } else { e = pcl_sqrt(e2); }
  acc.se += m ? e : 0.0f;
  acc.sm += m ? 1.0f : 0.0f;

  if (BWD) {
    const float qx = M.qx, qy = M.qy, qz = M.qz, rinv = M.rinv;
    const float xp = qx + 1e-6f, zp = qz + 1e-6f;
    const float rho2 = fmaf(qx, qx, qy * qy);
    const float rho = rho2 * rinv;
    // g_s = m (s - c)/e ; texel-unit and pixel-unit factors (tex_scale, W/2·(-1/pi), H/2·(2/pi)) are
    // constants of the pose and are applied once to the reduced sums (pcl_finish_gradient).
    const float w = m ? einv : 0.0f;
    float gix = fmaf(d[2], dsdx[2], fmaf(d[1], dsdx[1], d[0] * dsdx[0])) * w;
    float giy = fmaf(d[2], dsdy[2], fmaf(d[1], dsdy[1], d[0] * dsdy[0])) * w;
    gix = M.px_pass ? gix : 0.0f;
    giy = M.py_pass ? giy : 0.0f;
    const float dphi = fmaf(xp, xp, qy * qy);
    const float dth = fmaf(zp, zp, rho2);
    // one reciprocal for both denominators.  (dphi or dth == 0 needs qx == -1e-6 exactly; the reference's
    // atan2 backward is 0/0 there as well.)
    const float rr = pcl_rcp(dphi * dth);
    // ix = cx - kx·phi, iy = ky·theta + cy   =>   g_phi = -kx·g_ix, g_theta = ky·g_iy
    // phi = atan2(qy, xp):   dphi/dqx = -qy/dphi, dphi/dqy = xp/dphi
    // theta = atan2(rho, zp): dtheta/drho = zp/dth, dtheta/dzp = -rho/dth ; drho/dq{x,y} = q{x,y}/rho
    // Everything is accumulated DIVIDED BY ky (kxy = -kx/ky); ky is applied once in pcl_finish_gradient.
    const float A = (I.kxy * gix) * (rr * dth);         // g_phi / (xp² + qy²) / ky
    const float B = giy * (rr * dphi);                  // g_theta / (rho² + zp²) / ky
    const float Bz = B * zp * rinv;                     // rho == 0  ->  multiplied by qx = qy = 0 below
    const float gqx = fmaf(Bz, qx, -A * qy);
    const float gqy = fmaf(Bz, qy, A * xp);
    const float gqz = -B * rho;
    acc.ax += gqx; acc.ay += gqy; acc.az += gqz;
    acc.tx = fmaf(qy, gqz, fmaf(-qz, gqy, acc.tx));     // τ += q × g_q
    acc.ty = fmaf(qz, gqx, fmaf(-qx, gqz, acc.ty));
    acc.tz = fmaf(qx, gqy, fmaf(-qy, gqx, acc.tz));
  }
}

template <int FMT, bool BWD>
PCL_HD void pcl_eval(const PclPose& P, const PclImage& I, float px, float py, float pz,
                     float cr, float cg, float cb, bool valid, PclAcc& acc) {
  PclMid M;
  PclAddr A;
  PclRaw<FMT> raw;
  pcl_eval_a<BWD>(P, I, px, py, pz, M, A);
  pcl_fetch_raw<FMT>(I, A.idx, A.x0f, A.y0f, raw);
  pcl_eval_b<FMT, BWD>(I, M, raw, cr, cg, cb, valid, acc);
}

// ---------------------------------------------------------------------------------------------
// Structured start-pose grids (forward only).  Rotations that differ by an in-plane rotation about the camera's
// z axis, R_j = Rz(delta_j)·R_base, see every point at the same elevation theta and at azimuth phi_base + delta_j:
// all yaw-only grids are one such group, the 24 distinct rotations of the 4x4x4 Euler lattice are 6 groups of 4.
// pcl_grid_base does the rigid transform, rho, theta, the row coordinate and phi_base ONCE per (point, translation,
// group); pcl_grid_member finishes one member rotation (azimuth shift, column coordinate, fetch, blend, residual).
// Deviation from evaluating every pose separately: the `+1e-6` of cloud2idx is applied in the base frame only,
// which moves phi by <= 1e-6/rho rad — 1.6e-7 relative on the loss (measured against the reference, fp64).
// ---------------------------------------------------------------------------------------------
struct PclGridBase {
  float phi;              // atan2(qy, qx + 1e-6) in the base frame, (-pi, pi]
  float fy, y0f;          // fractional / integer part of the clipped row coordinate
  unsigned int yrow;      // yi * pitch - idx_bias: the table index is yrow + xi
};

PCL_HD void pcl_grid_base(const PclPose& P, const PclImage& I, float px, float py, float pz, PclGridBase& b) {
  const float dx = px - P.tx, dy = py - P.ty, dz = pz - P.tz;
  const float qx = fmaf(P.r02, dz, fmaf(P.r01, dy, P.r00 * dx));
  const float qy = fmaf(P.r12, dz, fmaf(P.r11, dy, P.r10 * dx));
  const float qz = fmaf(P.r22, dz, fmaf(P.r21, dy, P.r20 * dx));
  const float rho = pcl_sqrt(fmaf(qx, qx, qy * qy));
  b.phi = pcl_atan2(qy, qx + 1e-6f);
  const float theta = pcl_atan2_pos(rho, qz + 1e-6f);
  const float iy = fminf(fmaxf(fmaf(theta, I.ky, I.cy), I.iy_lo), I.iy_hi);
  const float ty = pcl_floor_magic(iy);
  b.y0f = ty - PCL_MAGIC_F;
  b.fy = iy - b.y0f;
  unsigned int yi;
#if defined(__CUDA_ARCH__)
  yi = __float_as_uint(ty);
#else
  { float a = ty; memcpy(&yi, &a, 4); }
#endif
  b.yrow = yi * (unsigned int)I.pitch - I.idx_bias;
}

template <int FMT>
PCL_HD void pcl_grid_member(const PclImage& I, const PclGridBase& b, float delta, float cr, float cg, float cb, bool valid,
                            float& se, float& sm) {
  // azimuth of the member rotation, wrapped back into (-pi, pi] (delta is given in (-pi, pi])
  float phi = b.phi + delta;
  phi = (phi > PCL_PI_F) ? (phi - 2.0f * PCL_PI_F) : phi;
  phi = (phi <= -PCL_PI_F) ? (phi + 2.0f * PCL_PI_F) : phi;
  const float ix = fminf(fmaxf(fmaf(-phi, I.kx, I.cx), I.ix_lo), I.ix_hi);
  const float tx = pcl_floor_magic(ix);
  const float x0f = tx - PCL_MAGIC_F;
  const float fx = ix - x0f;
  unsigned int xi;
#if defined(__CUDA_ARCH__)
  xi = __float_as_uint(tx);
#else
  { float a = tx; memcpy(&xi, &a, 4); }
#endif
  PclBasis bs;
  pcl_fetch_basis<FMT>(I, b.yrow + xi, x0f, b.y0f, bs);
  float d[3], ssum = -0.0f;
  bool all_zero = true;
  const float col[3] = {cr, cg, cb};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float top = fmaf(fx, bs.dxt[c], bs.nw[c]);
    const float dyv = fmaf(fx, bs.ddx[c], bs.dy0[c]);
    const float s = fmaf(b.fy, dyv, top);
    ssum += s;
    if (FMT == PCL_FMT_F32) all_zero = all_zero && (s == 0.0f);
    d[c] = fmaf(s, I.tex_scale, -col[c]);
  }
  const bool m = valid && (FMT == PCL_FMT_F32 ? !all_zero : (ssum > 0.0f));
  const float e = pcl_sqrt(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0])));
  se += m ? e : 0.0f;
  sm += m ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// pose set-up and gradient finish (one thread per pose)
// ---------------------------------------------------------------------------------------------
// R = (Rz·Ry)·Rx in fp32 like the reference's two torch.mm calls (omniloc.py:187-188), from the six trigonometric values
PCL_HD void pcl_pose_from_trig(const float cy, const float sy, const float cp, const float sp, const float cr, const float sr,
                               const float tx, const float ty, const float tz, PclPose& P) {
  const float m00 = cy * cp, m01 = -sy, m02 = cy * sp;
  const float m10 = sy * cp, m11 = cy, m12 = sy * sp;
  const float m20 = -sp, m21 = 0.0f, m22 = cp;
  P.r00 = m00; P.r01 = m01 * cr + m02 * sr; P.r02 = m02 * cr - m01 * sr;
  P.r10 = m10; P.r11 = m11 * cr + m12 * sr; P.r12 = m12 * cr - m11 * sr;
  P.r20 = m20; P.r21 = m21 * cr + m22 * sr; P.r22 = m22 * cr - m21 * sr;
  P.tx = tx; P.ty = ty; P.tz = tz;
}

PCL_HD void pcl_pose_from_params(const float* p6, PclPose& P) {
  pcl_pose_from_trig(cosf(p6[3]), sinf(p6[3]), cosf(p6[4]), sinf(p6[4]), cosf(p6[5]), sinf(p6[5]), p6[0], p6[1], p6[2], P);
}

// sums[8] = {Σ m e, Σ m, a(3), τ(3)} (already reduced over all points)  ->  loss, grad[6]
PCL_HD void pcl_finish_gradient(const float* p6, const PclPose& P, const PclImage& I, const double* sums,
                                float* loss, float* count, float* grad6) {
  const double M = sums[1];
  *loss = (float)(sums[0] / M);               // 0/0 -> NaN, the reference's empty mean
  if (count) *count = (float)M;
  if (!grad6) return;
  const double k = (double)I.tex_scale * (double)I.ky / M;   // g_s carries tex_scale; sums were accumulated / ky
  const double ax = sums[2] * k, ay = sums[3] * k, az = sums[4] * k;
  const double tx = sums[5] * k, ty = sums[6] * k, tz = sums[7] * k;
  // dL/dt = -Rᵀ a
  grad6[0] = (float)(-(P.r00 * ax + P.r10 * ay + P.r20 * az));
  grad6[1] = (float)(-(P.r01 * ax + P.r11 * ay + P.r21 * az));
  grad6[2] = (float)(-(P.r02 * ax + P.r12 * ay + P.r22 * az));
  // dL/dangle = ω·τ with ω_yaw = z, ω_pitch = Rz·y, ω_roll = Rz·Ry·x
  const double cy = (double)cosf(p6[3]), sy = (double)sinf(p6[3]);    // fp32 trig, like the reference's rotation
  const double cp = (double)cosf(p6[4]), sp = (double)sinf(p6[4]);
  grad6[3] = (float)tz;
  grad6[4] = (float)(-sy * tx + cy * ty);
  grad6[5] = (float)(cy * cp * tx + sy * cp * ty - sp * tz);
}
