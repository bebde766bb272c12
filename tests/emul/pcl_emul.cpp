// Host emulation of the device arithmetic in piccolo_b200/csrc/pcl_eval.cuh  —  TEST ONLY.
// Compiled by tests/test_emul_math.py with g++ so that the per-point maths (projection, bilinear
// taps, analytic gradient, finish) can be checked against the oracle on a machine without a GPU.
// The product library never links or loads this file.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../piccolo_b200/csrc/pcl_eval.cuh"

static inline uint32_t pack_rgba(const float* img, int H, int W, int y, int x) {
  if (x < 0 || y < 0 || x >= W || y >= H) return 0u;
  const float* p = img + ((size_t)y * W + x) * 3;
  uint32_t r = (uint32_t)lrintf(p[0] * 255.0f), g = (uint32_t)lrintf(p[1] * 255.0f), b = (uint32_t)lrintf(p[2] * 255.0f);
  return r | (g << 8) | (b << 16);
}

template <int FMT, bool BWD>
static void run(const PclImage& I, const float* xyz, const float* rgb, long n, const float* poses, int P,
                float* loss, float* cnt, float* grad) {
  for (int p = 0; p < P; ++p) {
    PclPose pose; pcl_pose_from_params(poses + 6 * p, pose);
    double sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long i = 0; i < n; ++i) {
      PclAcc a; std::memset(&a, 0, sizeof(a));
      pcl_eval<FMT, BWD>(pose, I, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], true, a);
      sums[0] += a.se; sums[1] += a.sm; sums[2] += a.ax; sums[3] += a.ay; sums[4] += a.az; sums[5] += a.tx; sums[6] += a.ty; sums[7] += a.tz;
    }
    pcl_finish_gradient(poses + 6 * p, pose, I, sums, loss + p, cnt + p, BWD ? grad + 6 * p : nullptr);
  }
}

// structured grid: losses of (translation t, base rotation ypr) for members with azimuth offsets delta[]
template <int FMT>
static void run_grid(const PclImage& I, const float* xyz, const float* rgb, long n, const float* base_pose6, const float* delta, int nm,
                     float* loss, float* cnt) {
  PclPose pose; pcl_pose_from_params(base_pose6, pose);
  for (int m = 0; m < nm; ++m) {
    double se = 0, sm = 0;
    for (long i = 0; i < n; ++i) {
      PclGridBase b; pcl_grid_base(pose, I, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], b);
      float e = 0.f, c = 0.f;
      pcl_grid_member<FMT>(I, b, delta[m], rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], true, e, c);
      se += e; sm += c;
    }
    loss[m] = (float)(se / sm); cnt[m] = (float)sm;
  }
}

extern "C" int emul_grid(const float* xyz, const float* rgb, long n, const float* img, int H, int W, const float* base_pose6,
                         const float* delta, int nm, float* loss, float* cnt) {
  PclImage I; std::memset(&I, 0, sizeof(I));
  pcl_image_set_geometry(I, H, W, W + 2);
  I.fmt = PCL_FMT_F32; I.tex_scale = 1.0f;
  std::vector<float> f32((size_t)(H + 2) * (W + 2) * 4, 0.0f);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
    for (int c = 0; c < 3; ++c) f32[((size_t)(y + 1) * (W + 2) + (x + 1)) * 4 + c] = img[((size_t)y * W + x) * 3 + c];
  I.data = f32.data();
  run_grid<PCL_FMT_F32>(I, xyz, rgb, n, base_pose6, delta, nm, loss, cnt);
  return 0;
}

extern "C" int emul_loss_grad(const float* xyz, const float* rgb, long n, const float* img, int H, int W, int fmt,
                              const float* poses, int P, int bwd, float* loss, float* cnt, float* grad) {
  PclImage I; std::memset(&I, 0, sizeof(I));
  pcl_image_set_geometry(I, H, W, (fmt == PCL_FMT_U8Q || fmt == PCL_FMT_F16D) ? W + 1 : W + 2);
  I.fmt = fmt;
  std::vector<uint32_t> u8;
  std::vector<float> f32;
  if (fmt == PCL_FMT_U8Q) {
    I.tex_scale = 1.0f / 255.0f;
    u8.resize((size_t)(H + 1) * (W + 1) * 4);
    for (int y0 = -1; y0 < H; ++y0) for (int x0 = -1; x0 < W; ++x0) {
      uint32_t* e = &u8[((size_t)(y0 + 1) * (W + 1) + (x0 + 1)) * 4];
      e[0] = pack_rgba(img, H, W, y0, x0); e[1] = pack_rgba(img, H, W, y0, x0 + 1);
      e[2] = pack_rgba(img, H, W, y0 + 1, x0); e[3] = pack_rgba(img, H, W, y0 + 1, x0 + 1);
    }
    I.data = u8.data();
  } else if (fmt == PCL_FMT_F16D) {
    I.tex_scale = 1.0f / 255.0f;
    u8.resize((size_t)(H + 1) * (W + 1) * 8);
    auto h16 = [](int v) -> uint32_t {            // exact fp16 bits of a small integer
      if (v == 0) return 0u;
      uint32_t sgn = v < 0 ? 0x8000u : 0u; uint32_t a = (uint32_t)(v < 0 ? -v : v); int e = 0;
      while ((a >> (e + 1)) != 0) ++e;            // a in [2^e, 2^(e+1))
      uint32_t man = (a << (10 - e)) & 0x3ffu;
      return sgn | ((uint32_t)(e + 15) << 10) | man;
    };
    for (int y0 = -1; y0 < H; ++y0) for (int x0 = -1; x0 < W; ++x0) {
      uint32_t t[4] = {pack_rgba(img, H, W, y0, x0), pack_rgba(img, H, W, y0, x0 + 1), pack_rgba(img, H, W, y0 + 1, x0), pack_rgba(img, H, W, y0 + 1, x0 + 1)};
      uint32_t* e = &u8[((size_t)(y0 + 1) * (W + 1) + (x0 + 1)) * 8];
      for (int c = 0; c < 3; ++c) {
        int nw = (t[0] >> (8 * c)) & 0xff, ne = (t[1] >> (8 * c)) & 0xff, sw = (t[2] >> (8 * c)) & 0xff, se = (t[3] >> (8 * c)) & 0xff;
        e[2 * c] = h16(nw) | (h16(ne - nw) << 16);
        e[2 * c + 1] = h16(sw - nw) | (h16((se - sw) - (ne - nw)) << 16);
      }
      e[6] = e[7] = 0;
    }
    I.data = u8.data();
  } else if (fmt == PCL_FMT_U8P) {
    I.tex_scale = 1.0f / 255.0f;
    u8.resize((size_t)(H + 2) * (W + 2));
    for (int y = -1; y <= H; ++y) for (int x = -1; x <= W; ++x) u8[(size_t)(y + 1) * (W + 2) + (x + 1)] = pack_rgba(img, H, W, y, x);
    I.data = u8.data();
  } else {
    I.tex_scale = 1.0f;
    f32.assign((size_t)(H + 2) * (W + 2) * 4, 0.0f);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
      for (int c = 0; c < 3; ++c) f32[((size_t)(y + 1) * (W + 2) + (x + 1)) * 4 + c] = img[((size_t)y * W + x) * 3 + c];
    I.data = f32.data();
  }
#define GO(F) do { if (bwd) run<F, true>(I, xyz, rgb, n, poses, P, loss, cnt, grad); else run<F, false>(I, xyz, rgb, n, poses, P, loss, cnt, grad); } while (0)
  if (fmt == PCL_FMT_U8Q) GO(PCL_FMT_U8Q); else if (fmt == PCL_FMT_U8P) GO(PCL_FMT_U8P); else if (fmt == PCL_FMT_F16D) GO(PCL_FMT_F16D); else GO(PCL_FMT_F32);
  return 0;
}
