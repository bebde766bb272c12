// Microbenchmark: does a packed fp32x2 FMA (FFMA2) save issue slots on sm_100a, and what do ALU-pipe ops cost?
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }                          // 2 FFMA
      if (MODE == 1) { x[i] = __ffma2_rn(x[i], A, B); }                                                       // 1 FFMA2
      if (MODE == 2) { x[i] = __ffma2_rn(x[i], A, B); x[i].x = fminf(x[i].x, 3.0f); x[i].y = fminf(x[i].y, 3.0f); }   // FFMA2 + 2 FMNMX
      if (MODE == 3) { x[i] = __ffma2_rn(x[i], A, B); x[i].x = fminf(x[i].x, 3.0f); }                         // FFMA2 + 1 FMNMX
      if (MODE == 4) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); x[i].x = fminf(x[i].x, 3.0f); x[i].y = fminf(x[i].y, 3.0f); }  // 2 FFMA + 2 FMNMX
      if (MODE == 5) { x[i].x = fminf(x[i].x * 1.0f, 3.0f + i); x[i].y = fminf(x[i].y, 2.0f + it); }           // ~2 FMNMX (+1 FMUL)
      if (MODE == 6) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); x[i].x = fminf(x[i].x, 3.0f); }  // 2 FFMA + 1 FMNMX
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int M> float run(float* out, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<M><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<M><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const char* names[] = {"2 FFMA", "1 FFMA2", "FFMA2 + 2 FMNMX", "FFMA2 + 1 FMNMX", "2 FFMA + 2 FMNMX", "2 FMNMX + FMUL", "2 FFMA + 1 FMNMX"};
  float ms[7] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters), run<4>(out, iters), run<5>(out, iters), run<6>(out, iters)};
  // cycles per inner statement per SMSP: 8 CTAs x 8 warps per SM = 16 warps per SMSP
  for (int m = 0; m < 7; ++m) printf("%-18s %.3f ms  -> %.2f SMSP-cycles per statement per warp (at 1.965 GHz)\n", names[m], ms[m], ms[m] * 1e-3 * 1.965e9 / (16.0 * 8 * iters));
  return 0;
}
