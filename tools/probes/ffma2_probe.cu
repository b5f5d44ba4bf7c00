// ffma2_probe.cu - packed FP32 (fma.rn.f32x2, SASS FFMA2) on sm_100a: latency of a dependent chain and issue throughput against scalar FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
// Question behind it: 57 % of the physics launch's instructions are FFMA / FMUL / FADD on 3-vectors and 3 x 3 blocks; would
// pairing them shorten one warp's dependent instruction stream?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int MODE, int ILP>
__global__ void k(int iters, float a, float b, float* out, long long* cyc) {
  // MODE 0: scalar FFMA on 2 * ILP independent chains (same flops as MODE 1); MODE 1: FFMA2 on ILP independent chains of pairs
  float2 x[ILP];
  for (int j = 0; j < ILP; j++) x[j] = make_float2(threadIdx.x * 0.001f + j, threadIdx.x * 0.002f - j);
  const float2 A = make_float2(a, a * 1.0001f), Bv = make_float2(b, b * 0.9999f);
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      if (MODE == 0) { x[j].x = fmaf(x[j].x, A.x, Bv.x); x[j].y = fmaf(x[j].y, A.y, Bv.y); }
      else x[j] = __ffma2_rn(x[j], A, Bv);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < ILP; j++) s += x[j].x + x[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ILP> int run(int warps, const char* name) {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMallocManaged(&cyc, 8));
  const int iters = 4096;
  k<MODE, ILP><<<148, warps * 32>>>(iters, 0.999f, 0.5f, out, cyc);
  CK(cudaDeviceSynchronize());
  const double per_iter = (double)*cyc / iters;
  printf("%-12s ILP %d, %2d warps/SM: %.2f cycles per iteration = %.2f cycles per pair-FMA (%.1f lane-FMAs per cycle per SM)\n", name, ILP, warps, per_iter,
         per_iter / ILP, 2.0 * ILP * warps * 32 / per_iter);
  cudaFree(out); cudaFree(cyc);
  return 0;
}
int main() {
  for (int w : {1, 4, 8, 16}) {
    if (w == 1) { run<0, 1>(w, "FFMA x2"); run<1, 1>(w, "FFMA2"); }      // dependent chain: latency
    if (w == 1) { run<0, 4>(w, "FFMA x2"); run<1, 4>(w, "FFMA2"); }
    if (w == 4) { run<0, 1>(w, "FFMA x2"); run<1, 1>(w, "FFMA2"); run<0, 4>(w, "FFMA x2"); run<1, 4>(w, "FFMA2"); }
    if (w == 8) { run<0, 4>(w, "FFMA x2"); run<1, 4>(w, "FFMA2"); }
    if (w == 16) { run<0, 4>(w, "FFMA x2"); run<1, 4>(w, "FFMA2"); }
  }
  return 0;
}
