// tmem_probe.cu - is tensor memory usable as a lane-private scratch store for a non-GEMM kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
// (1) correctness of tcgen05.st / tcgen05.ld 32x32b with per-warp column windows and mixed shapes (x1, x2, x4, x8), run-time block
//     offsets, 14 warps per CTA;  (2) round-trip latency of a dependent  ld -> wait -> math -> st -> wait  chain against the same
//     chain through shared memory, for 1 / 4 / 7 / 14 resident warps;  (3) whether column offsets have to be aligned to the shape.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N> __device__ __forceinline__ void tm_ld(uint32_t a, float* r);
template <int N> __device__ __forceinline__ void tm_st(uint32_t a, const float* r);
template <> __device__ __forceinline__ void tm_ld<1>(uint32_t a, float* r) {
  uint32_t x; asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(a)); r[0] = __uint_as_float(x); }
template <> __device__ __forceinline__ void tm_ld<2>(uint32_t a, float* r) {
  uint32_t x, y; asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
  r[0] = __uint_as_float(x); r[1] = __uint_as_float(y); }
template <> __device__ __forceinline__ void tm_ld<4>(uint32_t a, float* r) {
  uint32_t x[4]; asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(a));
  for (int k = 0; k < 4; k++) r[k] = __uint_as_float(x[k]); }
template <> __device__ __forceinline__ void tm_ld<8>(uint32_t a, float* r) {
  uint32_t x[8]; asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "r"(a));
  for (int k = 0; k < 8; k++) r[k] = __uint_as_float(x[k]); }
template <> __device__ __forceinline__ void tm_st<1>(uint32_t a, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(__float_as_uint(r[0])) : "memory"); }
template <> __device__ __forceinline__ void tm_st<2>(uint32_t a, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])) : "memory"); }
template <> __device__ __forceinline__ void tm_st<4>(uint32_t a, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])),
               "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])) : "memory"); }
template <> __device__ __forceinline__ void tm_st<8>(uint32_t a, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(__float_as_uint(r[0])),
               "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])),
               "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])) : "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float val(int w, int lane, int col) { return (float)(w * 100000 + lane * 1000 + col) + 0.25f; }

// mode 0: correctness (aligned layout of the physics kernel: 3 blocks x 40 columns); mode 1: unaligned shapes
__global__ void __launch_bounds__(448, 1) k_check(int mode, int* err, int* first) {
  __shared__ uint32_t s_base;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + (((uint32_t)(w & 3) * 32u) << 16) + (uint32_t)(w >> 2) * 128u;
  int bad = 0;
  if (mode == 0) {
    for (int blk = 0; blk < 3; blk++) {   // run-time block offset
      float r[40];
      for (int k = 0; k < 40; k++) r[k] = val(w, lane, blk * 40 + k);
      const uint32_t a = base + blk * 40;
      tm_st<8>(a, r); tm_st<2>(a + 8, r + 8);                                   // QJ WT PD
      tm_st<2>(a + 10, r + 10); tm_st<4>(a + 12, r + 12); tm_st<4>(a + 16, r + 16); tm_st<2>(a + 20, r + 20);   // R ZETA U
      tm_st<2>(a + 22, r + 22); tm_st<4>(a + 24, r + 24);                       // E
      tm_st<4>(a + 28, r + 28); tm_st<8>(a + 32, r + 32);                       // C bn bf
    }
    wait_st();
    __syncthreads();
    for (int blk = 2; blk >= 0; blk--) {
      float r[40];
      const uint32_t a = base + blk * 40;
      tm_ld<4>(a, r); tm_ld<2>(a + 4, r + 4); tm_ld<1>(a + 6, r + 6); tm_ld<1>(a + 7, r + 7); tm_ld<2>(a + 8, r + 8);
      tm_ld<2>(a + 10, r + 10); tm_ld<4>(a + 12, r + 12); tm_ld<2>(a + 16, r + 16); tm_ld<1>(a + 18, r + 18); tm_ld<1>(a + 19, r + 19);
      tm_ld<4>(a + 20, r + 20); tm_ld<8>(a + 24, r + 24); tm_ld<8>(a + 32, r + 32);
      wait_ld();
      for (int k = 0; k < 40; k++) if (r[k] != val(w, lane, blk * 40 + k)) { bad++; atomicMin(first, w * 100000 + lane * 1000 + blk * 40 + k); }
    }
    // overwrite a middle run and re-read its neighbours (no clobbering across run boundaries)
    { float r[3] = {-1.f, -2.f, -3.f}, q[8];
      const uint32_t a = base + 40;
      tm_st<1>(a + 19, r); tm_st<2>(a + 20, r + 1); wait_st();
      tm_ld<8>(a + 16, q); wait_ld();
      for (int k = 0; k < 8; k++) { const float e = (k >= 3 && k <= 5) ? -(float)(k - 2) : val(w, lane, 40 + 16 + k); if (q[k] != e) bad++; } }
  } else {
    float r[12], q[12];
    for (int k = 0; k < 12; k++) r[k] = val(w, lane, 300 + k);
    const uint32_t a = base + 1;   // odd column
    tm_st<4>(a, r); tm_st<8>(a + 4, r + 4); wait_st();
    tm_ld<8>(a, q); tm_ld<4>(a + 8, q + 8); wait_ld();
    for (int k = 0; k < 12; k++) if (q[k] != r[k]) bad++;
  }
  if (bad) atomicAdd(err, bad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512) : "memory");
}

// dependent chain: 12 floats in, a little math, 6 floats out - through TMEM (kind 0) or shared memory (kind 1)
__global__ void __launch_bounds__(448, 1) k_lat(int kind, int active_warps, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t s_base;
  extern __shared__ __align__(16) float dsm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + (((uint32_t)(w & 3) * 32u) << 16) + (uint32_t)(w >> 2) * 128u;
  float* my = dsm + (size_t)threadIdx.x * 28;   // 28-float records: conflict-free for 128-bit accesses of neighbouring lanes
  float r[12];
  for (int k = 0; k < 12; k++) r[k] = 0.001f * (lane + k);
  if (kind == 0) { tm_st<4>(base + 12, r); tm_st<8>(base + 16, r + 4); wait_st(); }
  else { for (int k = 0; k < 12; k++) my[12 + k] = r[k]; }
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (w < active_warps) {
    t0 = clock64();
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
      const int blk = it % 3;   // run-time offsets as in the physics kernel
      if (kind == 0) {
        const uint32_t a = base + blk * 40;
        tm_ld<4>(a + 12, r); tm_ld<8>(a + 16, r + 4); wait_ld();
      } else {
        const float4 v0 = *reinterpret_cast<const float4*>(my + 12), v1 = *reinterpret_cast<const float4*>(my + 16), v2 = *reinterpret_cast<const float4*>(my + 20);
        r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w; r[8] = v2.x; r[9] = v2.y; r[10] = v2.z; r[11] = v2.w;
      }
      float o[6];
      for (int k = 0; k < 6; k++) o[k] = r[k] * 1.0001f + r[k + 6] * 0.5f + acc * 1e-6f;
      acc += o[0] + o[3];
      if (kind == 0) {
        const uint32_t a = base + ((it + 1) % 3) * 40;
        tm_st<4>(a + 12, o); tm_st<2>(a + 16, o + 4); wait_st();
      } else {
        *reinterpret_cast<float4*>(my + 12) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float2*>(my + 16) = make_float2(o[4], o[5]);
      }
    }
    t1 = clock64();
    sink[blockIdx.x * 448 + threadIdx.x] = acc;
  }
  if (lane == 0 && blockIdx.x == 0) cycles[w] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512) : "memory");
}

int main() {
  int *err, *first; long long* cyc; float* sink;
  CK(cudaMallocManaged(&err, 4)); CK(cudaMallocManaged(&first, 4)); CK(cudaMallocManaged(&cyc, 16 * 8)); CK(cudaMalloc(&sink, 148 * 448 * 4));
  *err = 0; *first = 1 << 30;
  k_check<<<148, 448>>>(0, err, first);
  CK(cudaDeviceSynchronize());
  printf("check aligned layout, 148 CTAs x 14 warps: %d mismatches (first code %d)\n", *err, *first);
  CK(cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 448 * 28 * 4));
  const int iters = 2000;
  for (int kind = 0; kind < 2; kind++)
    for (int aw : {1, 4, 7, 14}) {
      k_lat<<<148, 448, 448 * 28 * 4>>>(kind, aw, iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      long long mx = 0; for (int w = 0; w < aw; w++) mx = cyc[w] > mx ? cyc[w] : mx;
      printf("%s chain, %2d warps: %.1f cycles per ld+math+st round trip\n", kind == 0 ? "TMEM" : "smem", aw, (double)mx / iters);
    }
  *err = 0;
  k_check<<<1, 448>>>(1, err, first);
  cudaError_t e = cudaDeviceSynchronize();
  printf("unaligned columns (x4 at an odd column, x8 at column 5): %s, %d mismatches\n", e == cudaSuccess ? "ran" : cudaGetErrorString(e), *err);
  return 0;
}
