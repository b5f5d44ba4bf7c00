// cuda_compat.h - TEST INFRASTRUCTURE: lets the device code of vid2player3d_b200/csrc/{dyn_common,packed*}.cuh compile as host C++.
// A warp is emulated by 32 host threads running the same function ("lanes"); __syncwarp() / __syncthreads() are a barrier,
// shuffles exchange values through a per-warp mailbox between two barriers.  Sequential consistency of the host threads at the
// barrier gives exactly the visibility rules the kernels rely on (shared memory written before a __syncwarp is visible after it).
// Nothing here is used by the product path.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define FULL 0xffffffffu

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }

struct EmuWarp {
  std::atomic<int> count{0};
  std::atomic<int> phase{0};
  double box[32];      // shuffle mailbox (a double holds a float, a double or an int exactly)
  int pred[32];
  void barrier() {
    const int ph = phase.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) == 31) {
      count.store(0, std::memory_order_relaxed);
      phase.store(ph + 1, std::memory_order_release);
    } else {
      while (phase.load(std::memory_order_acquire) == ph) std::this_thread::yield();
    }
  }
};
extern thread_local EmuWarp* emu_warp;
extern thread_local int emu_lane;

static inline void __syncwarp(unsigned = FULL) { emu_warp->barrier(); }
static inline void __syncthreads() { emu_warp->barrier(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
  emu_warp->box[emu_lane] = (double)v;
  emu_warp->barrier();
  const T r = (T)emu_warp->box[src & 31];
  emu_warp->barrier();
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return __shfl_sync(m, v, emu_lane ^ x); }
static inline unsigned __ballot_sync(unsigned, int p) {
  emu_warp->pred[emu_lane] = p;
  emu_warp->barrier();
  unsigned r = 0;
  for (int k = 0; k < 32; k++) r |= (emu_warp->pred[k] ? 1u : 0u) << k;
  emu_warp->barrier();
  return r;
}
static inline int __any_sync(unsigned, int p) {
  emu_warp->pred[emu_lane] = p;
  emu_warp->barrier();
  int r = 0;
  for (int k = 0; k < 32; k++) r |= emu_warp->pred[k];
  emu_warp->barrier();
  return r;
}
