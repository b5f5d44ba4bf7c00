// b200nn.cu - libb200nn.so (C ABI: include/b200nn.h): the policy MLP and the MVAE mixture-of-experts decoder layers as
// hand-written sm_100a GEMMs.  One launch per layer:
//
//     out[:, col0 : col0 + N] = act( sum_e coef[:, e] * (A W_e^T + bias_e) )        bf16 operands, fp32 accumulate
//
// Structure of `linear_kernel` (one 128-row output tile per CTA, 6 warps):
//   warp 0   TMA producer: cp.async.bulk.tensor tiles of A [128 x 64] and W [E x BN x 64] (128-byte swizzle) into a ring of
//            STAGES shared-memory stages, completion on mbarriers (expect_tx)
//   warp 1   owns TMEM (tcgen05.alloc / dealloc) and issues tcgen05.mma.cta_group::1.kind::f16 from ONE thread: per 64-wide
//            k block 4 MMAs of K = 16 per accumulator chunk; tcgen05.commit releases the stage / signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld of the warp's 32 TMEM lanes (= 32 output rows), bias, expert blend, activation, bf16 / fp32
//            stores straight from registers (every thread writes whole 32-byte sectors of its row)
// E = 1: BN = 128 accumulator columns.  E > 1 (mixture of experts, vid2player/motion_vae/model.py:237-252): the tile is 64
// output columns x E experts = E * 64 accumulator columns in TMEM (384 for the reference's 6 experts); the per-row softmax
// coefficients are applied in the epilogue in fp32, so the per-env blended weight matrices of the reference never exist.
// Two CTAs of the E = 1 form fit on an SM (96 KB of stages, 128 TMEM columns each): one tile's epilogue overlaps the next
// tile's main loop without a persistent scheduler.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/b200nn.h"

static thread_local char g_err[512] = "";
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
#define CUDA_OK(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) {                                                                         \
      snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -100;                                                                                   \
    }                                                                                                \
  } while (0)

// ------------------------------------------------------------------------------------------ device helpers (PTX)
namespace {

// mixture-of-experts layers: output tile = MOE_BN columns x E experts.  32 (x 6 experts = 192 accumulator columns, 40 KB per stage,
// two CTAs per SM so that one tile's blend epilogue runs under the other's main loop) measured against 64 (384 columns, 64 KB per
// stage, one CTA per SM) in profiles/r2j_nn.md; -DMOE_BN=64 -DMOE_STAGES=3 rebuilds the wide form.
#ifndef MOE_BN
#define MOE_BN 32
#endif
#ifndef MOE_STAGES
#define MOE_STAGES 2
#endif
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                        // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a broken pipeline traps after ~2 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // surfaces as a launch failure on the host
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major bf16 tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes,
// 8-row swizzle atoms of 1024 bytes (stride byte offset), start address in 16-byte units, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  (bit layout: cute/arch/mma_sm100_desc.hpp UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address            bits [0,14)
  d |= (uint64_t)0 << 16;                    // leading byte offset: unused for swizzled K-major tiles
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                    // version                  bits [46,48)
  d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B bits [61,64)
  return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n  (UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: the warp's 32 lanes x 16 (32) consecutive columns, one row per thread.  The load is asynchronous: several are
// issued back to back and ONE tcgen05.wait::ld covers them (the round trip is ~0.5 us when every load is waited for on its own).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == B200NN_ACT_RELU) return fmaxf(x, 0.f);
  if (act == B200NN_ACT_ELU) return x > 0.f ? x : expm1f(x);   // F.elu, alpha = 1
  return x;
}


#ifndef B200NN_PROBE
#define B200NN_PROBE 0
#endif
#if B200NN_PROBE
__device__ unsigned long long g_probe[16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory"); return t; }
#define PROBE(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_probe[i] = gtime(); } while (0)
__device__ unsigned long long g_cta[4096][8];
#define PROBE_CTA(k) do { const int c_ = blockIdx.y * gridDim.x + blockIdx.x; if (c_ < 4096) g_cta[c_][k] = gtime(); } while (0)
#else
#define PROBE_CTA(k) do { } while (0)
#define PROBE(i) do { } while (0)
#endif

struct LinearParams {
  const float* bias;
  const float* coef;
  void* out;
  int ldo, out_col0, rows, n, n_padded, k_padded, act, out_bf16;
  float out_min, out_max;
};

__device__ __forceinline__ float finish(float x, const LinearParams& p) {
  x = apply_act(x, p.act);
  return p.out_min < p.out_max ? fminf(fmaxf(x, p.out_min), p.out_max) : x;
}

__host__ __device__ constexpr int tmem_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// 16 consecutive outputs of one row -> global memory
__device__ __forceinline__ void store16(const LinearParams& p, int row, int col, const float* y) {
  if (col >= p.n) return;
  if (p.out_bf16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + p.out_col0 + col;
    if (col + 16 <= p.n) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      uint4* o4 = reinterpret_cast<uint4*>(o);   // ldo, out_col0 and col are multiples of 8 elements: 16-byte aligned
      o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      for (int i = 0; i < 16 && col + i < p.n; i++) o[i] = __float2bfloat16_rn(y[i]);
    }
  } else {
    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + p.out_col0 + col;
    if (col + 16 <= p.n && ((uintptr_t)o & 15) == 0) {          // widest store the row's alignment allows
#pragma unroll
      for (int i = 0; i < 4; i++) reinterpret_cast<float4*>(o)[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    } else if (col + 16 <= p.n && ((uintptr_t)o & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) reinterpret_cast<float2*>(o)[i] = make_float2(y[2 * i], y[2 * i + 1]);
    } else {
      for (int i = 0; i < 16 && col + i < p.n; i++) o[i] = y[i];
    }
  }
}

template <int E, int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 2)
linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
              const LinearParams p) {
  constexpr int ACC = E * BN;                          // accumulator columns in TMEM
  constexpr int B_STAGE_BYTES = ACC * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int TMEM_COLS = tmem_cols(ACC);
  static_assert(ACC <= 512 && BN % 16 == 0, "tile does not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle-128B atoms: 1024-byte aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_holder + 4);        // [E][BN] bias of this tile's columns

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M, n0 = blockIdx.y * BN;
  const int nkb = p.k_padded / BLOCK_K;

  if (threadIdx.x == 0) { PROBE(0); PROBE_CTA(0); }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // whole warp: allocate the accumulator columns, publish the TMEM base address through shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) PROBE(1);
  // Programmatic dependent launch: the next layer's CTAs may be scheduled as soon as ours have all started (they take the SM
  // slots our last wave leaves free, run their set-up and park at griddepcontrol.wait), and everything ABOVE this line ran
  // while the previous layer was still finishing.  Below it, every read of data a previous kernel produced (the A operand, the
  // blend coefficients) comes after griddepcontrol.wait = previous grid complete and visible.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {   // ---------------- TMA producer
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(sa, &tmA, kb * BLOCK_K, m0, &full[s]);
        tma_load_3d(sa + A_STAGE_BYTES, &tmW, kb * BLOCK_K, n0, 0, &full[s]);
        if (kb == 0) PROBE(2);
      }
      PROBE(3);
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------- MMA issuer (one thread)
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        if (kb == 0) PROBE(4);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
          const uint64_t adesc = umma_desc_sw128(a_addr + k * UMMA_K * 2);
#pragma unroll
          for (int c0 = 0; c0 < ACC; c0 += 256) {      // accumulator chunks of at most 256 columns (UMMA N <= 256)
            const int nc = ACC - c0 < 256 ? ACC - c0 : 256;
            const uint64_t bdesc = umma_desc_sw128(b_addr + c0 * 128 + k * UMMA_K * 2);   // B row r <-> accumulator column r
            umma_bf16(tmem_base + c0, adesc, bdesc, umma_idesc_bf16(nc), (uint32_t)((kb | k) != 0));
          }
        }
        umma_commit(&empty[s]);      // the stage is free again once these MMAs have read it
      }
      umma_commit(tmem_full);        // accumulators complete
      PROBE(5);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31 = rows m0 + 32 (w % 4) + lane
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    // everything that does not depend on the accumulators happens BEFORE the wait: the tile's bias goes to shared memory, the
    // row's blend coefficients to registers
    for (int i = (int)threadIdx.x - 64; i < ACC; i += 128) sbias[i] = __ldg(p.bias + (size_t)(i / BN) * p.n_padded + n0 + (i % BN));
    float coef[E];
    if (E > 1) asm volatile("griddepcontrol.wait;" ::: "memory");    // the coefficients come from the gate kernel before us
#pragma unroll
    for (int e = 0; e < E; e++) coef[e] = (E == 1) ? 1.f : (row < p.rows ? p.coef[(size_t)row * E + e] : 0.f);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps only
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) PROBE(6);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // bf16 outputs leave through shared memory and ONE TMA store per warp and 64-column half: the warp's 32 rows x 128 bytes are
    // staged in the (now idle) pipeline stage 0 in the 128-byte swizzle the tensor map expects - thread = row, 16-byte chunk j of
    // the row at ((j ^ (row & 7)) << 4): conflict-free per quarter warp - and written to global memory as whole 128-byte lines
    // (the direct form wrote 16-byte pieces of 32 different lines per instruction: ~5 us per tile against ~1.5 us now).
    uint8_t* stage_out = smem + (size_t)q * 4096;           // + half * 16384
    const int rsw = lane & 7;
    if constexpr (E == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        if (n0 + c0 >= p.n) break;                   // padding columns of the last tile: nothing to store
        uint32_t r[64];
        tmem_ld32_issue(trow + c0, r);
        tmem_ld32_issue(trow + c0 + 32, r + 32);
        tmem_ld_wait();
        if (p.out_bf16) {
          uint8_t* dst = stage_out + (c0 >> 6) * 16384 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float a = finish(__uint_as_float(r[8 * j + 2 * i]) + sbias[c0 + 8 * j + 2 * i], p);
              const float b = finish(__uint_as_float(r[8 * j + 2 * i + 1]) + sbias[c0 + 8 * j + 2 * i + 1], p);
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst + ((j ^ rsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; i++) y[i] = finish(__uint_as_float(r[16 * j + i]) + sbias[c0 + 16 * j + i], p);
            if (row < p.rows) store16(p, row, n0 + c0 + 16 * j, y);
          }
        }
      }
      if (p.out_bf16) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m0 + q * 32 < p.rows) {
          for (int c0 = 0; c0 < BN && n0 + c0 < p.n; c0 += 64) tma_store_2d(&tmO, stage_out + (c0 >> 6) * 16384, p.out_col0 + n0 + c0, m0 + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
      }
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= p.n) break;
        uint32_t r[E][16];
#pragma unroll
        for (int e = 0; e < E; e++) tmem_ld16_issue(trow + e * BN + c0, r[e]);     // E loads in flight, one wait
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = 0.f;
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
          for (int i = 0; i < 16; i++) y[i] = fmaf(coef[e], __uint_as_float(r[e][i]) + sbias[e * BN + c0 + i], y[i]);
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = finish(y[i], p);
        if (BN == 64 && p.out_bf16) {
          uint8_t* dst = stage_out + lane * 128;
#pragma unroll
          for (int jj = 0; jj < 2; jj++) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              __nv_bfloat162 h = __floats2bfloat162_rn(y[8 * jj + 2 * i], y[8 * jj + 2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst + ((((c0 >> 3) + jj) ^ rsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else if (row < p.rows) {
          store16(p, row, n0 + c0, y);      // 32-column tiles: 64 contiguous bytes per row from registers
        }
      }
      if (BN == 64 && p.out_bf16) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m0 + q * 32 < p.rows) {
          tma_store_2d(&tmO, stage_out, p.out_col0 + n0, m0 + q * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
      }
    }
    if (threadIdx.x == 64) PROBE(7);
    if (lane == 0) PROBE_CTA(1 + q);
    tcgen05_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) PROBE(8);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if (lane == 0) PROBE_CTA(5);
  }
}

// ------------------------------------------------------------------------------------------ small SIMT kernels
// rows strided by the grid, columns by the block: coalesced, no integer division; up to three destinations of the same leading
// dimension get the same values (the latent block of the three MixedDecoder layer inputs)
__global__ void cast_rows_kernel(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, __nv_bfloat16* __restrict__ dst2,
                                 __nv_bfloat16* __restrict__ dst3, int ld_dst, int rows, int cols, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, float lo, float hi, const uint8_t* __restrict__ row_mask) {
  for (int r = blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += gridDim.x * blockDim.y) {
    if (row_mask && !row_mask[r]) continue;
    const float* s = src + (size_t)r * ld_src;
    const size_t o = (size_t)r * ld_dst;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      float x = s[c];
      if (mean) x = (x - __ldg(mean + c)) * __ldg(rstd + c);
      const __nv_bfloat16 v = __float2bfloat16_rn(fminf(fmaxf(x, lo), hi));
      dst[o + c] = v;
      if (dst2) dst2[o + c] = v;
      if (dst3) dst3[o + c] = v;
    }
  }
}

// one warp per row: logits_e = h . w_e + b_e (k <= 256), softmax over E <= 8
__global__ void gate_softmax_kernel(const __nv_bfloat16* __restrict__ h, int ldh, int k, const float* __restrict__ w, const float* __restrict__ b,
                                    int E, float* __restrict__ coef, int rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; e++) acc[e] = 0.f;
  for (int j = lane; j < k; j += 32) {
    const float x = __bfloat162float(h[(size_t)row * ldh + j]);
#pragma unroll
    for (int e = 0; e < 8; e++)
      if (e < E) acc[e] = fmaf(x, w[e * k + j], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < 8; e++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  if (lane == 0) {
    float m = -1e30f, s = 0.f;
    for (int e = 0; e < E; e++) { acc[e] += b[e]; m = fmaxf(m, acc[e]); }
    for (int e = 0; e < E; e++) { acc[e] = expf(acc[e] - m); s += acc[e]; }
    for (int e = 0; e < E; e++) coef[(size_t)row * E + e] = acc[e] / s;
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

template <int E, int BN, int STAGES> constexpr int smem_bytes() { return STAGES * (A_STAGE_BYTES + E * BN * BLOCK_K * 2) + 1024 + 256 + E * BN * 4; }

}  // namespace

// A/B switch (B200NN_CARVEOUT=1): every kernel of this library asks for the same L1 / shared-memory split (all shared), so that
// neighbouring launches never re-partition the SMs' L1.  Measured on B200 (profiles/r2j_nn.md): no gain for the policy (91.6 vs 91.8 us),
// a loss for the decoder chain (118.9 vs 110.7 us: the small cast / gate kernels like their L1) - off by default.
template <class K> static void prefer_max_shared(K kernel) {
  static const bool on = getenv("B200NN_CARVEOUT") && getenv("B200NN_CARVEOUT")[0] == '1';
  if (on) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

struct b200nn_linear {
  b200nn_linear_desc_t d;
  CUtensorMap tmA, tmW, tmO;
  LinearParams p;
  int device, bn, pdl;
  dim3 grid;
};

template <int E, int BN, int STAGES> static int launch(const b200nn_linear* h, cudaStream_t st) {
  static bool attr_set[16] = {};
  if (!attr_set[h->device & 15]) {
    CUDA_OK(cudaFuncSetAttribute(linear_kernel<E, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<E, BN, STAGES>()));
    prefer_max_shared(linear_kernel<E, BN, STAGES>);
    attr_set[h->device & 15] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = h->grid;
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes<E, BN, STAGES>();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = h->pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, linear_kernel<E, BN, STAGES>, h->tmA, h->tmW, h->tmO, h->p));
  return 0;
}

extern "C" {

int b200nn_abi_version(void) { return B200NN_ABI_VERSION; }
const char* b200nn_last_error(void) { return g_err; }

int b200nn_linear_create(const b200nn_linear_desc_t* d, int32_t device, b200nn_linear_handle* out) {
  if (!d || !out) return fail(-1, "b200nn_linear_create: null argument");
  if (!d->a || !d->w || !d->bias || !d->out) return fail(-1, "b200nn_linear_create: null operand pointer");
  const int E = d->num_experts;
  if (E < 1 || E > 6) return fail(-2, "b200nn_linear_create: num_experts must be 1..6");
  if (E > 1 && !d->coef) return fail(-1, "b200nn_linear_create: mixture layer needs the coefficient tensor");
  const int bn = E == 1 ? 128 : MOE_BN;
  if (d->k_padded < BLOCK_K || d->k_padded % BLOCK_K || d->lda % BLOCK_K || d->ldw % BLOCK_K || d->lda < d->k_padded || d->ldw < d->k_padded)
    return fail(-2, "b200nn_linear_create: K and the leading dimensions must be padded to multiples of 64");
  if (d->n_padded < bn || d->n_padded % bn || d->n < 1 || d->n > d->n_padded) return fail(-2, "b200nn_linear_create: n_padded must be a multiple of the output tile");
  if (d->rows < 1) return fail(-2, "b200nn_linear_create: rows < 1");
  if (d->out_bf16 && (d->ldo % 8 || d->out_col0 % 8)) return fail(-2, "b200nn_linear_create: bf16 output needs ldo and out_col0 multiples of 8");
  if (((uintptr_t)d->a | (uintptr_t)d->w | (uintptr_t)d->out) & 15) return fail(-2, "b200nn_linear_create: operands must be 16-byte aligned");
  encode_tiled_fn enc = get_encode();
  if (!enc) return fail(-3, "b200nn_linear_create: cuTensorMapEncodeTiled not available (no CUDA driver)");
  CUDA_OK(cudaSetDevice(device));
  b200nn_linear* h = new b200nn_linear();
  h->d = *d;
  h->device = device;
  h->bn = bn;
  {
    const char* e = getenv("B200NN_PDL");    // A/B switch: B200NN_PDL=0 launches the layers without programmatic dependent launch
    h->pdl = !(e && e[0] == '0');
  }
  const int rows_padded = (d->rows + BLOCK_M - 1) / BLOCK_M * BLOCK_M;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->k_padded, (cuuint64_t)rows_padded};
    cuuint64_t strides[1] = {(cuuint64_t)d->lda * 2};
    cuuint32_t box[2] = {BLOCK_K, BLOCK_M}, es[2] = {1, 1};
    CUresult r = enc(&h->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->a), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete h; snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(A) failed: %d", (int)r); return -3; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->k_padded, (cuuint64_t)d->n_padded, (cuuint64_t)E};
    cuuint64_t strides[2] = {(cuuint64_t)d->ldw * 2, (cuuint64_t)d->ldw * 2 * (cuuint64_t)d->n_padded};
    cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)bn, (cuuint32_t)E}, es[3] = {1, 1, 1};
    CUresult r = enc(&h->tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete h; snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(W) failed: %d", (int)r); return -3; }
  }
  if (d->out_bf16) {   // output tile store: 64 columns x 32 rows per TMA store, clipped at the real extent of the block that is written
    cuuint64_t dims[2] = {(cuuint64_t)(d->out_col0 + d->n), (cuuint64_t)d->rows};
    cuuint64_t strides[1] = {(cuuint64_t)d->ldo * 2};
    cuuint32_t box[2] = {64, 32}, es[2] = {1, 1};
    CUresult r = enc(&h->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete h; snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(out) failed: %d", (int)r); return -3; }
  } else {
    h->tmO = h->tmA;   // unused by the float-output epilogue
  }
  h->p = LinearParams{d->bias, d->coef, d->out, d->ldo, d->out_col0, d->rows, d->n, d->n_padded, d->k_padded, d->act, d->out_bf16, d->out_min, d->out_max};
  h->grid = dim3(rows_padded / BLOCK_M, (d->n + bn - 1) / bn, 1);
  *out = h;
  return 0;
}

int b200nn_linear_destroy(b200nn_linear_handle h) {
  delete h;
  return 0;
}

int b200nn_linear_run(b200nn_linear_handle h, void* stream) {
  if (!h) return fail(-1, "b200nn_linear_run: null handle");
  cudaStream_t st = (cudaStream_t)stream;
  switch (h->d.num_experts) {
    case 1: return launch<1, 128, 3>(h, st);
    case 2: return launch<2, MOE_BN, 3>(h, st);
    case 3: return launch<3, MOE_BN, 3>(h, st);
    case 4: return launch<4, MOE_BN, 3>(h, st);
    case 5: return launch<5, MOE_BN, MOE_STAGES>(h, st);
    case 6: return launch<6, MOE_BN, MOE_STAGES>(h, st);
  }
  return fail(-2, "b200nn_linear_run: unsupported expert count");
}

#if B200NN_PROBE
int b200nn_probe_read_ctas(unsigned long long* out, int n) { return cudaMemcpyFromSymbol(out, g_cta, sizeof(unsigned long long) * 8 * n) == cudaSuccess ? 0 : -1; }
int b200nn_probe_read(unsigned long long* out16) { return cudaMemcpyFromSymbol(out16, g_probe, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1; }
#endif

static int cast_launch(const float* src, int32_t ld_src, void* dst, void* dst2, void* dst3, int32_t ld_dst, int32_t rows, int32_t cols,
                       const float* mean, const float* rstd, float lo, float hi, void* stream, const uint8_t* row_mask = nullptr) {
  if (!src || !dst || rows < 1 || cols < 1 || cols > ld_src || cols > ld_dst) return fail(-2, "b200nn_cast_rows: bad argument");
  if ((mean == nullptr) != (rstd == nullptr)) return fail(-2, "b200nn_cast_rows: mean and rstd go together");
  const int tx = cols >= 256 ? 256 : (cols >= 128 ? 128 : (cols >= 64 ? 64 : 32));
  const dim3 block(tx, 256 / tx, 1);
  const int want = (rows + (int)block.y - 1) / (int)block.y;
  const int blocks = want < 148 * 16 ? want : 148 * 16;
  static bool once = (prefer_max_shared(cast_rows_kernel), true);
  (void)once;
  cast_rows_kernel<<<blocks, block, 0, (cudaStream_t)stream>>>(src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<__nv_bfloat16*>(dst2),
                                                             reinterpret_cast<__nv_bfloat16*>(dst3), ld_dst, rows, cols, mean, rstd, lo, hi, row_mask);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int b200nn_cast_rows(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows, int32_t cols, const float* mean, const float* rstd,
                     float lo, float hi, void* stream) {
  return cast_launch(src, ld_src, dst, nullptr, nullptr, ld_dst, rows, cols, mean, rstd, lo, hi, stream);
}

int b200nn_cast_rows_masked(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows, int32_t cols, const uint8_t* row_mask,
                            float lo, float hi, void* stream) {
  if (!row_mask) return fail(-2, "b200nn_cast_rows_masked: null mask");
  return cast_launch(src, ld_src, dst, nullptr, nullptr, ld_dst, rows, cols, nullptr, nullptr, lo, hi, stream, row_mask);
}

int b200nn_cast_rows3(const float* src, int32_t ld_src, void* dst, void* dst2, void* dst3, int32_t ld_dst, int32_t rows, int32_t cols, void* stream) {
  return cast_launch(src, ld_src, dst, dst2, dst3, ld_dst, rows, cols, nullptr, nullptr, -3.0e38f, 3.0e38f, stream);
}

int b200nn_gate_softmax(const void* h, int32_t ldh, int32_t k, const float* w, const float* b, int32_t E, float* coef, int32_t rows, void* stream) {
  if (!h || !w || !b || !coef || E < 1 || E > 8 || rows < 1 || k < 1 || k > ldh) return fail(-2, "b200nn_gate_softmax: bad argument");
  static bool once = (prefer_max_shared(gate_softmax_kernel), true);
  (void)once;
  gate_softmax_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(h), ldh, k, w, b, E, coef, rows);
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
