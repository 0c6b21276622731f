// Shared host/device helpers for the drk_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "drk_b200.h"

namespace drk {

constexpr int kWarp = 32;
constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- host side: thread-local error text + launch accounting
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int finish_launch(const char* what, int n_launches = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DRK_ECUDA;
  }
  g_launches.fetch_add(n_launches, std::memory_order_relaxed);
  return DRK_OK;
}

#define DRK_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::drk::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

// ---- device side
#ifdef __CUDACC__

// streaming (read-once) loads: keep them out of L1 so the gathered rows stay resident there
__device__ __forceinline__ int ld_stream_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ld_stream_i64(const int64_t* p) {
  long long v;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// gathered rows: read-only path, allocate in L1 (neighbouring destinations re-use them)
__device__ __forceinline__ float4 ld_gather_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ld_gather_f2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- asynchronous global -> shared copies (LDGSTS): fire-and-forget, so one thread keeps dozens of
// loads in flight without holding registers.  BYTES in {4, 8, 16}; if !pred the destination is zero-filled
// (src-size 0) and the source is not dereferenced.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src, bool pred) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int src_size = pred ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(gmem_src), "n"(BYTES), "r"(src_size) : "memory");
}
// 16-byte copy that bypasses L1 (data another phase of the same CTA wrote to global memory)
__device__ __forceinline__ void cp_async_cg16(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk: SASS UBLKCP) tracked by an mbarrier in shared memory.  One thread issues a copy of
// any multiple of 16 bytes (16-byte aligned on both sides); the copy engine moves it without occupying the LSU or any register, and
// signals completion by transaction bytes on the barrier.  Waiters spin on the barrier's phase parity (acquire: the data is visible).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DRK_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DRK_MBAR_DONE;\n"
      "bra DRK_MBAR_WAIT;\n"
      "DRK_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// generic-proxy accesses to shared memory made so far are ordered before async-proxy (bulk copy) accesses issued afterwards
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

// ---- tensor cores with fp32-level accuracy: mma.sync m16n8k8 TF32 with error compensation ("3xTF32").
// An fp32 operand is split as v = hi + lo with hi = tf32(v), lo = tf32(v - hi); a product is accumulated (fp32) as
// lo_a*hi_b + hi_a*lo_b + hi_a*hi_b: the dropped lo*lo term is ~2^-22 relative, far inside the 1e-5 parity bar, while a plain
// TF32 product (2^-11) would not be.
// Fragment layout (PTX ISA, g = lane >> 2, t = lane & 3):  A 16x8 row-major: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B 8x8: b0 (k = t, n = g) b1 (k = t+4, n = g);  C 16x8: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// fp32 -> tf32 (10-bit mantissa), round to nearest with ties away from zero: exactly what `cvt.rna.tf32.f32` computes for finite values,
// done with one integer add and one mask.  The conversion instruction runs on the XU pipe (16 lanes/clk/SM); the 3xTF32 split needs
// two conversions per operand element and was what bounded every tensor-core phase of the step kernel (ncu: math-pipe throttle on the
// cvt lines; ~7 k of the 12 k cycles of the dW1 phase), while IADD / LOP3 issue at the full rate.
__device__ __forceinline__ uint32_t to_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(v);
  lo = to_tf32(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a * b for fp32 a (already split) and fp32 b (already split)
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint2 bhi, uint2 blo) {
  mma_tf32(c, alo, bhi.x, bhi.y);
  mma_tf32(c, ahi, blo.x, blo.y);
  mma_tf32(c, ahi, bhi.x, bhi.y);
}

#endif  // __CUDACC__

}  // namespace drk
