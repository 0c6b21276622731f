// Microbenchmark: issue rate of legacy mma.sync (TF32 m16n8k8, BF16 m16n8k16), packed FFMA2, shared-memory atomics (ATOMS.OR with and
// without a consumed return value) and FLO on one SM of a B200.  nvcc -arch=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KIND, int CHAINS>
__global__ void k_mma(float* out, long long* cyc, int iters) {
  float acc[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  uint32_t b0 = threadIdx.x * 11u, b1 = threadIdx.x * 13u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (KIND == 0) mma_tf32(acc[i], a, b0, b1);
      else mma_bf16(acc[i], a, b0, b1);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2(float* out, long long* cyc, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
  float2 x = make_float2(threadIdx.x * 1e-3f, 1.f), y = make_float2(1.0001f, 0.9999f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(x, y, acc[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>  // 0: atomicOr with used return, 1: atomicOr result unused (RED), 2: plain st (no atomic), 3: atomicAdd used
__global__ void k_atoms(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t bm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) bm[i] = 0;
  __syncthreads();
  uint32_t h = threadIdx.x * 2654435761u + 12345u, seen = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    h = h * 1664525u + 1013904223u;
    const uint32_t w = (h >> 12) & 4095u, bit = 1u << (h & 31);
    if (MODE == 0) seen |= atomicOr(&bm[w], bit) & bit;
    else if (MODE == 1) atomicOr(&bm[w], bit);
    else if (MODE == 2) bm[w] = bit;
    else seen += atomicAdd(&bm[w], 1u);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = seen;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s += bm[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_flo(uint32_t* out, long long* cyc, int iters) {
  uint32_t w = threadIdx.x * 2654435761u | 1u, s = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int b = __ffs((int)w) - 1;
    s += b;
    w = (w & (w - 1u)) | (w << 7) | 0x80000000u;
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  long long h;
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16, 32}) {
    k_mma<0, 8><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("tf32 m16n8k8  8 chains, %2d warps: %.2f cycles per MMA per SM (%.1f MAC/clk/SM)\n", warps, (double)h / (iters * 8.0 * warps), 1024.0 * iters * 8 * warps / h);
    k_mma<0, 1><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("tf32 m16n8k8  1 chain,  %2d warps: %.2f cycles per dependent MMA (latency when 1 warp)\n", warps, (double)h / iters);
    k_mma<1, 8><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("bf16 m16n8k16 8 chains, %2d warps: %.2f cycles per MMA per SM (%.1f MAC/clk/SM)\n", warps, (double)h / (iters * 8.0 * warps), 2048.0 * iters * 8 * warps / h);
    k_ffma2<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ffma2         8 chains, %2d warps: %.1f MAC/clk/SM\n", warps, 64.0 * iters * 8 * warps / h);
    k_atoms<0><<<1, warps * 32>>>((uint32_t*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ATOMS.OR ret  %2d warps: %.2f cycles per lane-op per SM\n", warps, (double)h / (iters * 32.0 * warps));
    k_atoms<1><<<1, warps * 32>>>((uint32_t*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ATOMS.OR nort %2d warps: %.2f cycles per lane-op per SM\n", warps, (double)h / (iters * 32.0 * warps));
    k_atoms<3><<<1, warps * 32>>>((uint32_t*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ATOMS.ADD ret %2d warps: %.2f cycles per lane-op per SM\n", warps, (double)h / (iters * 32.0 * warps));
    k_atoms<2><<<1, warps * 32>>>((uint32_t*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS random    %2d warps: %.2f cycles per lane-op per SM\n", warps, (double)h / (iters * 32.0 * warps));
    k_flo<<<1, warps * 32>>>((uint32_t*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ffs loop      %2d warps: %.2f cycles per iteration per warp-slot\n", warps, (double)h / (iters * 1.0 * warps));
  }
  return 0;
}
