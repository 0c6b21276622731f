// Microbenchmark: latency of cp.async.bulk (1-D TMA) global -> shared on one SM / all SMs of a B200, as one request or split in chunks.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/micro_tma profiles/micro_tma.cu && /tmp/micro_tma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const char* src, int bytes, int chunks, long long* out, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  long long tot = 0;
  uint32_t ph = 0;
  const char* my = src + (size_t)blockIdx.x * bytes;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
      const int cb = bytes / chunks;
      for (int c = 0; c < chunks; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + c * cb)), "l"(my + c * cb), "r"(cb), "r"(s32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)), "r"(ph) : "memory");
    ph ^= 1;
    const long long t1 = clock64();
    if (it > 0) tot += t1 - t0;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = tot / (iters - 1);
}
int main() {
  char* src; long long* out; long long h[148];
  cudaMalloc(&src, 148 * 65536); cudaMemset(src, 1, 148 * 65536); cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int grid : {1, 148})
    for (int bytes : {4096, 32768, 65536})
      for (int chunks : {1, 4, 16}) {
        k<<<grid, 128, 65536>>>(src, bytes, chunks, out, 20);
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  %6d bytes in %2d request(s): %6lld cycles (max over CTAs), %.1f B/clk/SM   [%s]\n", grid, bytes, chunks, mx, (double)bytes / mx, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
