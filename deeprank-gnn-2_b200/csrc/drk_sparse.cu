// Segmented gather-reduce kernels: the aggregation half of every DeepRank2 convolution and
// the per-graph readout.
//
//   drk_spmm          out[i,:] = epi( reduce_{s in seg i} w[s] * src[idx[s],:] )
//                     == x[col] -> scatter_sum(.., row)           ginet.py:45,58  vanilla_gnn.py:30,35
//                     == per-node mean loop of FoutLayer           foutnet.py:56-58 (MEAN_NAN)
//                     == scatter_mean(edge_attr * .., row, out=0)  sgat.py:68-72    (MEAN_CLAMP, w)
//   drk_segment_mean  scatter_mean(x, batch, dim=0)               ginet_nocluster.py:103-104
//
// Layout: one sub-warp of LPR lanes per destination row, each lane owning VEC consecutive
// columns (LPR*VEC >= width for the common widths 16/32/64 -> one 128-bit load per lane per
// gathered row, a full 128 B line per 32-wide row).  Edges of a row are visited in CSR order
// (ascending edge id = the order of the reference's CPU scatter_add_), accumulated sequentially
// in fp32, no atomics -> bit-reproducible.  Consecutive rows go to the same CTA so the gathered
// rows of one graph (~300 x 128 B = 38 KB) stay resident in that SM's L1; the index stream is
// loaded with L1::no_allocate so it does not evict them.
#include <algorithm>

#include "drk_common.cuh"

namespace drk {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = ld_gather_f4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  __device__ __forceinline__ void fence() { asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3])); }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ void load(const float* p) {
    const float2 t = ld_gather_f2(p);
    v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
  __device__ __forceinline__ void fence() { asm volatile("" : "+f"(v[0]), "+f"(v[1])); }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
  __device__ __forceinline__ void fence() { asm volatile("" : "+f"(v[0])); }
};

__device__ __forceinline__ float relu_keep_nan(float x) { return x < 0.f ? 0.f : x; }           // torch.relu(NaN) = NaN
__device__ __forceinline__ float relu_grad_mask(float g, float m) { return m <= 0.f ? 0.f : g; }  // threshold_backward

struct SpmmArgs {
  const int32_t* ptr;
  const int32_t* idx;
  const float* w;
  const float* src;
  const float* addend;
  const float* mask;
  float* out;
  uint32_t ld_src;  // leading dimensions in ELEMENTS; 32-bit: tensors of up to 2^32 elements (16 GB)
  uint32_t ld_addend;
  uint32_t ld_mask;
  uint32_t ld_out;
  int32_t n_out;
  int32_t width;
  int32_t reduce;
  int32_t act;
  int32_t rows_per_block;
};

constexpr int kSpmmThreads = 256;

// LPR lanes per destination row, VEC floats per lane.  All control flow is WARP-UNIFORM: the 32/LPR rows a
// warp works on are walked in lock-step up to the longest of them, shorter rows run predicated.  (Letting
// each sub-warp loop to its own row length makes the sub-warps diverge and the hardware then issues them
// one after the other: 8x slower for LPR = 4.)  Per trip every lane has kInFlight gathers outstanding.
template <int LPR, int VEC, bool HAS_W, bool HAS_IDX>
__global__ void __launch_bounds__(kSpmmThreads, 4) k_spmm(const SpmmArgs a) {
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kRowsPerPass = (kSpmmThreads / 32) * kRowsPerWarp;
  constexpr int kChunks = LPR >= 8 ? 1 : 8 / LPR;  // index chunks (of LPR edges) per trip
  constexpr int kInFlight = kChunks * LPR;         // edges consumed per trip
  constexpr int kBatch = 8;                        // gathers in flight per lane (8 x float4 = 32 registers)
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = lane_id();
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const int group_base = sub * LPR;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n_out);
  const uint32_t ld_src_bytes = a.ld_src * 4u;

  for (int cbase = 0; cbase < a.width; cbase += LPR * VEC) {  // one trip for width <= LPR*VEC
    const int c = cbase + sl * VEC;
    const bool col_ok = c < a.width;  // width % VEC == 0 is guaranteed by the dispatcher
    const char* src_c = reinterpret_cast<const char*>(a.src + c);
    for (int rw = row0 + warp * kRowsPerWarp; rw < row_end; rw += kRowsPerPass) {  // rw is warp-uniform
      const int r = rw + sub;
      const bool row_ok = r < row_end;
      int beg = 0, len = 0;
      if (row_ok) {
        beg = __ldg(a.ptr + r);
        len = __ldg(a.ptr + r + 1) - beg;
      }
      int max_len = len;
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) max_len = max(max_len, __shfl_xor_sync(kFull, max_len, o));
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;

      // software pipeline: the index chunks of the NEXT trip are loaded while this trip's gathers are in flight
      int next_idx[kChunks];
      float next_w[kChunks];
#pragma unroll
      for (int q = 0; q < kChunks; ++q) {
        const int e = q * LPR + sl;
        next_idx[q] = -1;
        next_w[q] = 0.f;
        if (e < len) {
          next_idx[q] = HAS_IDX ? ld_stream_i32(a.idx + beg + e) : beg + e;
          if (HAS_W) next_w[q] = ld_stream_f32(a.w + beg + e);
        }
      }
      for (int off = 0; off < max_len; off += kInFlight) {
        int my_idx[kChunks];
        float my_w[kChunks];
#pragma unroll
        for (int q = 0; q < kChunks; ++q) {
          my_idx[q] = next_idx[q];
          my_w[q] = next_w[q];
          const int e = off + kInFlight + q * LPR + sl;
          next_idx[q] = -1;
          if (e < len) {
            next_idx[q] = HAS_IDX ? ld_stream_i32(a.idx + beg + e) : beg + e;
            if (HAS_W) next_w[q] = ld_stream_f32(a.w + beg + e);
          }
        }
#pragma unroll
        for (int u0 = 0; u0 < kInFlight; u0 += kBatch) {
          Vec<VEC> t[kBatch];
          int srow[kBatch];
          float tw[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            srow[u] = __shfl_sync(kFull, my_idx[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
            if (HAS_W) tw[u] = __shfl_sync(kFull, my_w[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            if (srow[u] >= 0 && col_ok) t[u].load(reinterpret_cast<const float*>(src_c + (uint64_t)(uint32_t)srow[u] * ld_src_bytes));
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            if (srow[u] >= 0 && col_ok) {  // strictly sequential accumulation in CSR order
#pragma unroll
              for (int k = 0; k < VEC; ++k) acc[k] = HAS_W ? fmaf(tw[u], t[u].v[k], acc[k]) : acc[k] + t[u].v[k];
            }
          }
        }
      }
      if (!col_ok || !row_ok) continue;
      if (a.reduce != DRK_REDUCE_SUM) {
        const float deg = (float)len;
        const float den = a.reduce == DRK_REDUCE_MEAN_CLAMP ? fmaxf(deg, 1.f) : deg;  // MEAN_NAN: 0/0 = NaN like torch.mean([])
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = acc[k] / den;
      }
      if (a.addend != nullptr) {
        Vec<VEC> ad;
        ad.load(a.addend + (size_t)r * a.ld_addend + c);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += ad.v[k];
      }
      if (a.act == DRK_ACT_RELU) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = relu_keep_nan(acc[k]);
      }
      if (a.mask != nullptr) {
        Vec<VEC> m;
        m.load(a.mask + (size_t)r * a.ld_mask + c);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = relu_grad_mask(acc[k], m.v[k]);
      }
      Vec<VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) o.v[k] = acc[k];
      o.store(a.out + (size_t)r * a.ld_out + c);
    }
  }
}

template <int LPR, int VEC>
static void launch_spmm(const SpmmArgs& a, int blocks, cudaStream_t stream) {
  const bool has_w = a.w != nullptr, has_idx = a.idx != nullptr;
  if (has_w && has_idx) k_spmm<LPR, VEC, true, true><<<blocks, kSpmmThreads, 0, stream>>>(a);
  else if (has_w) k_spmm<LPR, VEC, true, false><<<blocks, kSpmmThreads, 0, stream>>>(a);
  else if (has_idx) k_spmm<LPR, VEC, false, true><<<blocks, kSpmmThreads, 0, stream>>>(a);
  else k_spmm<LPR, VEC, false, false><<<blocks, kSpmmThreads, 0, stream>>>(a);
}

// ---------------------------------------------------------------- per-graph mean (one CTA per graph)
constexpr int kMeanThreads = 256;

template <int VEC>
__global__ void __launch_bounds__(kMeanThreads) k_segment_mean(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ graph_ptr,
                                                               int32_t width, float* __restrict__ out, int64_t ld_out) {
  extern __shared__ float partial[];  // [row_lanes][width]
  const int g = blockIdx.x;
  const int beg = graph_ptr[g];
  const int end = graph_ptr[g + 1];
  const int cv = width / VEC;              // vector columns
  const int row_lanes = kMeanThreads / cv;  // >= 1 (dispatcher guarantees cv <= kMeanThreads)
  const int cl = threadIdx.x % cv;
  const int rl = threadIdx.x / cv;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  if (rl < row_lanes) {
    for (int i = beg + rl; i < end; i += row_lanes) {
      Vec<VEC> t;
      t.load(x + (int64_t)i * ldx + cl * VEC);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] += t.v[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) partial[rl * width + cl * VEC + k] = acc[k];
  }
  __syncthreads();
  if (rl == 0) {
    const float den = fmaxf((float)(end - beg), 1.f);  // scatter_mean: count clamped to >= 1
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float s = 0.f;
      for (int p = 0; p < row_lanes; ++p) s += partial[p * width + cl * VEC + k];  // fixed order
      out[(int64_t)g * ld_out + cl * VEC + k] = s / den;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) k_segment_mean_bwd(const float* __restrict__ dg, uint32_t ld_dg, const int32_t* __restrict__ graph_ptr,
                                                          const int32_t* __restrict__ batch32, const float* __restrict__ mask,
                                                          uint32_t ld_mask, int32_t num_nodes, int32_t width, float* __restrict__ dx,
                                                          uint32_t ld_dx) {
  // thread -> (node i, vector column cv); cols = width / VEC vector columns per node (32-bit index math only)
  const uint32_t cols = (uint32_t)width / VEC;
  const uint32_t total = (uint32_t)num_nodes * cols;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t i = t / cols;
    const uint32_t c = (t - i * cols) * VEC;
    const int b = __ldg(batch32 + i);
    const float den = fmaxf((float)(__ldg(graph_ptr + b + 1) - __ldg(graph_ptr + b)), 1.f);  // true division, like scatter_mean
    Vec<VEC> g;
    g.load(dg + (size_t)b * ld_dg + c);
    float o[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) o[k] = g.v[k] / den;
    if (mask != nullptr) {
      Vec<VEC> m;
      m.load(mask + (size_t)i * ld_mask + c);
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = relu_grad_mask(o[k], m.v[k]);
    }
    Vec<VEC> ov;
#pragma unroll
    for (int k = 0; k < VEC; ++k) ov.v[k] = o[k];
    ov.store(dx + (size_t)i * ld_dx + c);
  }
}

static int pick_vec(int32_t width, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds) {
  int vec = 4;
  if (width % 4 != 0) vec = width % 2 == 0 ? 2 : 1;
  for (int64_t ld : lds) {
    if (vec == 4 && ld % 4 != 0) vec = ld % 2 == 0 ? 2 : 1;
    if (vec == 2 && ld % 2 != 0) vec = 1;
  }
  for (const void* p : ptrs) {
    if (p == nullptr) continue;
    if (vec == 4 && !aligned16(p)) vec = aligned8(p) ? 2 : 1;
    if (vec == 2 && !aligned8(p)) vec = 1;
  }
  return vec;
}

}  // namespace drk

extern "C" {

int drk_spmm(const int32_t* ptr, const int32_t* idx, const float* w, const float* src, int64_t ld_src, const float* addend,
             int64_t ld_addend, const float* mask, int64_t ld_mask, float* out, int64_t ld_out, int32_t n_out, int32_t width,
             int32_t reduce, int32_t act, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n_out >= 0 && width >= 0, DRK_EINVAL, "spmm: negative size");
  if (n_out == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(ptr && src && out, DRK_EINVAL, "spmm: null pointer");
  DRK_REQUIRE(reduce >= DRK_REDUCE_SUM && reduce <= DRK_REDUCE_MEAN_NAN, DRK_EINVAL, "spmm: unknown reduce %d", reduce);
  DRK_REQUIRE(act == DRK_ACT_NONE || act == DRK_ACT_RELU, DRK_EINVAL, "spmm: unknown activation %d", act);
  DRK_REQUIRE(ld_src >= 0 && ld_src < (int64_t)1 << 30 && ld_out >= 0 && ld_out < (int64_t)1 << 30 && ld_addend < (int64_t)1 << 30 &&
                  ld_mask < (int64_t)1 << 30,
              DRK_EUNSUPPORTED, "spmm: leading dimension out of range");
  SpmmArgs a{ptr, idx, w, src, addend, mask, out, (uint32_t)ld_src, (uint32_t)ld_addend, (uint32_t)ld_mask, (uint32_t)ld_out,
             n_out, width, reduce, act, 0};
  const int vec = pick_vec(width, {src, addend, mask, out}, {ld_src, addend ? ld_addend : 4, mask ? ld_mask : 4, ld_out});
  const int vcols = width / vec;
  int lpr = 4;
  while (lpr < 32 && lpr < vcols) lpr <<= 1;
  const int rows_per_pass = (kSpmmThreads / 32) * (32 / lpr);
  // contiguous rows per CTA: aim for ~8 CTAs per SM, at least one pass, at most 8 passes
  int passes = (int)ceil_div<int64_t>(n_out, (int64_t)kNumSM * 8 * rows_per_pass);
  passes = std::max(1, std::min(passes, 8));
  a.rows_per_block = rows_per_pass * passes;
  const int blocks = ceil_div(n_out, a.rows_per_block);
  cudaStream_t st = as_stream(stream);
#define DRK_SPMM_CASE(L, V) \
  if (lpr == L && vec == V) launch_spmm<L, V>(a, blocks, st)
  DRK_SPMM_CASE(4, 4); else DRK_SPMM_CASE(8, 4); else DRK_SPMM_CASE(16, 4); else DRK_SPMM_CASE(32, 4);
  else DRK_SPMM_CASE(4, 2); else DRK_SPMM_CASE(8, 2); else DRK_SPMM_CASE(16, 2); else DRK_SPMM_CASE(32, 2);
  else DRK_SPMM_CASE(4, 1); else DRK_SPMM_CASE(8, 1); else DRK_SPMM_CASE(16, 1); else DRK_SPMM_CASE(32, 1);
#undef DRK_SPMM_CASE
  return finish_launch("spmm");
}

int drk_segment_mean(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int32_t width, float* out,
                     int64_t ld_out, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_graphs >= 0 && width >= 0, DRK_EINVAL, "segment mean: negative size");
  if (num_graphs == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(x && graph_ptr && out, DRK_EINVAL, "segment mean: null pointer");
  int vec = pick_vec(width, {x}, {ldx});
  while (width / vec > kMeanThreads && vec > 1) vec >>= 1;
  DRK_REQUIRE(width / vec <= kMeanThreads, DRK_EUNSUPPORTED, "segment mean: width %d too large", width);
  const int row_lanes = kMeanThreads / (width / vec);
  const size_t smem = (size_t)row_lanes * width * sizeof(float);
  cudaStream_t st = as_stream(stream);
  if (vec == 4) k_segment_mean<4><<<num_graphs, kMeanThreads, smem, st>>>(x, ldx, graph_ptr, width, out, ld_out);
  else if (vec == 2) k_segment_mean<2><<<num_graphs, kMeanThreads, smem, st>>>(x, ldx, graph_ptr, width, out, ld_out);
  else k_segment_mean<1><<<num_graphs, kMeanThreads, smem, st>>>(x, ldx, graph_ptr, width, out, ld_out);
  return finish_launch("segment mean");
}

int drk_segment_mean_bwd(const float* dg, int64_t ld_dg, const int32_t* graph_ptr, const int32_t* batch32, const float* mask,
                         int64_t ld_mask, int32_t num_nodes, int32_t width, float* dx, int64_t ld_dx, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && width >= 0, DRK_EINVAL, "segment mean bwd: negative size");
  if (num_nodes == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(dg && graph_ptr && batch32 && dx, DRK_EINVAL, "segment mean bwd: null pointer");
  DRK_REQUIRE((int64_t)num_nodes * width < (int64_t)1 << 31, DRK_EUNSUPPORTED, "segment mean bwd: more than 2^31 elements");
  const int vec = pick_vec(width, {dg, mask, dx}, {ld_dg, mask ? ld_mask : 4, ld_dx});
  const int64_t work = (int64_t)num_nodes * (width / vec);
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(work, 256), (int64_t)kNumSM * 32);
  cudaStream_t st = as_stream(stream);
  if (vec == 4) k_segment_mean_bwd<4><<<blocks, 256, 0, st>>>(dg, (uint32_t)ld_dg, graph_ptr, batch32, mask, (uint32_t)ld_mask, num_nodes, width, dx, (uint32_t)ld_dx);
  else if (vec == 2) k_segment_mean_bwd<2><<<blocks, 256, 0, st>>>(dg, (uint32_t)ld_dg, graph_ptr, batch32, mask, (uint32_t)ld_mask, num_nodes, width, dx, (uint32_t)ld_dx);
  else k_segment_mean_bwd<1><<<blocks, 256, 0, st>>>(dg, (uint32_t)ld_dg, graph_ptr, batch32, mask, (uint32_t)ld_mask, num_nodes, width, dx, (uint32_t)ld_dx);
  return finish_launch("segment mean bwd");
}

}  // extern "C"
