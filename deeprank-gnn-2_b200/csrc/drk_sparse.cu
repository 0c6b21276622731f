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

#include <cooperative_groups.h>

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

// ---------------------------------------------------------------- per-graph tiled variant (block-diagonal adjacency)
// For collated batches whose graphs are too big for the whole-step kernel but whose source rows still fit one SM: atom-level graphs
// of ~3 k nodes.  A work unit is (graph g, group of 16 columns, slice of the graph's destination rows): the CTA
// copies the [n_g x 16] column slice of `src` into shared memory ONCE (cp.async, 64-byte rows, <= 3584 rows = 224 KB) and every
// gathered row is then a shared-memory read instead of an L2 sector request; the destination rows of a graph are split over several
// CTAs so that the grid covers all 148 SMs (the tile is re-read from L2 by each of them, from HBM once).  Same lane layout, visiting
// order (CSR order, sequential fp32 accumulation) and epilogues as k_spmm: bit-identical results.  A source row outside the graph's
// range (not a block-diagonal adjacency) is fetched from global memory instead -- correct, just not fast.
constexpr int kTileThreads = 1024;  // one CTA per SM (the tile takes the shared memory): 32 warps to hide the latency of the index stream
constexpr int kTileCols = 16;
constexpr int kTileMaxRows = 3584;  // 3584 x 64 B = 224 KB

struct SpmmTileArgs {
  SpmmArgs base;
  const int32_t* graph_ptr;
  int32_t num_graphs, col_groups, row_splits;
};

template <bool HAS_W>
__global__ void __launch_bounds__(kTileThreads, 1) k_spmm_tiled(const SpmmTileArgs t) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  constexpr int LPR = 4, VEC = 4, kRowsPerWarp = 8, kWarps = kTileThreads / 32, kBatch = 8;
  constexpr unsigned kFull = 0xffffffffu;
  const SpmmArgs& a = t.base;
  const int lane = lane_id(), sub = lane / LPR, sl = lane % LPR, group_base = sub * LPR, warp = threadIdx.x >> 5;
  int unit = blockIdx.x;
  const int split = unit % t.row_splits;
  unit /= t.row_splits;
  const int cg = unit % t.col_groups;
  const int g = unit / t.col_groups;
  const int n0 = __ldg(t.graph_ptr + g), n = __ldg(t.graph_ptr + g + 1) - n0;
  const int c = cg * kTileCols + sl * VEC;
  // ---- the column slice of the graph's source rows -> shared memory (row r at byte 64 r), by the copy engine (cp.async.bulk): one
  // request when the slice is contiguous in global memory (width 16, dense rows), else one 64-byte request per row, issued by all the
  // threads of the CTA in parallel.  (Thread-issued cp.async moves only ~10 B/clk/SM here: 20 k cycles for a 3 k-row tile.)
  __shared__ __align__(8) unsigned long long tile_bar;
  if (threadIdx.x == 0) {
    mbar_init(&tile_bar, 1);
    mbar_init_fence();
  }
  __syncthreads();
  {
    const char* src_cg = reinterpret_cast<const char*>(a.src + cg * kTileCols);
    const uint64_t ld_bytes = (uint64_t)a.ld_src * 4u;
    if (threadIdx.x == 0) mbar_expect_tx(&tile_bar, (uint32_t)n * 64u);
    __syncthreads();  // the expectation is registered before any copy can complete
    if (ld_bytes == 64) {
      if (threadIdx.x == 0 && n > 0) bulk_copy_g2s(tile_smem, src_cg + (uint64_t)(uint32_t)n0 * 64u, (uint32_t)n * 64u, &tile_bar);
    } else {
      for (int r = threadIdx.x; r < n; r += kTileThreads) bulk_copy_g2s(tile_smem + (size_t)r * 64, src_cg + (uint64_t)(uint32_t)(n0 + r) * ld_bytes, 64u, &tile_bar);
    }
  }
  const int per = (((n + t.row_splits - 1) / t.row_splits) + kRowsPerWarp - 1) / kRowsPerWarp * kRowsPerWarp;
  const int row0 = n0 + split * per, row_end = min(row0 + per, n0 + n);
  const char* src_c = reinterpret_cast<const char*>(a.src + c);
  const uint32_t ld_src_bytes = a.ld_src * 4u;
  bool tile_ready = false;
  for (int rw = row0 + warp * kRowsPerWarp; rw < row_end; rw += kWarps * kRowsPerWarp) {
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
    float acc[VEC] = {0.f, 0.f, 0.f, 0.f};
    // 8 edges per trip: every lane of the row's group fetches two index words, the group shares them by shuffle
    int next_idx[2];
    float next_w[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = q * LPR + sl;
      next_idx[q] = -1;
      next_w[q] = 0.f;
      if (e < len) {
        next_idx[q] = ld_stream_i32(a.idx + beg + e);
        if (HAS_W) next_w[q] = ld_stream_f32(a.w + beg + e);
      }
    }
    if (!tile_ready) {  // the first rows' offsets and indices are on their way while the tile lands
      mbar_wait(&tile_bar, 0);
      tile_ready = true;
    }
    for (int off = 0; off < max_len; off += kBatch) {
      int my_idx[2];
      float my_w[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        my_idx[q] = next_idx[q];
        my_w[q] = next_w[q];
        const int e = off + kBatch + q * LPR + sl;
        next_idx[q] = -1;
        if (e < len) {
          next_idx[q] = ld_stream_i32(a.idx + beg + e);
          if (HAS_W) next_w[q] = ld_stream_f32(a.w + beg + e);
        }
      }
      float4 v[kBatch];
      int srow[kBatch];
      float tw[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        srow[u] = __shfl_sync(kFull, my_idx[u / LPR], group_base + (u % LPR));
        if (HAS_W) tw[u] = __shfl_sync(kFull, my_w[u / LPR], group_base + (u % LPR));
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (srow[u] >= 0) {
          const unsigned local = (unsigned)(srow[u] - n0);
          if (local < (unsigned)n) v[u] = *reinterpret_cast<const float4*>(tile_smem + (size_t)local * 64 + sl * 16);
          else v[u] = ld_gather_f4(reinterpret_cast<const float*>(src_c + (uint64_t)(uint32_t)srow[u] * ld_src_bytes));
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (srow[u] >= 0) {  // strictly sequential accumulation in CSR order
          if (HAS_W) {
            acc[0] = fmaf(tw[u], v[u].x, acc[0]); acc[1] = fmaf(tw[u], v[u].y, acc[1]); acc[2] = fmaf(tw[u], v[u].z, acc[2]); acc[3] = fmaf(tw[u], v[u].w, acc[3]);
          } else {
            acc[0] += v[u].x; acc[1] += v[u].y; acc[2] += v[u].z; acc[3] += v[u].w;
          }
        }
      }
    }
    if (!row_ok) continue;
    if (a.reduce != DRK_REDUCE_SUM) {
      const float deg = (float)len;
      const float den = a.reduce == DRK_REDUCE_MEAN_CLAMP ? fmaxf(deg, 1.f) : deg;
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
    *reinterpret_cast<float4*>(a.out + (size_t)r * a.ld_out + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------- per-graph mean (one CTA per graph)
constexpr int kMeanThreads = 256;

// A thread-block CLUSTER per graph (1 CTA for residue-level graphs, 8 for atom-level ones): every CTA sums a contiguous slice of the
// graph's rows, the CTA sums meet in rank 0's hands through distributed shared memory and are added in rank order -- no workspace, no
// second launch, no atomics, a fixed association (bit-reproducible).  A thread keeps four independent partial sums over its rows so
// that four loads are in flight (one CTA of 16 row lanes walking 3 k rows one dependent load at a time took 61 us on the C3 batch).
template <int VEC>
__global__ void __launch_bounds__(kMeanThreads) k_segment_mean(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ graph_ptr,
                                                               int32_t width, float* __restrict__ out, int64_t ld_out) {
  extern __shared__ float partial[];  // [row_lanes][width], then the CTA's sum [width]
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int g = blockIdx.x / csize;
  const int gbeg = graph_ptr[g], gend = graph_ptr[g + 1];
  const int per = (gend - gbeg + csize - 1) / csize;
  const int beg = min(gend, gbeg + crank * per), end = min(gend, beg + per);
  const int cv = width / VEC;              // vector columns
  const int row_lanes = kMeanThreads / cv;  // >= 1 (dispatcher guarantees cv <= kMeanThreads)
  const int cl = threadIdx.x % cv;
  const int rl = threadIdx.x / cv;
  float* cta_sum = partial + row_lanes * width;
  if (rl < row_lanes) {
    float acc[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[u][k] = 0.f;
    int i = beg + rl;
    for (; i + 3 * row_lanes < end; i += 4 * row_lanes) {
      Vec<VEC> t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u].load(x + (int64_t)(i + u * row_lanes) * ldx + cl * VEC);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[u][k] += t[u].v[k];
    }
    for (int u = 0; i < end; i += row_lanes, ++u) {
      Vec<VEC> t;
      t.load(x + (int64_t)i * ldx + cl * VEC);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[u & 3][k] += t.v[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) partial[rl * width + cl * VEC + k] = (acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]);
  }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float s = 0.f;
      for (int p = 0; p < row_lanes; ++p) s += partial[p * width + cl * VEC + k];  // fixed order
      cta_sum[cl * VEC + k] = s;
    }
  }
  cluster.sync();
  if (crank == 0 && rl == 0) {
    const float den = fmaxf((float)(gend - gbeg), 1.f);  // scatter_mean: count clamped to >= 1
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float s = 0.f;
      for (int r = 0; r < csize; ++r) s += *cluster.map_shared_rank(cta_sum + cl * VEC + k, r);  // rank order
      out[(int64_t)g * ld_out + cl * VEC + k] = s / den;
    }
  }
  cluster.sync();  // the other CTAs' shared memory stays alive until rank 0 has read it
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

int drk_spmm_tiled_supported(int32_t max_graph_nodes, int32_t width) {
  return (max_graph_nodes >= 0 && max_graph_nodes <= drk::kTileMaxRows && width > 0 && width % drk::kTileCols == 0) ? 1 : 0;
}

int drk_spmm_tiled(const int32_t* ptr, const int32_t* idx, const float* w, const float* src, int64_t ld_src, const float* addend, int64_t ld_addend,
                   const float* mask, int64_t ld_mask, float* out, int64_t ld_out, const int32_t* graph_ptr, int32_t num_graphs,
                   int32_t max_graph_nodes, int32_t width, int32_t reduce, int32_t act, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_graphs >= 0 && width >= 0 && max_graph_nodes >= 0, DRK_EINVAL, "spmm tiled: negative size");
  if (num_graphs == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(drk_spmm_tiled_supported(max_graph_nodes, width), DRK_EUNSUPPORTED, "spmm tiled: graphs of %d nodes / width %d do not fit (<= %d rows, width %% %d == 0)",
              max_graph_nodes, width, kTileMaxRows, kTileCols);
  DRK_REQUIRE(ptr && idx && src && out && graph_ptr, DRK_EINVAL, "spmm tiled: null pointer");
  DRK_REQUIRE(reduce >= DRK_REDUCE_SUM && reduce <= DRK_REDUCE_MEAN_NAN, DRK_EINVAL, "spmm tiled: unknown reduce %d", reduce);
  DRK_REQUIRE(act == DRK_ACT_NONE || act == DRK_ACT_RELU, DRK_EINVAL, "spmm tiled: unknown activation %d", act);
  DRK_REQUIRE(ld_src >= 0 && ld_src < (int64_t)1 << 30 && ld_out >= 0 && ld_out < (int64_t)1 << 30 && ld_addend < (int64_t)1 << 30 && ld_mask < (int64_t)1 << 30,
              DRK_EUNSUPPORTED, "spmm tiled: leading dimension out of range");
  DRK_REQUIRE(aligned16(src) && aligned16(out) && ld_src % 4 == 0 && ld_out % 4 == 0 && (!addend || (aligned16(addend) && ld_addend % 4 == 0)) &&
                  (!mask || (aligned16(mask) && ld_mask % 4 == 0)),
              DRK_EUNSUPPORTED, "spmm tiled: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  SpmmTileArgs t{};
  t.base = SpmmArgs{ptr, idx, w, src, addend, mask, out, (uint32_t)ld_src, (uint32_t)ld_addend, (uint32_t)ld_mask, (uint32_t)ld_out, 0, width, reduce, act, 0};
  t.graph_ptr = graph_ptr;
  t.num_graphs = num_graphs;
  t.col_groups = width / kTileCols;
  // destination rows of a graph over enough CTAs for ~2 waves of the 148 SMs (one CTA per SM: the tile takes the whole shared memory)
  const int units = num_graphs * t.col_groups;
  t.row_splits = std::max(1, std::min(8, ceil_div(2 * kNumSM, units)));
  const size_t smem = (size_t)std::max(max_graph_nodes, 1) * 64;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaFuncSetAttribute(w ? k_spmm_tiled<true> : k_spmm_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "spmm tiled: smem opt-in: %s", cudaGetErrorString(e));
  if (w) k_spmm_tiled<true><<<units * t.row_splits, kTileThreads, smem, st>>>(t);
  else k_spmm_tiled<false><<<units * t.row_splits, kTileThreads, smem, st>>>(t);
  return finish_launch("spmm tiled");
}

int drk_segment_mean_rows(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int64_t num_rows_hint, int32_t width, float* out,
                          int64_t ld_out, void* stream);

int drk_segment_mean(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int32_t width, float* out,
                     int64_t ld_out, void* stream) {
  return drk_segment_mean_rows(x, ldx, graph_ptr, num_graphs, 0, width, out, ld_out, stream);
}

int drk_segment_mean_rows(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int64_t num_rows_hint, int32_t width, float* out,
                          int64_t ld_out, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_graphs >= 0 && width >= 0, DRK_EINVAL, "segment mean: negative size");
  if (num_graphs == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(x && graph_ptr && out, DRK_EINVAL, "segment mean: null pointer");
  int vec = pick_vec(width, {x}, {ldx});
  while (width / vec > kMeanThreads && vec > 1) vec >>= 1;
  DRK_REQUIRE(width / vec <= kMeanThreads, DRK_EUNSUPPORTED, "segment mean: width %d too large", width);
  const int row_lanes = kMeanThreads / (width / vec);
  const size_t smem = (size_t)(row_lanes + 1) * width * sizeof(float);
  cudaStream_t st = as_stream(stream);
  // cluster size from the mean graph size, which the caller's offsets imply only on the device: use the row count hint num_rows_hint
  int csize = 1;
  if (num_rows_hint > 0) {
    const int64_t mean_rows = num_rows_hint / num_graphs;
    csize = mean_rows >= 2048 ? 8 : mean_rows >= 1024 ? 4 : mean_rows >= 512 ? 2 : 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)num_graphs * csize);
  cfg.blockDim = dim3(kMeanThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (vec == 4) e = cudaLaunchKernelEx(&cfg, k_segment_mean<4>, x, ldx, graph_ptr, width, out, ld_out);
  else if (vec == 2) e = cudaLaunchKernelEx(&cfg, k_segment_mean<2>, x, ldx, graph_ptr, width, out, ld_out);
  else e = cudaLaunchKernelEx(&cfg, k_segment_mean<1>, x, ldx, graph_ptr, width, out, ld_out);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "segment mean: launch: %s", cudaGetErrorString(e));
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

// =====================================================================================================
// Per-edge ReLU messages of VanillaConvolutionalLayer (vanilla_gnn.py:26-35), message size 32:
//   m_e = relu( U[row_e] + V[col_e] + C attr_e ),   S[i] = sum_{e in row i} m_e
// with U = x Wa^T + b (destination half of _edge_mlp), V = x Wb^T (source half), C = the edge-feature
// columns of _edge_mlp.  The reference materialises cat[x_i, x_j, e] as an [E, 2F+Fe] matrix and runs a
// dense GEMM over it (7.4 GFLOP per layer at C2); here the two node halves are projected once per NODE
// and only the 32-wide V rows are gathered per edge.
//   forward also emits, per edge (by ORIGINAL edge id), the 32-bit ReLU mask, and per node cnt[i,c] =
//   number of active edges of channel c (both consumed by the backward kernels).
// =====================================================================================================
namespace drk {

constexpr int kMsg = 32;       // message size fixed by the reference (vanilla_gnn.py:20)
constexpr int kMaxEdgeFeat = 8;

struct EdgeMsgArgs {
  const int32_t* ptr;    // CSR by destination
  const int32_t* idx;    // source node of each CSR slot
  const int32_t* perm;   // original edge id of each CSR slot
  const float* uv;       // [N, 64]: U | V
  const float* attr;     // [E, fe] in ORIGINAL edge order
  const float* cmat;     // [32, ldc]: edge-feature block of the edge-MLP weight (row-major, row stride ldc)
  float* s;              // [N, 32]
  float* cnt;            // [N, 32] (may be NULL)
  uint32_t* mask;        // [E] by original edge id (may be NULL)
  uint32_t ld_uv, ld_attr, ldc, ld_s;
  int32_t n, fe, rows_per_block;
};

// 8 lanes per destination row (4 columns each), 4 rows per warp -- same walk as k_spmm<8,4>.  FE = number of edge features as a
// compile-time constant (1..8; 0 = none): the kernel is INSTRUCTION bound (ncu: 60 M warp instructions for 1.57 M edges with the
// generic 8-feature loops, 124 registers -> 2 CTAs per SM), so the per-edge feature loops and the weight registers must not exist
// for features the model does not have.
template <int FE>
__global__ void __launch_bounds__(256) k_edge_msg_fwd(const EdgeMsgArgs a) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = lane_id();
  const int sub = lane >> 3, sl = lane & 7;
  const int group_base = sub * 8;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n);
  const int c = sl * 4;
  constexpr int kF = FE > 0 ? FE : 1;
  float cw[4][kF];  // this lane's 4 rows of C
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < kF; ++k) cw[q][k] = k < FE ? __ldg(a.cmat + (size_t)(c + q) * a.ldc + k) : 0.f;

  for (int rw = row0 + warp * 4; rw < row_end; rw += 32) {
    const int r = rw + sub;
    const bool row_ok = r < row_end;
    int beg = 0, len = 0;
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok) {
      beg = __ldg(a.ptr + r);
      len = __ldg(a.ptr + r + 1) - beg;
      u = ld_gather_f4(a.uv + (size_t)r * a.ld_uv + c);
    }
    int max_len = len;
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 16));
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 8));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float cn[4] = {0.f, 0.f, 0.f, 0.f};
    for (int off = 0; off < max_len; off += 8) {
      // lane sl of the group fetches source, edge id and attributes of the chunk's sl-th edge: 8 edges in flight per group, the
      // dependent loads (edge id -> attributes) happen once per chunk instead of once per edge
      int my_src = -1, my_eid = 0;
      float my_attr[kF];
#pragma unroll
      for (int k = 0; k < kF; ++k) my_attr[k] = 0.f;
      if (off + sl < len) {
        my_src = ld_stream_i32(a.idx + beg + off + sl);
        my_eid = a.perm != nullptr ? ld_stream_i32(a.perm + beg + off + sl) : beg + off + sl;  // perm == NULL: attributes and masks live in slot order
#pragma unroll
        for (int k = 0; k < FE; ++k) my_attr[k] = __ldg(a.attr + (size_t)my_eid * a.ld_attr + k);
      }
      float4 vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // all gathers of the chunk before the first use
        const int srow = __shfl_sync(kFull, my_src, group_base + j);
        vv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (srow >= 0) vv[j] = ld_gather_f4(a.uv + (size_t)srow * a.ld_uv + kMsg + c);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int srow = __shfl_sync(kFull, my_src, group_base + j);
        const int eid = __shfl_sync(kFull, my_eid, group_base + j);
        const bool on_edge = srow >= 0;  // uniform inside the 8-lane group; everything below is predicated, not branched
        const float4 v = vv[j];
        float m[4] = {u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w};
#pragma unroll
        for (int k = 0; k < FE; ++k) {
          const float av = __shfl_sync(kFull, my_attr[k], group_base + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) m[q] = fmaf(cw[q][k], av, m[q]);
        }
        unsigned bits = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool on = on_edge && m[q] > 0.f;
          bits |= on ? (1u << q) : 0u;
          acc[q] += on ? m[q] : 0.f;
          cn[q] += on ? 1.f : 0.f;
        }
        if (a.mask != nullptr) {  // warp-uniform branch
          // assemble the 32-bit mask of the edge from the 8 lanes of the group (4 bits each); xor 1/2/4 stay in the group
          unsigned word = bits << (4 * sl);
          word |= __shfl_xor_sync(kFull, word, 1);
          word |= __shfl_xor_sync(kFull, word, 2);
          word |= __shfl_xor_sync(kFull, word, 4);
          if (on_edge && sl == 0) a.mask[eid] = word;
        }
      }
    }
    if (!row_ok) continue;
    *reinterpret_cast<float4*>(a.s + (size_t)r * a.ld_s + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (a.cnt != nullptr) *reinterpret_cast<float4*>(a.cnt + (size_t)r * kMsg + c) = make_float4(cn[0], cn[1], cn[2], cn[3]);
  }
}

// backward, source side:  dV[j,c] = sum_{t in CSC segment j} dS[row_t, c] * mask[eid_t][c]
struct EdgeMsgBwdSrcArgs {
  const int32_t* ptr;   // CSC by source
  const int32_t* idx;   // destination node of each CSC slot
  const int32_t* perm;  // original edge id of each CSC slot
  const float* ds;      // [N, 32]
  const uint32_t* mask; // [E]
  float* dv;            // [N, 32] (strided: ld_dv)
  uint32_t ld_ds, ld_dv;
  int32_t n, rows_per_block;
};

__global__ void __launch_bounds__(256) k_edge_msg_bwd_src(const EdgeMsgBwdSrcArgs a) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = lane_id();
  const int sub = lane >> 3, sl = lane & 7;
  const int group_base = sub * 8;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n);
  const int c = sl * 4;
  for (int rw = row0 + warp * 4; rw < row_end; rw += 32) {
    const int r = rw + sub;
    const bool row_ok = r < row_end;
    int beg = 0, len = 0;
    if (row_ok) {
      beg = __ldg(a.ptr + r);
      len = __ldg(a.ptr + r + 1) - beg;
    }
    int max_len = len;
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 16));
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 8));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int off = 0; off < max_len; off += 8) {
      int my_dst = -1;
      unsigned my_mask = 0;
      if (off + sl < len) {
        my_dst = ld_stream_i32(a.idx + beg + off + sl);
        my_mask = a.mask[ld_stream_i32(a.perm + beg + off + sl)];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int drow = __shfl_sync(kFull, my_dst, group_base + j);
        const unsigned word = __shfl_sync(kFull, my_mask, group_base + j);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (drow >= 0) g = ld_gather_f4(a.ds + (size_t)drow * a.ld_ds + c);
        const unsigned bits = drow >= 0 ? word >> (4 * sl) : 0u;
        acc[0] += (bits & 1u) ? g.x : 0.f;
        acc[1] += (bits & 2u) ? g.y : 0.f;
        acc[2] += (bits & 4u) ? g.z : 0.f;
        acc[3] += (bits & 8u) ? g.w : 0.f;
      }
    }
    if (row_ok) *reinterpret_cast<float4*>(a.dv + (size_t)r * a.ld_dv + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// backward, edge-feature weights:  dC[c,k] = sum_e dS[row_e,c] * mask_e[c] * attr[e,k]
// per destination row: t[c,k] = sum_{e in row} mask_e[c] attr[e,k];  partial[block][c,k] += dS[i,c] t[c,k]
struct EdgeMsgBwdCArgs {
  const int32_t* ptr;
  const int32_t* perm;
  const float* ds;
  const uint32_t* mask;
  const float* attr;
  float* partial;  // [gridDim.x][32][kMaxEdgeFeat]
  uint32_t ld_ds, ld_attr;
  int32_t n, fe;
};

constexpr int kBwdCWarps = 32;  // 1024 threads: the kernel is a chain of dependent global loads per row, it needs warps, not registers
template <int FE>
__global__ void __launch_bounds__(kBwdCWarps * 32) k_edge_msg_bwd_c(const EdgeMsgBwdCArgs a) {
  // One warp per destination row, lane = channel c for the accumulation.  The row's edges are fetched LANE-PARALLEL (edge id, ReLU mask
  // word and attributes of up to 32 edges in flight at once: three dependent global loads per 32 edges instead of per edge) and then
  // broadcast to the channels with shuffles, in CSR order, so the sum keeps its fixed association.
  __shared__ float red[kBwdCWarps][kMsg][kMaxEdgeFeat];
  const int c = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float acc[FE];
#pragma unroll
  for (int k = 0; k < FE; ++k) acc[k] = 0.f;
  const int rows_per_block = (a.n + gridDim.x - 1) / gridDim.x;
  const int row0 = blockIdx.x * rows_per_block;
  const int row_end = min(row0 + rows_per_block, a.n);
  for (int r = row0 + warp; r < row_end; r += kBwdCWarps) {
    const int beg = __ldg(a.ptr + r), end = __ldg(a.ptr + r + 1);
    const float g = __ldg(a.ds + (size_t)r * a.ld_ds + c);
    float t[FE];
#pragma unroll
    for (int k = 0; k < FE; ++k) t[k] = 0.f;
    for (int s0 = beg; s0 < end; s0 += 32) {
      const int s = s0 + c;
      const bool valid = s < end;
      const int eid = valid ? (a.perm != nullptr ? __ldg(a.perm + s) : s) : 0;
      const uint32_t m = valid ? __ldg(a.mask + eid) : 0u;
      float av[FE];
#pragma unroll
      for (int k = 0; k < FE; ++k) av[k] = valid ? __ldg(a.attr + (size_t)eid * a.ld_attr + k) : 0.f;
      const int cnt = min(32, end - s0);
      for (int j = 0; j < cnt; ++j) {
        const bool on = (__shfl_sync(0xffffffffu, m, j) >> c) & 1u;
#pragma unroll
        for (int k = 0; k < FE; ++k) {
          const float v = __shfl_sync(0xffffffffu, av[k], j);
          t[k] += on ? v : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < FE; ++k) acc[k] = fmaf(g, t[k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < kMaxEdgeFeat; ++k) red[warp][c][k] = k < FE ? acc[k < FE ? k : 0] : 0.f;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < kMaxEdgeFeat; ++k) {
      float s = 0.f;
      for (int w = 0; w < kBwdCWarps; ++w) s += red[w][c][k];
      a.partial[((size_t)blockIdx.x * kMsg + c) * kMaxEdgeFeat + k] = s;
    }
  }
}

__global__ void k_edge_msg_bwd_c_reduce(const float* __restrict__ partial, int n_partials, int fe, float* __restrict__ dc, int64_t ld_dc) {
  const int c = threadIdx.x & 31;
  const int k = threadIdx.x >> 5;
  if (k >= fe) return;
  float s = 0.f;
  for (int p = 0; p < n_partials; ++p) s += partial[((size_t)p * kMsg + c) * kMaxEdgeFeat + k];
  dc[(size_t)c * ld_dc + k] = s;
}

static int rows_per_block_for(int n, int rows_per_pass) {
  int passes = (int)ceil_div<int64_t>(n, (int64_t)kNumSM * 8 * rows_per_pass);
  passes = std::max(1, std::min(passes, 8));
  return rows_per_pass * passes;
}

}  // namespace drk

extern "C" {

int drk_edge_msg_fwd(const int32_t* rowptr, const int32_t* colidx, const int32_t* perm, const float* uv, int64_t ld_uv,
                     const float* edge_attr, int64_t ld_attr, int32_t num_edge_features, const float* cmat, int64_t ld_c, float* s,
                     int64_t ld_s, float* cnt, uint32_t* mask, int32_t num_nodes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && num_edge_features >= 0, DRK_EINVAL, "edge msg: negative size");
  DRK_REQUIRE(num_edge_features <= kMaxEdgeFeat, DRK_EUNSUPPORTED, "edge msg: at most %d edge features are supported", kMaxEdgeFeat);
  if (num_nodes == 0) return DRK_OK;
  DRK_REQUIRE(rowptr && uv && s && (num_edge_features == 0 || (edge_attr && cmat)), DRK_EINVAL, "edge msg: null pointer");
  DRK_REQUIRE(ld_uv % 4 == 0 && ld_s % 4 == 0 && aligned16(uv) && aligned16(s) && (cnt == nullptr || aligned16(cnt)), DRK_EUNSUPPORTED,
              "edge msg: U|V and S must be 16-byte aligned with leading dimensions divisible by 4");
  EdgeMsgArgs a{rowptr, colidx, perm, uv, edge_attr, cmat, s, cnt, mask, (uint32_t)ld_uv, (uint32_t)ld_attr, (uint32_t)ld_c, (uint32_t)ld_s,
                num_nodes, num_edge_features, rows_per_block_for(num_nodes, 32)};
  const int blocks = ceil_div(num_nodes, a.rows_per_block);
  cudaStream_t st = as_stream(stream);
  switch (num_edge_features) {
    case 0: k_edge_msg_fwd<0><<<blocks, 256, 0, st>>>(a); break;
    case 1: k_edge_msg_fwd<1><<<blocks, 256, 0, st>>>(a); break;
    case 2: k_edge_msg_fwd<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: k_edge_msg_fwd<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: k_edge_msg_fwd<4><<<blocks, 256, 0, st>>>(a); break;
    case 5: k_edge_msg_fwd<5><<<blocks, 256, 0, st>>>(a); break;
    case 6: k_edge_msg_fwd<6><<<blocks, 256, 0, st>>>(a); break;
    case 7: k_edge_msg_fwd<7><<<blocks, 256, 0, st>>>(a); break;
    default: k_edge_msg_fwd<8><<<blocks, 256, 0, st>>>(a); break;
  }
  return finish_launch("edge msg fwd");
}

int drk_edge_msg_bwd_src(const int32_t* colptr, const int32_t* rowidx, const int32_t* permT, const float* ds, int64_t ld_ds,
                         const uint32_t* mask, float* dv, int64_t ld_dv, int32_t num_nodes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0, DRK_EINVAL, "edge msg bwd src: negative size");
  if (num_nodes == 0) return DRK_OK;
  DRK_REQUIRE(colptr && ds && mask && dv, DRK_EINVAL, "edge msg bwd src: null pointer");
  DRK_REQUIRE(ld_ds % 4 == 0 && ld_dv % 4 == 0 && aligned16(ds) && aligned16(dv), DRK_EUNSUPPORTED, "edge msg bwd src: alignment");
  EdgeMsgBwdSrcArgs a{colptr, rowidx, permT, ds, mask, dv, (uint32_t)ld_ds, (uint32_t)ld_dv, num_nodes, rows_per_block_for(num_nodes, 32)};
  k_edge_msg_bwd_src<<<ceil_div(num_nodes, a.rows_per_block), 256, 0, as_stream(stream)>>>(a);
  return finish_launch("edge msg bwd src");
}

size_t drk_edge_msg_bwd_c_workspace_bytes(void) { return (size_t)drk::kNumSM * 2 * drk::kMsg * drk::kMaxEdgeFeat * sizeof(float); }

int drk_edge_msg_bwd_c(const int32_t* rowptr, const int32_t* perm, const float* ds, int64_t ld_ds, const uint32_t* mask,
                       const float* edge_attr, int64_t ld_attr, int32_t num_edge_features, float* dc, int64_t ld_dc, int32_t num_nodes,
                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && num_edge_features >= 0 && num_edge_features <= kMaxEdgeFeat, DRK_EINVAL, "edge msg bwd c: bad size");
  if (num_edge_features == 0) return DRK_OK;
  DRK_REQUIRE(rowptr && ds && mask && edge_attr && dc, DRK_EINVAL, "edge msg bwd c: null pointer");
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= drk_edge_msg_bwd_c_workspace_bytes(), DRK_EWORKSPACE, "edge msg bwd c: workspace too small");
  const int blocks = kNumSM * 2;
  EdgeMsgBwdCArgs a{rowptr, perm, ds, mask, edge_attr, static_cast<float*>(workspace), (uint32_t)ld_ds, (uint32_t)ld_attr, num_nodes, num_edge_features};
  cudaStream_t st = as_stream(stream);
  switch (num_edge_features) {
    case 1: k_edge_msg_bwd_c<1><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 2: k_edge_msg_bwd_c<2><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 3: k_edge_msg_bwd_c<3><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 4: k_edge_msg_bwd_c<4><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 5: k_edge_msg_bwd_c<5><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 6: k_edge_msg_bwd_c<6><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    case 7: k_edge_msg_bwd_c<7><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
    default: k_edge_msg_bwd_c<8><<<blocks, kBwdCWarps * 32, 0, st>>>(a); break;
  }
  k_edge_msg_bwd_c_reduce<<<1, 32 * kMaxEdgeFeat, 0, as_stream(stream)>>>(static_cast<float*>(workspace), blocks, num_edge_features, dc, ld_dc);
  return finish_launch("edge msg bwd c", 2);
}

}  // extern "C"

// =====================================================================================================
// Segment max with argmax (torch_scatter.scatter_max, community_pooling.py:209; max_pool_x ginet.py:103)
// and the cluster-offset pass of get_preloaded_cluster (community_pooling.py:23-27).
// =====================================================================================================
namespace drk {

// thread -> (segment c, column f).  Elements of a segment are visited in ascending element id, ">" keeps the
// FIRST maximum (torch_scatter's CPU rule); an empty segment gives out = 0, arg = n_src.
__global__ void __launch_bounds__(256) k_segment_max(const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                                                     const float* __restrict__ src, uint32_t ld_src, int32_t n_src, int32_t n_seg,
                                                     int32_t width, float* __restrict__ out, uint32_t ld_out, int32_t* __restrict__ arg) {
  const uint32_t total = (uint32_t)n_seg * (uint32_t)width;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t c = t / (uint32_t)width;
    const uint32_t f = t - c * (uint32_t)width;
    const int beg = __ldg(ptr + c), end = __ldg(ptr + c + 1);
    float best = 0.f;
    int best_i = n_src;
    for (int s = beg; s < end; ++s) {
      const int i = perm != nullptr ? __ldg(perm + s) : s;
      const float v = __ldg(src + (size_t)i * ld_src + f);
      if (s == beg || v > best || (v != v && best == best)) {  // NaN propagates like torch's max
        best = v;
        best_i = i;
      }
    }
    out[(size_t)c * ld_out + f] = best;
    if (arg != nullptr) arg[(size_t)c * width + f] = best_i;
  }
}

// dsrc must be zero-filled by the caller; each (c,f) owns a distinct target element -> no atomics.
__global__ void __launch_bounds__(256) k_segment_max_bwd(const float* __restrict__ dout, uint32_t ld_dout, const int32_t* __restrict__ arg,
                                                         int32_t n_src, int32_t n_seg, int32_t width, float* __restrict__ dsrc,
                                                         uint32_t ld_dsrc) {
  const uint32_t total = (uint32_t)n_seg * (uint32_t)width;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t c = t / (uint32_t)width;
    const uint32_t f = t - c * (uint32_t)width;
    const int i = arg[t];
    if (i >= 0 && i < n_src) dsrc[(size_t)i * ld_dsrc + f] = dout[(size_t)c * ld_dout + f];
  }
}

// get_preloaded_cluster: cluster[i] += sum_{h < batch[i]} (max_{j in graph h} cluster[j] + 1)
__global__ void __launch_bounds__(256) k_cluster_graph_max(const int64_t* __restrict__ cluster, const int32_t* __restrict__ graph_ptr,
                                                           long long* __restrict__ gmax) {
  __shared__ long long red[8];
  const int g = blockIdx.x;
  const int beg = graph_ptr[g], end = graph_ptr[g + 1];
  long long m = -1;  // empty graph contributes max+1 = 0
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) m = max(m, (long long)cluster[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane_id() == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = max(m, red[w]);
    gmax[g] = m;
  }
}

__global__ void __launch_bounds__(1024) k_cluster_offsets_scan(const long long* __restrict__ gmax, int32_t num_graphs, long long* __restrict__ offs) {
  // single block; offs[g] = sum_{h<g} (gmax[h] + 1), offs[num_graphs] = total number of cluster ids
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < num_graphs; base += 1024) {
    const int g = base + threadIdx.x;
    const long long v = g < num_graphs ? gmax[g] + 1 : 0;
    long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane_id() >= o) incl += up;
    }
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    long long woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_tot[w];
    if (g < num_graphs) offs[g] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offs[num_graphs] = carry;
}

__global__ void __launch_bounds__(256) k_cluster_add_offsets(int64_t* __restrict__ cluster, const int32_t* __restrict__ batch32,
                                                             const long long* __restrict__ offs, int32_t n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cluster[i] += offs[batch32[i]];
}

}  // namespace drk

extern "C" {

int drk_segment_max(const int32_t* ptr, const int32_t* perm, const float* src, int64_t ld_src, int32_t n_src, int32_t n_seg, int32_t width,
                    float* out, int64_t ld_out, int32_t* arg, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n_src >= 0 && n_seg >= 0 && width >= 0, DRK_EINVAL, "segment max: negative size");
  if (n_seg == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(ptr && out && (src || n_src == 0), DRK_EINVAL, "segment max: null pointer");
  DRK_REQUIRE((int64_t)n_seg * width < (int64_t)1 << 31, DRK_EUNSUPPORTED, "segment max: more than 2^31 outputs");
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>((int64_t)n_seg * width, 256), (int64_t)kNumSM * 32);
  k_segment_max<<<blocks, 256, 0, as_stream(stream)>>>(ptr, perm, src, (uint32_t)ld_src, n_src, n_seg, width, out, (uint32_t)ld_out, arg);
  return finish_launch("segment max");
}

int drk_segment_max_bwd(const float* dout, int64_t ld_dout, const int32_t* arg, int32_t n_src, int32_t n_seg, int32_t width, float* dsrc,
                        int64_t ld_dsrc, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n_src >= 0 && n_seg >= 0 && width >= 0, DRK_EINVAL, "segment max bwd: negative size");
  if (n_seg == 0 || width == 0 || n_src == 0) return DRK_OK;
  DRK_REQUIRE(dout && arg && dsrc, DRK_EINVAL, "segment max bwd: null pointer");
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>((int64_t)n_seg * width, 256), (int64_t)kNumSM * 32);
  k_segment_max_bwd<<<blocks, 256, 0, as_stream(stream)>>>(dout, (uint32_t)ld_dout, arg, n_src, n_seg, width, dsrc, (uint32_t)ld_dsrc);
  return finish_launch("segment max bwd");
}

size_t drk_cluster_offsets_workspace_bytes(int32_t num_graphs) { return num_graphs < 0 ? 0 : ((size_t)2 * num_graphs + 2) * sizeof(long long); }

int drk_cluster_offsets(int64_t* cluster, const int32_t* graph_ptr, const int32_t* batch32, int32_t num_nodes, int32_t num_graphs,
                        int64_t* total_ids, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && num_graphs >= 0, DRK_EINVAL, "cluster offsets: negative size");
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= drk_cluster_offsets_workspace_bytes(num_graphs), DRK_EWORKSPACE,
              "cluster offsets: workspace too small");
  if (num_graphs == 0) return DRK_OK;
  DRK_REQUIRE(graph_ptr && batch32 && (cluster || num_nodes == 0), DRK_EINVAL, "cluster offsets: null pointer");
  long long* gmax = static_cast<long long*>(workspace);
  long long* offs = gmax + num_graphs;
  cudaStream_t st = as_stream(stream);
  k_cluster_graph_max<<<num_graphs, 256, 0, st>>>(cluster, graph_ptr, gmax);
  k_cluster_offsets_scan<<<1, 1024, 0, st>>>(gmax, num_graphs, offs);
  if (num_nodes > 0) k_cluster_add_offsets<<<ceil_div(num_nodes, 256), 256, 0, st>>>(cluster, batch32, offs, num_nodes);
  if (total_ids != nullptr) {
    cudaError_t e = cudaMemcpyAsync(total_ids, offs + num_graphs, sizeof(long long), cudaMemcpyDeviceToDevice, st);
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "cluster offsets: copy: %s", cudaGetErrorString(e));
  }
  return finish_launch("cluster offsets", 3);
}

}  // extern "C"
