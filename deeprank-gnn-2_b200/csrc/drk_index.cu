// Graph index construction on the device: stable destination-sorted CSR + source-sorted CSC
// of a batch's int64 edge list, generic segment index of one key vector, batch offsets.
//
// Replaces the implicit index handling of the reference path (`row, col = edge_index`,
// torch_scatter's index broadcast + CPU scatter_add_ visiting order; ginet.py:41,58,
// vanilla_gnn.py:28,35, foutnet.py:57; PyG collate `ptr`).  All outputs are integers and are
// bit-exact against torch.sort(stable=True) / bincount / cumsum (oracle/restate.py:graph_csr).
//
// Algorithm (counting sort, deterministic result):
//   1. histogram of keys        (integer atomics: the counts do not depend on their order)
//   2. exclusive scan           (two small kernels; N+1 counters per key array)
//   3. fill                     (atomic cursor per segment -> arbitrary order inside a segment)
//   4. per-segment rank sort    (edge ids ascending inside each segment -> the stable order),
//      fused with the gather of the "other" endpoint (colidx = col[perm]).
// HBM traffic ~ 16E (read int64 edges twice) + 8E (tmp) + 16E (perm/idx) bytes per key array.
#include <algorithm>

#include "drk_common.cuh"

namespace drk {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                             // per thread
constexpr int kScanTile = kScanThreads * kScanItems;      // 2048 counters per block

// ---------------------------------------------------------------- 1. histogram
__global__ void __launch_bounds__(256) k_key_hist(const int64_t* __restrict__ key0, const int64_t* __restrict__ key1,
                                                  int64_t n, int32_t num_segments, int32_t* __restrict__ cnt0,
                                                  int32_t* __restrict__ cnt1, int32_t* __restrict__ status) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const unsigned long long a = (unsigned long long)ld_stream_i64(key0 + e);
    const bool ok0 = a < (unsigned long long)num_segments;
    bool ok1 = true;
    unsigned long long b = 0;
    if (key1 != nullptr) {
      b = (unsigned long long)ld_stream_i64(key1 + e);
      ok1 = b < (unsigned long long)num_segments;
    }
    if (ok0 && ok1) {  // an edge with a bad endpoint is dropped from BOTH orders
      atomicAdd(cnt0 + a, 1);
      if (key1 != nullptr) atomicAdd(cnt1 + b, 1);
    } else {
      bad = true;
    }
  }
  if (bad && status != nullptr) atomicOr(status, DRK_STATUS_INDEX_RANGE);
}

// ---------------------------------------------------------------- 2. exclusive scan (grid.y = key array)
__device__ __forceinline__ int block_reduce_sum(int v, int* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 0) smem[warp] = v;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += smem[w];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int32_t* __restrict__ cnt, int32_t len, int64_t array_stride,
                                                                 int32_t* __restrict__ tile_sums, int32_t tiles) {
  __shared__ int smem[kScanThreads / 32];
  const int32_t* c = cnt + (int64_t)blockIdx.y * array_stride;
  const int base = blockIdx.x * kScanTile;
  int v = 0;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
    const int idx = base + i;
    if (idx < len) v += c[idx];
  }
  const int total = block_reduce_sum(v, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.y * tiles + blockIdx.x] = total;
}

// ptr[i] = sum_{j<i} cnt[j] for i in [0, len); cnt is zeroed afterwards (it becomes the fill cursor).
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(int32_t* __restrict__ cnt, int32_t len, int64_t array_stride,
                                                             const int32_t* __restrict__ tile_sums, int32_t tiles,
                                                             int32_t* __restrict__ ptr0, int32_t* __restrict__ ptr1) {
  __shared__ int smem[kScanThreads / 32];
  __shared__ int warp_prefix[kScanThreads / 32];
  int32_t* c = cnt + (int64_t)blockIdx.y * array_stride;
  int32_t* ptr = blockIdx.y == 0 ? ptr0 : ptr1;
  // prefix of the tiles before this one (tiles is small: N / 2048)
  int before = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += kScanThreads) before += tile_sums[blockIdx.y * tiles + t];
  before = block_reduce_sum(before, smem);

  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int item[kScanItems];
  int local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int idx = base + i;
    item[i] = idx < len ? c[idx] : 0;
    local += item[i];
  }
  // exclusive scan of `local` across the block
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane_id() >= o) incl += up;
  }
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 31) warp_prefix[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_prefix[w];
  int run = before + woff + incl - local;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int idx = base + i;
    if (idx < len) {
      ptr[idx] = run;
      c[idx] = 0;
    }
    run += item[i];
  }
}

// ---------------------------------------------------------------- 3. fill (unordered inside a segment)
__global__ void __launch_bounds__(256) k_key_fill(const int64_t* __restrict__ key0, const int64_t* __restrict__ key1, int64_t n,
                                                  int32_t num_segments, const int32_t* __restrict__ ptr0,
                                                  const int32_t* __restrict__ ptr1, int32_t* __restrict__ cur0,
                                                  int32_t* __restrict__ cur1, int32_t* __restrict__ tmp0,
                                                  int32_t* __restrict__ tmp1) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const unsigned long long a = (unsigned long long)ld_stream_i64(key0 + e);
    unsigned long long b = 0;
    bool ok = a < (unsigned long long)num_segments;
    if (key1 != nullptr) {
      b = (unsigned long long)ld_stream_i64(key1 + e);
      ok = ok && b < (unsigned long long)num_segments;
    }
    if (!ok) continue;
    tmp0[ptr0[a] + atomicAdd(cur0 + a, 1)] = (int32_t)e;
    if (key1 != nullptr) tmp1[ptr1[b] + atomicAdd(cur1 + b, 1)] = (int32_t)e;
  }
}

// the gathered endpoint is only range-checked by the histogram when it is also a sort key (CSR+CSC build);
// in the single-key build it is checked here so that a bad index can never become an out-of-bounds gather.
__device__ __forceinline__ int32_t checked_endpoint(long long v, int32_t n, int32_t* status) {
  if ((unsigned long long)v < (unsigned long long)n) return (int32_t)v;
  if (status != nullptr) atomicOr(status, DRK_STATUS_INDEX_RANGE);
  return 0;
}

// ---------------------------------------------------------------- 4. rank sort inside each segment
// One warp per segment.  Element ids are distinct, so rank(v) = #{u in segment : u < v}.
// O(L^2/32) shuffles per lane; L is a node degree (~20, max ~130 on 15 A residue graphs).
__global__ void __launch_bounds__(256) k_segment_rank_sort(const int32_t* __restrict__ ptr0, const int32_t* __restrict__ ptr1,
                                                           const int32_t* __restrict__ tmp0, const int32_t* __restrict__ tmp1,
                                                           const int64_t* __restrict__ other0, const int64_t* __restrict__ other1,
                                                           int32_t num_segments, int32_t* __restrict__ perm0,
                                                           int32_t* __restrict__ perm1, int32_t* __restrict__ idx0,
                                                           int32_t* __restrict__ idx1, int32_t* __restrict__ status) {
  const int warps_per_block = blockDim.x >> 5;
  const int64_t total = (int64_t)num_segments * (ptr1 != nullptr ? 2 : 1);
  const int lane = lane_id();
  for (int64_t w = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < total; w += (int64_t)gridDim.x * warps_per_block) {
    const bool second = w >= num_segments;
    const int seg = (int)(second ? w - num_segments : w);
    const int32_t* ptr = second ? ptr1 : ptr0;
    const int32_t* tmp = second ? tmp1 : tmp0;
    const int64_t* other = second ? other1 : other0;
    int32_t* perm = second ? perm1 : perm0;
    int32_t* idx = second ? idx1 : idx0;
    const int begin = ptr[seg];
    const int len = ptr[seg + 1] - begin;
    if (len <= 0) continue;
    if (len <= 32) {
      const int v = lane < len ? tmp[begin + lane] : 0x7fffffff;
      int rank = 0;
      for (int j = 0; j < len; ++j) rank += (__shfl_sync(0xffffffffu, v, j) < v) ? 1 : 0;
      if (lane < len) {
        perm[begin + rank] = v;
        if (idx != nullptr) idx[begin + rank] = checked_endpoint(other[v], num_segments, status);
      }
    } else {
      for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        const int v = i < len ? tmp[begin + i] : 0x7fffffff;
        int rank = 0;
        for (int j0 = 0; j0 < len; j0 += 32) {
          const int u = (j0 + lane) < len ? tmp[begin + j0 + lane] : 0x7fffffff;
          const int lim = min(32, len - j0);
          for (int j = 0; j < lim; ++j) rank += (__shfl_sync(0xffffffffu, u, j) < v) ? 1 : 0;
        }
        if (i < len) {
          perm[begin + rank] = v;
          if (idx != nullptr) idx[begin + rank] = checked_endpoint(other[v], num_segments, status);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- batch offsets
__global__ void __launch_bounds__(256) k_batch_offsets(const int64_t* __restrict__ batch, int32_t n, int32_t num_graphs,
                                                       int32_t* __restrict__ graph_ptr, int32_t* __restrict__ batch32,
                                                       int32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    if (i <= num_graphs) graph_ptr[i] = 0;
    return;
  }
  if (i >= n) return;
  const long long b = batch[i];
  int flags = 0;
  if (b < 0 || b >= num_graphs) flags |= DRK_STATUS_INDEX_RANGE;
  const long long bc = b < 0 ? 0 : (b >= num_graphs ? num_graphs - 1 : b);
  if (batch32 != nullptr) batch32[i] = (int32_t)bc;
  long long prev = -1;
  if (i > 0) {
    prev = batch[i - 1];
    if (prev > b) flags |= DRK_STATUS_UNSORTED;
    prev = prev < 0 ? 0 : (prev >= num_graphs ? num_graphs - 1 : prev);
  }
  // node i opens every graph in (prev, b]: empty graphs in between start (and end) here
  for (long long g = prev + 1; g <= bc; ++g) graph_ptr[g] = i;
  if (i == n - 1)
    for (long long g = bc + 1; g <= num_graphs; ++g) graph_ptr[g] = n;
  if (flags != 0 && status != nullptr) atomicOr(status, flags);
}

// ---------------------------------------------------------------- gather rows
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ perm,
                                                     int64_t n, int32_t width, float* __restrict__ out, int64_t ld_out) {
  const int64_t total = n * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / width;
    const int c = (int)(t - r * width);
    out[r * ld_out + c] = src[(int64_t)perm[r] * ld_src + c];
  }
}

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct IndexWorkspace {
  int32_t* cnt;        // [keys][num_segments + 1]
  int64_t cnt_stride;  // elements between the two counter arrays
  int32_t* tile_sums;  // [keys][tiles]
  int32_t* tmp0;       // [n]
  int32_t* tmp1;       // [n]
  int32_t tiles;
  size_t bytes;
};

static IndexWorkspace carve(void* base, int64_t n, int32_t num_segments, int keys) {
  IndexWorkspace w{};
  const int32_t len = num_segments + 1;
  w.tiles = (int32_t)ceil_div<int64_t>(len, kScanTile);
  w.cnt_stride = (int64_t)(align_up((size_t)len * 4) / 4);
  size_t off = 0;
  char* p = static_cast<char*>(base);
  w.cnt = reinterpret_cast<int32_t*>(p + off);
  off += (size_t)w.cnt_stride * 4 * keys;
  w.tile_sums = reinterpret_cast<int32_t*>(p + off);
  off += align_up((size_t)w.tiles * 4 * keys);
  w.tmp0 = reinterpret_cast<int32_t*>(p + off);
  off += align_up((size_t)n * 4);
  w.tmp1 = reinterpret_cast<int32_t*>(p + off);
  if (keys == 2) off += align_up((size_t)n * 4);
  w.bytes = off;
  return w;
}

// key0/key1: the vectors sorted on (key1 may be NULL); other0/other1: the int64 vectors gathered into idx0/idx1
// in the sorted order (NULL = no gather).
static int build(const int64_t* key0, const int64_t* key1, const int64_t* other0, const int64_t* other1, int64_t n,
                 int32_t num_segments, int32_t* ptr0, int32_t* idx0, int32_t* perm0, int32_t* ptr1, int32_t* idx1, int32_t* perm1,
                 int32_t* status, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int keys = key1 != nullptr ? 2 : 1;
  DRK_REQUIRE(n >= 0 && num_segments >= 0, DRK_EINVAL, "index build: negative size");
  DRK_REQUIRE(n < (int64_t)0x7fffffff, DRK_EUNSUPPORTED, "index build: more than 2^31-1 elements");
  DRK_REQUIRE(ptr0 != nullptr && (perm0 != nullptr || n == 0), DRK_EINVAL, "index build: null output");
  IndexWorkspace w = carve(workspace, n, num_segments, keys);
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= w.bytes, DRK_EWORKSPACE, "index build: workspace %zu < %zu bytes",
              workspace_bytes, w.bytes);
  const int32_t len = num_segments + 1;
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, (size_t)w.cnt_stride * 4 * keys, stream);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "index build: memset: %s", cudaGetErrorString(e));
  int32_t* cnt1 = keys == 2 ? w.cnt + w.cnt_stride : nullptr;
  int launches = 0;
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256 * 4), (int64_t)kNumSM * 16);
    k_key_hist<<<blocks, 256, 0, stream>>>(key0, key1, n, num_segments, w.cnt, cnt1, status);
    ++launches;
  }
  {
    dim3 grid(w.tiles, keys);
    k_scan_tile_sums<<<grid, kScanThreads, 0, stream>>>(w.cnt, len, w.cnt_stride, w.tile_sums, w.tiles);
    k_scan_apply<<<grid, kScanThreads, 0, stream>>>(w.cnt, len, w.cnt_stride, w.tile_sums, w.tiles, ptr0, ptr1);
    launches += 2;
  }
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256 * 4), (int64_t)kNumSM * 16);
    k_key_fill<<<blocks, 256, 0, stream>>>(key0, key1, n, num_segments, ptr0, ptr1, w.cnt, cnt1, w.tmp0, w.tmp1);
    const int64_t warps = (int64_t)num_segments * keys;
    const int sblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(warps, 8), (int64_t)kNumSM * 32));
    k_segment_rank_sort<<<sblocks, 256, 0, stream>>>(ptr0, ptr1, w.tmp0, w.tmp1, other0, other1, num_segments, perm0, perm1,
                                                     other0 != nullptr ? idx0 : nullptr, other1 != nullptr ? idx1 : nullptr, status);
    launches += 2;
  }
  return finish_launch("index build", launches);
}

}  // namespace drk

extern "C" {

size_t drk_graph_index_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
  if (num_edges < 0 || num_nodes < 0) return 0;
  return drk::carve(nullptr, num_edges, num_nodes, 2).bytes;
}

int drk_graph_index_build(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes, int32_t* rowptr, int32_t* colidx,
                          int32_t* perm, int32_t* colptr, int32_t* rowidx, int32_t* permT, int32_t* status, void* workspace,
                          size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(edge_index != nullptr || num_edges == 0, DRK_EINVAL, "graph index: null edge_index");
  DRK_REQUIRE(rowptr && (num_edges == 0 || (colidx && perm)), DRK_EINVAL, "graph index: null CSR output");
  const bool with_csc = colptr != nullptr;  // (rowidx / permT may legitimately be NULL when E == 0)
  DRK_REQUIRE(!with_csc || num_edges == 0 || (rowidx && permT), DRK_EINVAL, "graph index: CSC outputs must be all set or all NULL");
  if (num_edges == 0) {  // no edges: both pointer arrays are all zero, nothing to sort
    DRK_REQUIRE(num_nodes >= 0, DRK_EINVAL, "graph index: negative size");
    cudaError_t e = cudaMemsetAsync(rowptr, 0, ((size_t)num_nodes + 1) * sizeof(int32_t), as_stream(stream));
    if (e == cudaSuccess && with_csc) e = cudaMemsetAsync(colptr, 0, ((size_t)num_nodes + 1) * sizeof(int32_t), as_stream(stream));
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "graph index: memset: %s", cudaGetErrorString(e));
    return DRK_OK;
  }
  const int64_t* row = edge_index;
  const int64_t* col = edge_index + num_edges;
  if (with_csc)
    return build(row, col, col, row, num_edges, num_nodes, rowptr, colidx, perm, colptr, rowidx, permT, status, workspace, workspace_bytes,
                 as_stream(stream));
  return build(row, nullptr, col, nullptr, num_edges, num_nodes, rowptr, colidx, perm, nullptr, nullptr, nullptr, status, workspace,
               workspace_bytes, as_stream(stream));
}

size_t drk_segment_index_workspace_bytes(int64_t n, int32_t num_segments) {
  if (n < 0 || num_segments < 0) return 0;
  return drk::carve(nullptr, n, num_segments, 1).bytes;
}

int drk_segment_index_build(const int64_t* index, int64_t n, int32_t num_segments, int32_t* ptr, int32_t* perm, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(index != nullptr || n == 0, DRK_EINVAL, "segment index: null index");
  return build(index, nullptr, nullptr, nullptr, n, num_segments, ptr, nullptr, perm, nullptr, nullptr, nullptr, status, workspace,
               workspace_bytes, as_stream(stream));
}

int drk_batch_offsets(const int64_t* batch, int32_t num_nodes, int32_t num_graphs, int32_t* graph_ptr, int32_t* batch32,
                      int32_t* status, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && num_graphs >= 0, DRK_EINVAL, "batch offsets: negative size");
  DRK_REQUIRE(graph_ptr != nullptr && (batch != nullptr || num_nodes == 0), DRK_EINVAL, "batch offsets: null pointer");
  const int work = num_nodes > 0 ? num_nodes : num_graphs + 1;
  k_batch_offsets<<<ceil_div(work, 256), 256, 0, as_stream(stream)>>>(batch, num_nodes, num_graphs, graph_ptr, batch32, status);
  return finish_launch("batch offsets");
}

int drk_gather_rows(const float* src, int64_t ld_src, const int32_t* perm, int64_t n, int32_t width, float* out, int64_t ld_out,
                    void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0 && width >= 0, DRK_EINVAL, "gather rows: negative size");
  if (n == 0 || width == 0) return DRK_OK;
  DRK_REQUIRE(src && perm && out, DRK_EINVAL, "gather rows: null pointer");
  const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n * width, 256 * 4), (int64_t)kNumSM * 16);
  k_gather_rows<<<blocks, 256, 0, as_stream(stream)>>>(src, ld_src, perm, n, width, out, ld_out);
  return finish_launch("gather rows");
}

}  // extern "C"
