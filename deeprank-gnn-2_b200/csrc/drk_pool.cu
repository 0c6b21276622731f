// Community pooling structure on the device (reference: deeprank2/utils/community_pooling.py:165-242 and the PyG 2.4 helpers it
// calls -- consecutive_cluster, pool_edge (relabel, remove_self_loops, coalesce), pool_batch).
//
// The data-dependent sizes (number of distinct clusters, number of distinct pooled edges) are properties of the GRAPHS, not of the
// weights: the host collate knows them (or an upper bound) when it builds the batch, so every output below is allocated by the caller
// and nothing is read back -- a whole train step of a clustered network can be captured into a CUDA graph.  Each kernel also writes
// the count it found to a device scalar and raises DRK_STATUS_INDEX_RANGE if it exceeds the capacity it was given.
//
//   consecutive_cluster(c):   drk_segment_index_build(c, K)            nodes grouped by cluster id, stable  (drk_index.cu)
//                             drk_compact_segments                     empty ids dropped: rank[id] (= torch.unique's inverse through
//                                                                      rank[c]), compact ptr, last member of every cluster (= perm)
//   pool_edge(c, ei, ea):     drk_pool_edge_keys                       key = dense id of the pooled pair inside its graph's C_g x C_g
//                                                                      block (self loops -> DRK_POOL_JUNK_SEGMENTS junk segments at the end,
//                                                                      spread by edge id: the counting sort ranks inside a segment)
//                             drk_segment_index_build(key, KK + 1)     edges grouped by pooled pair, ascending edge id inside a pair
//                             drk_compact_segments                     distinct pairs in (row, col) order = coalesce's order
//                             drk_pool_edge_decode                     pooled edge_index from the dense ids
//                             drk_spmm(SUM)                            merged edge attributes, summed in ascending edge id (drk_sparse.cu)
#include <algorithm>

#include "drk_common.cuh"

namespace drk {
namespace pool {

constexpr int kCT = 1024;

// Two launches over chunks of kCT * kCI segment sizes: (1) every CTA counts the non-empty segments of its chunk; (2) every CTA adds up the
// counts of the chunks before its own (a few hundred integers), then scans its chunk tile by tile and writes the kept segments.  The first
// version walked all K segments with ONE CTA: 141 us for the ~370 k dense pair ids of a pooled C2 batch, 0.42 of the 1.0 ms train step of the
// clustered networks (profiles/r02_launches_c4-ginet_v2).
constexpr int kCI = 4;  // tiles per chunk

__global__ void __launch_bounds__(kCT) k_compact_count(const int32_t* __restrict__ ptr, int32_t num_segments, int32_t* __restrict__ chunk_count) {
  const int base = blockIdx.x * kCT * kCI;
  int mine = 0;
#pragma unroll
  for (int i = 0; i < kCI; ++i) {
    const int k = base + i * kCT + threadIdx.x;
    if (k < num_segments) mine += __ldg(ptr + k + 1) > __ldg(ptr + k) ? 1 : 0;
  }
  const int total = __reduce_add_sync(0xffffffffu, mine);
  __shared__ int s_warp[kCT / 32];
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int v = __reduce_add_sync(0xffffffffu, s_warp[threadIdx.x]);
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kCT) k_compact_segments(const int32_t* __restrict__ ptr, int32_t num_segments, const int32_t* __restrict__ perm,
                                                         int64_t* __restrict__ rank, int32_t* __restrict__ ptr_out, int32_t* __restrict__ ids_out,
                                                         int64_t* __restrict__ last_out, int32_t cap, int32_t* __restrict__ count_out,
                                                         int32_t* __restrict__ status, const int32_t* __restrict__ chunk_count) {
  __shared__ int s_warp[kCT / 32 + 1];
  __shared__ int s_base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // kept segments before this chunk, and in total: fixed-order integer sums
  {
    int before = 0, all = 0;
    for (int b = tid; b < (int)gridDim.x; b += kCT) {
      const int c = __ldg(chunk_count + b);
      all += c;
      if (b < (int)blockIdx.x) before += c;
    }
    before = __reduce_add_sync(0xffffffffu, before);
    all = __reduce_add_sync(0xffffffffu, all);
    __shared__ int s_b[kCT / 32], s_a[kCT / 32];
    if (lane == 0) {
      s_b[warp] = before;
      s_a[warp] = all;
    }
    __syncthreads();
    if (warp == 0) {
      const int vb = __reduce_add_sync(0xffffffffu, s_b[lane]), va = __reduce_add_sync(0xffffffffu, s_a[lane]);
      if (lane == 0) {
        s_base[0] = vb;
        s_base[1] = va;
      }
    }
    __syncthreads();
  }
  int carry = s_base[0];
  const int kept = s_base[1];
  const int chunk0 = blockIdx.x * kCT * kCI;
  for (int base = chunk0; base < min(chunk0 + kCT * kCI, num_segments); base += kCT) {
    const int k = base + tid;
    int lo = 0, hi = 0;
    if (k < num_segments) {
      lo = __ldg(ptr + k);
      hi = __ldg(ptr + k + 1);
    }
    const int present = hi > lo ? 1 : 0;
    int inc = present;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;
      if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int pos = carry + s_warp[warp] + inc - present;  // compact index of this segment if it is present
    if (k < num_segments) {
      if (rank != nullptr) rank[k] = present ? (int64_t)pos : (int64_t)-1;
      if (present && pos < cap) {
        ptr_out[pos] = lo;
        if (ids_out != nullptr) ids_out[pos] = k;
        if (last_out != nullptr) last_out[pos] = (int64_t)__ldg(perm + hi - 1);  // stable grouping: the last member has the largest index
      }
    }
    carry += s_warp[32];
    __syncthreads();
  }
  const int total = num_segments > 0 ? __ldg(ptr + num_segments) : 0;
  for (int i = min(kept, cap) + (int)blockIdx.x * kCT + tid; i <= cap; i += (int)gridDim.x * kCT) ptr_out[i] = total;  // trailing (unused) capacity = empty segments
  if (blockIdx.x == 0 && tid == 0) {
    if (count_out != nullptr) *count_out = kept;
    if (kept > cap && status != nullptr) atomicOr(status, DRK_STATUS_INDEX_RANGE);
  }
}

// key[e] = kkptr[g] + (inv[row] - cptr[g]) * C_g + (inv[col] - cptr[g]) with g the graph of the edge's row node; self loops of the
// pooled graph and edges whose endpoints fall outside the graph's cluster range go to the junk segments [junk, junk + DRK_POOL_JUNK_SEGMENTS).
__global__ void __launch_bounds__(256) k_pool_edge_keys(const int64_t* __restrict__ erow, const int64_t* __restrict__ ecol, int64_t num_edges,
                                                       const int64_t* __restrict__ inv, int32_t num_nodes, const int32_t* __restrict__ batch32,
                                                       const int64_t* __restrict__ cptr, const int64_t* __restrict__ kkptr, int32_t num_graphs,
                                                       int64_t junk, int64_t* __restrict__ key, int32_t* __restrict__ status) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += stride) {
    const unsigned long long r = (unsigned long long)ld_stream_i64(erow + e), c = (unsigned long long)ld_stream_i64(ecol + e);
    int64_t k = junk + (e & (DRK_POOL_JUNK_SEGMENTS - 1));
    if (r < (unsigned long long)num_nodes && c < (unsigned long long)num_nodes) {
      const int g = __ldg(batch32 + r);
      if ((unsigned)g < (unsigned)num_graphs) {
        const int64_t c0 = __ldg(cptr + g), cg = __ldg(cptr + g + 1) - c0;
        const int64_t pr = __ldg(inv + r) - c0, pc = __ldg(inv + c) - c0;
        if (pr >= 0 && pr < cg && pc >= 0 && pc < cg) {
          if (pr != pc) k = __ldg(kkptr + g) + pr * cg + pc;
        } else {
          bad = true;  // an edge that joins two graphs, or a cluster id outside its graph's range
        }
      } else {
        bad = true;
      }
    } else {
      bad = true;
    }
    key[e] = k;
  }
  if (bad && status != nullptr) atomicOr(status, DRK_STATUS_CROSS_GRAPH);
}

// pooled edge_index [2, count] from the dense pair ids (ascending = sorted by (row, col), graphs in order)
__global__ void __launch_bounds__(256) k_pool_edge_decode(const int32_t* __restrict__ ids, int32_t cap, const int32_t* __restrict__ count,
                                                         const int64_t* __restrict__ cptr, const int64_t* __restrict__ kkptr, int32_t num_graphs,
                                                         int64_t* __restrict__ out_row, int64_t* __restrict__ out_col) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  if (i >= min(__ldg(count), cap)) {  // capacity beyond the pairs found (the host's bound was not tight): keep the arrays well defined
    out_row[i] = 0;
    out_col[i] = 0;
    return;
  }
  const int64_t k = (int64_t)__ldg(ids + i);
  int lo = 0, hi = num_graphs;  // last g with kkptr[g] <= k
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(kkptr + mid) <= k) lo = mid;
    else hi = mid;
  }
  const int64_t c0 = __ldg(cptr + lo), cg = __ldg(cptr + lo + 1) - c0, local = k - __ldg(kkptr + lo);
  out_row[i] = c0 + local / cg;
  out_col[i] = c0 + local % cg;
}


// ---------------------------------------------------------------------------------------------------------------------------------
// pool_edge for collated batches, one CTA per graph, in shared memory.  The edges of a graph are one slice of edge_index and its pooled
// pairs live in a C_g x C_g block (C_g clusters, a few dozen): count the edges per pair with shared-memory atomics, scan the block
// (offsets of the member lists + compact index of every non-empty pair = its position among the graph's pooled edges, which the
// collate knows in advance: pooled_edge_ptr), drop the members into their lists, and let one thread per pooled pair sort its (short)
// list by edge id and add the attributes in that order -- the same association as the global route (counting sort by dense pair id ->
// drk_compact_segments -> drk_pool_edge_decode -> drk_spmm), which took ~140 us of a 0.6-0.7 ms clustered train step on the C2 batch.
constexpr int kPT = 512;

struct PoolBlockedArgs {
  const int64_t* erow;
  const int64_t* ecol;
  const int32_t* edge_ptr;         // [G + 1]
  const int64_t* inv;              // [N] consecutive cluster id of every node
  const int64_t* cptr;             // [G + 1] first cluster of every graph
  const int32_t* pooled_edge_ptr;  // [G + 1] first pooled edge of every graph
  const float* attr;               // [E, fe] or NULL
  int64_t* out_row;                // [E1]
  int64_t* out_col;                // [E1]
  float* out_attr;                 // [E1, fe] or NULL
  int32_t* status;
  int64_t ld_attr, num_pooled;  // num_pooled = rows the caller allocated
  int32_t num_nodes, fe, cap_clusters, cap_edges;
};

__device__ __forceinline__ int block_excl_scan_512(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < kPT / 32 ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kPT / 32) s_warp[lane] = winc - w;
    if (lane == 31) s_warp[kPT / 32] = winc;
  }
  __syncthreads();
  const int excl = s_warp[warp] + inc - v;
  total = s_warp[kPT / 32];
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(kPT) k_pool_edge_blocked(const PoolBlockedArgs a) {
  extern __shared__ __align__(16) unsigned char pool_smem[];
  __shared__ int s_warp[kPT / 32 + 1];
  const int cc = a.cap_clusters * a.cap_clusters;
  int* s_start = reinterpret_cast<int*>(pool_smem);  // [cc + 1] edges per pair, then the offsets of the member lists
  int* s_pid = s_start + cc + 1;                     // [cc] compact index of the pair among the graph's pooled edges
  int* s_fill = s_pid + cc;                          // [cc]
  int* s_mem = s_fill + cc;                          // [cap_edges] members of every pair (graph-local edge ids)
  unsigned short* s_key = reinterpret_cast<unsigned short*>(s_mem + a.cap_edges);  // [cap_edges] pair of every edge, 0xffff = dropped
  const int g = blockIdx.x, tid = threadIdx.x;
  const int e0 = __ldg(a.edge_ptr + g), ne = __ldg(a.edge_ptr + g + 1) - e0;
  const int64_t c0 = __ldg(a.cptr + g);
  const int C = (int)(__ldg(a.cptr + g + 1) - c0);
  const int o0 = __ldg(a.pooled_edge_ptr + g), want = __ldg(a.pooled_edge_ptr + g + 1) - o0;
  if (C < 0 || C > a.cap_clusters || ne < 0 || ne > a.cap_edges || C * C >= 0xffff) {
    if (tid == 0 && a.status != nullptr) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
    return;
  }
  const int kk = C * C;
  for (int k = tid; k < kk; k += kPT) {
    s_start[k] = 0;
    s_fill[k] = 0;
  }
  __syncthreads();
  bool bad = false;
  for (int eb = tid; eb < ne; eb += 4 * kPT) {  // 4 edges per thread and trip: the endpoint -> cluster lookups are two dependent L2 round trips
    unsigned long long r[4], c[4];
    int64_t pr[4], pc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = eb + u * kPT;
      r[u] = e < ne ? (unsigned long long)ld_stream_i64(a.erow + e0 + e) : 0ull;
      c[u] = e < ne ? (unsigned long long)ld_stream_i64(a.ecol + e0 + e) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = r[u] < (unsigned long long)a.num_nodes && c[u] < (unsigned long long)a.num_nodes;
      pr[u] = in ? __ldg(a.inv + r[u]) - c0 : -1;
      pc[u] = in ? __ldg(a.inv + c[u]) - c0 : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = eb + u * kPT;
      if (e >= ne) continue;
      unsigned short key = 0xffffu;
      if (pr[u] >= 0 && pr[u] < C && pc[u] >= 0 && pc[u] < C) {
        if (pr[u] != pc[u]) {
          key = (unsigned short)(pr[u] * C + pc[u]);
          atomicAdd(&s_start[key], 1);
        }
      } else {
        bad = true;  // an endpoint outside the batch, an edge that joins two graphs, or a cluster id outside its graph's range
      }
      s_key[e] = key;
    }
  }
  __syncthreads();
  // offsets of the member lists and compact index of the non-empty pairs: one scan over the C x C block, kPT pairs at a time
  int carry_e = 0, carry_p = 0;
  for (int base = 0; base < kk; base += kPT) {
    const int k = base + tid;
    const int n = k < kk ? s_start[k] : 0;
    int tot_e, tot_p;
    const int ex_e = block_excl_scan_512(n, s_warp, tot_e);
    const int ex_p = block_excl_scan_512(n > 0 ? 1 : 0, s_warp, tot_p);
    if (k < kk) {
      s_start[k] = carry_e + ex_e;
      s_pid[k] = n > 0 ? carry_p + ex_p : -1;
    }
    carry_e += tot_e;
    carry_p += tot_p;
  }
  if (tid == 0) s_start[kk] = carry_e;
  const bool miscount = carry_p != want || (int64_t)o0 + want > a.num_pooled;  // the collate's count of distinct pooled edges does not hold
  __syncthreads();
  for (int e = tid; e < ne; e += kPT) {
    const unsigned short key = s_key[e];
    if (key != 0xffffu) s_mem[s_start[key] + atomicAdd(&s_fill[key], 1)] = e;
  }
  __syncthreads();
  for (int k = tid; k < kk; k += kPT) {
    const int pid = s_pid[k];
    if (pid < 0 || pid >= want || (int64_t)o0 + pid >= a.num_pooled) continue;
    const int lo = s_start[k], hi = s_start[k + 1];
    for (int i = lo + 1; i < hi; ++i) {  // the atomics dropped the members in arbitrary order: ascending edge id (short lists)
      const int v = s_mem[i];
      int j = i - 1;
      while (j >= lo && s_mem[j] > v) {
        s_mem[j + 1] = s_mem[j];
        --j;
      }
      s_mem[j + 1] = v;
    }
    const int64_t o = (int64_t)o0 + pid;
    a.out_row[o] = c0 + k / C;
    a.out_col[o] = c0 + k % C;
    if (a.out_attr != nullptr) {
      for (int f = 0; f < a.fe; ++f) {
        float sum = 0.f;
        for (int i0 = lo; i0 < hi; i0 += 8) {  // 8 gathers in flight, added in ascending edge id
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = i0 + u < hi ? __ldg(a.attr + (int64_t)(e0 + s_mem[i0 + u]) * a.ld_attr + f) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (i0 + u < hi) sum += v[u];
        }
        a.out_attr[o * a.fe + f] = sum;
      }
    }
  }
  if (a.status != nullptr) {
    if (bad) atomicOr(a.status, DRK_STATUS_CROSS_GRAPH);
    if (miscount && tid == 0) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// consecutive_cluster for collated batches, one CTA per graph: the ids of a graph's nodes (globally unique after get_preloaded_cluster,
// i.e. a contiguous range per graph) are relabelled 0..C_g-1 in ascending id order and offset by the graph's first new id, which the
// collate knows (cluster_ptr).  One launch instead of the eight of drk_segment_index_build + drk_compact_segments + the rank gather:
//   inv [N]        new id of every node                      (PyG: the `inverse` of torch.unique)
//   last [C]       largest node index of every cluster       (PyG's `perm` on CPU: last writer wins)
//   ptr_c [C + 1], perm [N]   nodes grouped by new id, ascending node index inside a cluster (the segment plan of scatter_max / _mean)
constexpr int kCBT = 256;

struct ConsecutiveArgs {
  const int64_t* cluster;   // [N]
  const int32_t* node_ptr;  // [G + 1]
  const int64_t* cptr;      // [G + 1] first new id of every graph
  int64_t* inv;
  int64_t* last;
  int32_t* ptr_c;
  int32_t* perm;
  int32_t* status;
  int32_t num_graphs, num_nodes, cap_nodes, cap_ids, capacity;  // capacity = clusters the caller allocated
};

__global__ void __launch_bounds__(kCBT) k_consecutive_blocked(const ConsecutiveArgs a) {
  extern __shared__ __align__(16) unsigned char pool_smem[];
  __shared__ long long s_min[kCBT / 32];
  __shared__ int s_scan[kCBT / 32 + 1];
  int* s_cnt = reinterpret_cast<int*>(pool_smem);  // [cap_ids] nodes per old id, then the first perm slot of the id
  int* s_rank = s_cnt + a.cap_ids;                 // [cap_ids] new (graph-local) id, -1 for absent ids
  int* s_key = s_rank + a.cap_ids;                 // [cap_nodes] old id - smallest id of the graph
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int node0 = __ldg(a.node_ptr + g), n = __ldg(a.node_ptr + g + 1) - node0;
  const int64_t c0 = __ldg(a.cptr + g);
  const int want = (int)(__ldg(a.cptr + g + 1) - c0);
  if (g == a.num_graphs - 1 && tid == 0) {
    const int64_t total = __ldg(a.cptr + a.num_graphs);
    a.ptr_c[min(total, (int64_t)a.capacity)] = a.num_nodes;
    if (total != a.capacity && a.status != nullptr) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);  // stale sizes
  }
  if (n < 0 || n > a.cap_nodes) {
    if (tid == 0 && a.status != nullptr) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
    return;
  }
  if (n == 0) return;
  long long lo = 0x7fffffffffffffffll;
  for (int i = tid; i < n; i += kCBT) lo = min(lo, (long long)ld_stream_i64(a.cluster + node0 + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  if (lane == 0) s_min[warp] = lo;
  for (int k = tid; k < a.cap_ids; k += kCBT) s_cnt[k] = 0;
  __syncthreads();
  lo = s_min[0];
#pragma unroll
  for (int w = 1; w < kCBT / 32; ++w) lo = min(lo, s_min[w]);
  bool bad = false;
  for (int i = tid; i < n; i += kCBT) {
    const long long k = (long long)ld_stream_i64(a.cluster + node0 + i) - lo;
    int key = -1;
    if (k < a.cap_ids) {
      key = (int)k;
      atomicAdd(&s_cnt[key], 1);
    } else {
      bad = true;  // the ids of this graph span more than the id bound recorded for the batch
    }
    s_key[i] = key;
    a.perm[node0 + i] = node0 + i;  // a well-defined entry even for nodes a flagged batch leaves without a slot
  }
  __syncthreads();
  // new ids = number of present ids below; perm slots = number of nodes with a smaller id
  int carry_n = 0, carry_c = 0;
  for (int base = 0; base < a.cap_ids; base += kCBT) {
    const int k = base + tid;
    const int c = k < a.cap_ids ? s_cnt[k] : 0;
    int incn = c, incc = c > 0 ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tn = __shfl_up_sync(0xffffffffu, incn, o), tc = __shfl_up_sync(0xffffffffu, incc, o);
      if (lane >= o) {
        incn += tn;
        incc += tc;
      }
    }
    if (lane == 31) s_scan[warp] = (incn << 12) | incc;  // n <= cap_nodes < 2^19, ids per tile <= 256
    __syncthreads();
    int wn = 0, wc = 0, tn_all = 0, tc_all = 0;
    for (int w = 0; w < kCBT / 32; ++w) {
      const int v = s_scan[w];
      if (w < warp) {
        wn += v >> 12;
        wc += v & 0xfff;
      }
      tn_all += v >> 12;
      tc_all += v & 0xfff;
    }
    __syncthreads();
    if (k < a.cap_ids) {
      s_cnt[k] = carry_n + wn + incn - c;
      s_rank[k] = c > 0 ? carry_c + wc + incc - 1 : -1;
    }
    carry_n += tn_all;
    carry_c += tc_all;
  }
  const bool miscount = carry_c != want;
  __syncthreads();
  for (int i = tid; i < n; i += kCBT) {
    const int key = s_key[i];
    const int r = key >= 0 ? s_rank[key] : -1;
    a.inv[node0 + i] = (r >= 0 && r < want && c0 + r < a.capacity) ? c0 + r : min(c0, (int64_t)max(a.capacity - 1, 0));  // (a flagged batch still gets in-range ids)
  }
  for (int k = tid; k < a.cap_ids; k += kCBT) {
    const int r = s_rank[k];
    if (r < 0) continue;
    const bool keep = r < want && c0 + r < a.capacity;
    int pos = node0 + s_cnt[k], last = node0;
    if (keep) a.ptr_c[c0 + r] = pos;
    for (int i = 0; i < n; ++i) {  // members in ascending node index (n is a few hundred, the ids a few dozen)
      if (s_key[i] == k) {
        a.perm[pos++] = node0 + i;
        last = node0 + i;
      }
    }
    if (keep) a.last[c0 + r] = last;
  }
  if (a.status != nullptr) {
    if (bad) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
    if (miscount && tid == 0) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);  // the collate's count of distinct ids does not hold for this graph
  }
}

static size_t pool_blocked_smem(int cap_clusters, int cap_edges) {
  const size_t cc = (size_t)cap_clusters * cap_clusters;
  return (3 * cc + 1 + (size_t)cap_edges) * sizeof(int) + (size_t)cap_edges * sizeof(unsigned short) + 16;
}

}  // namespace pool
}  // namespace drk

extern "C" {

size_t drk_compact_segments_workspace_bytes(int32_t num_segments) {
  using namespace drk;
  const int chunks = std::max(1, ceil_div(std::max(num_segments, 0), pool::kCT * pool::kCI));
  return (size_t)chunks * sizeof(int32_t);
}

int drk_compact_segments(const int32_t* ptr, int32_t num_segments, const int32_t* perm, int64_t* rank, int32_t* ptr_out, int32_t* ids_out,
                         int64_t* last_out, int32_t capacity, int32_t* count_out, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_segments >= 0 && capacity >= 0, DRK_EINVAL, "compact segments: negative size");
  DRK_REQUIRE(ptr && ptr_out, DRK_EINVAL, "compact segments: null pointer");
  DRK_REQUIRE(last_out == nullptr || perm != nullptr, DRK_EINVAL, "compact segments: last_out needs perm");
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= drk_compact_segments_workspace_bytes(num_segments), DRK_EWORKSPACE, "compact segments: workspace too small");
  const int chunks = std::max(1, ceil_div(num_segments, pool::kCT * pool::kCI));
  int32_t* chunk_count = static_cast<int32_t*>(workspace);
  pool::k_compact_count<<<chunks, pool::kCT, 0, as_stream(stream)>>>(ptr, num_segments, chunk_count);
  pool::k_compact_segments<<<chunks, pool::kCT, 0, as_stream(stream)>>>(ptr, num_segments, perm, rank, ptr_out, ids_out, last_out, capacity, count_out, status,
                                                                     chunk_count);
  return finish_launch("compact segments", 2);
}

int drk_pool_edge_keys(const int64_t* edge_index, int64_t num_edges, const int64_t* inv, int32_t num_nodes, const int32_t* batch32,
                       const int64_t* cluster_ptr, const int64_t* pair_ptr, int32_t num_graphs, int64_t junk_key, int64_t* key, int32_t* status,
                       void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_edges >= 0 && num_nodes >= 0 && num_graphs >= 0 && junk_key >= 0, DRK_EINVAL, "pool edge keys: negative size");
  if (num_edges == 0) return DRK_OK;
  DRK_REQUIRE(edge_index && inv && batch32 && cluster_ptr && pair_ptr && key, DRK_EINVAL, "pool edge keys: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(num_edges, 256), (int64_t)kNumSM * 8);
  pool::k_pool_edge_keys<<<grid, 256, 0, as_stream(stream)>>>(edge_index, edge_index + num_edges, num_edges, inv, num_nodes, batch32, cluster_ptr, pair_ptr,
                                                             num_graphs, junk_key, key, status);
  return finish_launch("pool edge keys");
}

int drk_pool_edge_decode(const int32_t* ids, int32_t capacity, const int32_t* count, const int64_t* cluster_ptr, const int64_t* pair_ptr, int32_t num_graphs,
                         int64_t* edge_index_out, void* stream) {
  using namespace drk;
  DRK_REQUIRE(capacity >= 0 && num_graphs >= 0, DRK_EINVAL, "pool edge decode: negative size");
  if (capacity == 0) return DRK_OK;
  DRK_REQUIRE(ids && count && cluster_ptr && pair_ptr && edge_index_out, DRK_EINVAL, "pool edge decode: null pointer");
  pool::k_pool_edge_decode<<<ceil_div(capacity, 256), 256, 0, as_stream(stream)>>>(ids, capacity, count, cluster_ptr, pair_ptr, num_graphs, edge_index_out,
                                                                                 edge_index_out + capacity);
  return finish_launch("pool edge decode");
}

int drk_pool_edge_blocked_supported(int32_t max_graph_clusters, int32_t max_graph_edges) {
  using namespace drk;
  if (max_graph_clusters < 1 || max_graph_edges < 0 || max_graph_clusters > 255) return 0;
  return pool::pool_blocked_smem(max_graph_clusters, max_graph_edges) <= (size_t)200 * 1024 ? 1 : 0;
}

int drk_pool_edge_blocked(const int64_t* edge_index, int64_t num_edges, const int32_t* edge_ptr, const int64_t* inv, int32_t num_nodes,
                          const int64_t* cluster_ptr, const int32_t* pooled_edge_ptr, int32_t num_graphs, int32_t max_graph_clusters, int32_t max_graph_edges,
                          const float* edge_attr, int64_t ld_attr, int32_t num_edge_features, int64_t* pooled_index, int64_t num_pooled, float* pooled_attr,
                          int32_t* status, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_edges >= 0 && num_nodes >= 0 && num_graphs >= 0 && num_pooled >= 0 && num_edge_features >= 0, DRK_EINVAL, "pool edge blocked: negative size");
  if (num_graphs == 0) return DRK_OK;
  DRK_REQUIRE(drk_pool_edge_blocked_supported(max_graph_clusters, max_graph_edges), DRK_EUNSUPPORTED,
              "pool edge blocked: %d clusters / %d edges per graph do not fit one CTA (use the global route)", max_graph_clusters, max_graph_edges);
  DRK_REQUIRE(edge_ptr && inv && cluster_ptr && pooled_edge_ptr && (num_edges == 0 || edge_index) && (num_pooled == 0 || pooled_index), DRK_EINVAL,
              "pool edge blocked: null pointer");
  DRK_REQUIRE(pooled_attr == nullptr || edge_attr != nullptr, DRK_EINVAL, "pool edge blocked: pooled_attr needs edge_attr");
  pool::PoolBlockedArgs a{edge_index, edge_index + num_edges, edge_ptr, inv, cluster_ptr, pooled_edge_ptr, edge_attr, pooled_index, pooled_index + num_pooled,
                          pooled_attr, status, ld_attr, num_pooled, num_nodes, num_edge_features, max_graph_clusters, max_graph_edges};
  const size_t smem = pool::pool_blocked_smem(max_graph_clusters, max_graph_edges);
  cudaError_t e = cudaFuncSetAttribute(pool::k_pool_edge_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "pool edge blocked: smem opt-in: %s", cudaGetErrorString(e));
  pool::k_pool_edge_blocked<<<num_graphs, pool::kPT, smem, as_stream(stream)>>>(a);
  return finish_launch("pool edge blocked");
}

int drk_consecutive_blocked_supported(int32_t max_graph_nodes, int32_t max_graph_ids) {
  if (max_graph_nodes < 1 || max_graph_ids < 1 || max_graph_nodes >= (1 << 19)) return 0;
  return ((size_t)2 * max_graph_ids + (size_t)max_graph_nodes) * sizeof(int) + 16 <= (size_t)200 * 1024 ? 1 : 0;
}

int drk_consecutive_blocked(const int64_t* cluster, int32_t num_nodes, const int32_t* node_ptr, const int64_t* cluster_ptr, int32_t num_graphs,
                            int32_t max_graph_nodes, int32_t max_graph_ids, int32_t capacity, int64_t* inv, int64_t* last, int32_t* ptr_c, int32_t* perm,
                            int32_t* status, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_nodes >= 0 && num_graphs >= 0 && capacity >= 0, DRK_EINVAL, "consecutive blocked: negative size");
  if (num_graphs == 0) return DRK_OK;
  DRK_REQUIRE(drk_consecutive_blocked_supported(max_graph_nodes, max_graph_ids), DRK_EUNSUPPORTED,
              "consecutive blocked: %d nodes / %d ids per graph do not fit one CTA (use the global route)", max_graph_nodes, max_graph_ids);
  DRK_REQUIRE(node_ptr && cluster_ptr && inv && last && ptr_c && perm && (cluster || num_nodes == 0), DRK_EINVAL, "consecutive blocked: null pointer");
  pool::ConsecutiveArgs a{cluster, node_ptr, cluster_ptr, inv, last, ptr_c, perm, status, num_graphs, num_nodes, max_graph_nodes, max_graph_ids, capacity};
  const size_t smem = ((size_t)2 * max_graph_ids + (size_t)max_graph_nodes) * sizeof(int) + 16;
  cudaError_t e = cudaFuncSetAttribute(pool::k_consecutive_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "consecutive blocked: smem opt-in: %s", cudaGetErrorString(e));
  pool::k_consecutive_blocked<<<num_graphs, pool::kCBT, smem, as_stream(stream)>>>(a);
  return finish_launch("consecutive blocked");
}

}  // extern "C"
