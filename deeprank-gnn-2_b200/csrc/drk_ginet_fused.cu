// Fused per-graph GINet kernels: the whole two-branch convolution stack of ginet_nocluster.GINet
// (reference ginet_nocluster.py:88-106) for ONE graph per CTA, with every intermediate tile in shared memory.
//
//   forward :  P = x [W1;W1e]^T -> H1 = relu(A P) -> A2 = A H1 -> H2 = relu([A2a W2^T | A2b W2e^T]) -> G[g] = mean_i H2[i]
//   backward:  dZ2 = dG[g]/n * (Z2 > 0) (Z2 recomputed from A2) ; dW2 += dZ2^T A2 ; dA2 = dZ2 W2
//              dZ1 = (A^T dA2) * (H1 > 0) ; Q = A^T dZ1 ; dW1 += Q^T x
//
// Why per graph: a batch is a block-diagonal union of ~300-node graphs, so all gathers of a graph hit a 300 x 32
// fp32 tile (38 KB).  Staged once in shared memory, the four aggregations of a train step read HBM for the
// index stream only; the unfused path re-reads/writes an [N,32] tensor around every one of ~13 launches.
// HBM traffic per graph, forward: x (4 n F) + colidx (2 x 4 e) + rowptr + saved H1, A2 (2 x 128 n) + 256 B out.
//
// Feature widths are the architecture's (conv1: F -> 16, conv2: 16 -> 32, two branches => 32 / 64 stacked); F (<= 64)
// and the graph sizes are runtime.  Graphs larger than the shared-memory budget make the launch wrapper
// return DRK_EUNSUPPORTED and the host falls back to the unfused kernels (same results).
// Accumulation order: edges of a destination in CSR order (the reference's scatter_add_ order), fp32, no atomics.
#include <algorithm>

#include "drk_common.cuh"

namespace drk {

constexpr int kS1 = 32;   // stacked conv1 outputs (2 x 16)
constexpr int kF1 = 16;   // conv1 outputs per branch = conv2 inputs per branch
constexpr int kS2 = 64;   // stacked conv2 outputs (2 x 32)
constexpr int kF2 = 32;
constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;

__host__ __device__ inline int fused_kp(int fi) {  // padded smem row stride of the x tile / W1 rows (see drk_dense.cu)
  int kp = (fi + 3) / 4 * 4;
  if (((kp / 4) & 1) == 0) kp += 4;
  return kp;
}

struct GinetFwdArgs {
  const float* x;
  int64_t ldx;
  int32_t fi;
  const int32_t* graph_ptr;
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* w1s;  // [32, fi]: conv1.fc.weight over conv1_ext.fc.weight
  const float* w2a;  // [32, 16] conv2.fc.weight
  const float* w2b;  // [32, 16] conv2_ext.fc.weight
  float* h1s;        // [N, 32] saved for backward (may be NULL)
  float* a2s;        // [N, 32] saved for backward (may be NULL)
  float* g;          // [B, 64]
  int32_t* status;
  int32_t num_graphs;
  int32_t rows_cap;  // shared-memory capacity in rows
  int32_t x_vec;     // 4 / 2 / 1
  int32_t idx_cap;   // shared-memory capacity in edge slots
};

// Stage the graph's slice of a CSR/CSC into shared memory: s_ptr[i] = ptr[node0+i] - ptr[node0] (n+1 entries) and, for every
// slot, the BYTE offset of the gathered row inside a [rows][32] fp32 tile.  Slots whose endpoint lies outside the graph
// (cannot happen for a collated batch) point at the all-zero row `rows_cap` and raise DRK_STATUS_CROSS_GRAPH, so the gather
// loop itself needs no range check.  Returns false (uniformly) if the graph has more edges than the staging buffer.
__device__ __forceinline__ bool fused_stage_index(int32_t* __restrict__ s_ptr, int32_t* __restrict__ s_off, const int32_t* __restrict__ ptr,
                                                  const int32_t* __restrict__ idx, int node0, int n, int idx_cap, int rows_cap,
                                                  int32_t* status) {
  const int e0 = __ldg(ptr + node0);
  const int ne = __ldg(ptr + node0 + n) - e0;
  for (int i = threadIdx.x; i <= n; i += kFusedThreads) s_ptr[i] = __ldg(ptr + node0 + i) - e0;
  if (ne > idx_cap) return false;
  bool bad = false;
  for (int e = threadIdx.x; e < ne; e += 4 * kFusedThreads) {
    int raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = (e + u * kFusedThreads < ne) ? ld_stream_i32(idx + e0 + e + u * kFusedThreads) : node0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (e + u * kFusedThreads < ne) {
        const int local = raw[u] - node0;
        const bool ok = (unsigned)local < (unsigned)n;
        bad |= !ok;
        s_off[e + u * kFusedThreads] = (ok ? local : rows_cap) * (kS1 * 4);
      }
    }
  }
  if (bad && status != nullptr) atomicOr(status, DRK_STATUS_CROSS_GRAPH);
  return true;
}

// One aggregation over the graph's rows: dst[i] = epi( sum_{s in row i} src[slot s] ), src/dst tiles in smem ([rows][32] floats,
// src has an all-zero row at index rows_cap).  8 lanes per row, 4 rows per warp, warp-uniform trip count, CSR order.
// MODE 0: relu, MODE 1: none, MODE 2: multiply by (mask_global[row] > 0).
// STAGED: slots come from shared memory (byte offsets, padding slots read the zero row -> no predicates in the inner loop);
// otherwise they are streamed from global memory (graphs with more edges than the staging buffer).
template <int MODE, bool STAGED>
__device__ __forceinline__ void fused_aggregate(const float* __restrict__ s_src, float* __restrict__ s_dst, const int32_t* __restrict__ s_ptr,
                                                const int32_t* __restrict__ s_off, const int32_t* __restrict__ g_ptr,
                                                const int32_t* __restrict__ g_idx, int node0, int n, int rows_cap,
                                                float* __restrict__ g_out, const float* __restrict__ g_mask) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int sub = lane >> 3, sl = lane & 7;
  const int group_base = sub * 8;
  const int zero_off = rows_cap * (kS1 * 4);
  const char* lane_base = reinterpret_cast<const char*>(s_src) + sl * 16;
  const int e0 = STAGED ? 0 : __ldg(g_ptr + node0);
  for (int rw = warp * 4; rw < n; rw += kFusedWarps * 4) {
    const int r = rw + sub;
    const bool row_ok = r < n;
    int beg = 0, len = 0;
    if (row_ok) {
      beg = s_ptr[r];
      len = s_ptr[r + 1] - beg;
    }
    int max_len = len;
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 16));
    max_len = max(max_len, __shfl_xor_sync(kFull, max_len, 8));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int off = 0; off < max_len; off += 8) {
      int my_off = zero_off;
      if (off + sl < len) {
        if (STAGED) {
          my_off = s_off[beg + off + sl];
        } else {
          const int local = ld_stream_i32(g_idx + e0 + beg + off + sl) - node0;
          my_off = ((unsigned)local < (unsigned)n ? local : rows_cap) * (kS1 * 4);
        }
      }
      // shared memory delivers one 128 B row per cycle per SM and this loop is bound by exactly that: padding slots
      // (rows shorter than the longest of the warp's four) are predicated off rather than pointed at the zero row
      const int rem = len - off;
#pragma unroll
      for (int j0 = 0; j0 < 8; j0 += 4) {  // four rows in flight per lane (16 registers)
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = __shfl_sync(kFull, my_off, group_base + j0 + j);
          if (j0 + j < rem) v[j] = *reinterpret_cast<const float4*>(lane_base + o);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j0 + j < rem) {
            acc.x += v[j].x;
            acc.y += v[j].y;
            acc.z += v[j].z;
            acc.w += v[j].w;
          }
        }
      }
    }
    if (!row_ok) continue;
    if (MODE == 0) {
      acc.x = acc.x < 0.f ? 0.f : acc.x;
      acc.y = acc.y < 0.f ? 0.f : acc.y;
      acc.z = acc.z < 0.f ? 0.f : acc.z;
      acc.w = acc.w < 0.f ? 0.f : acc.w;
    } else if (MODE == 2) {
      const float4 m = ld_stream_f4(g_mask + (size_t)(node0 + r) * kS1 + sl * 4);
      acc.x = m.x <= 0.f ? 0.f : acc.x;
      acc.y = m.y <= 0.f ? 0.f : acc.y;
      acc.z = m.z <= 0.f ? 0.f : acc.z;
      acc.w = m.w <= 0.f ? 0.f : acc.w;
    }
    *reinterpret_cast<float4*>(s_dst + r * kS1 + sl * 4) = acc;
    if (g_out != nullptr) *reinterpret_cast<float4*>(g_out + (size_t)(node0 + r) * kS1 + sl * 4) = acc;
  }
}

template <int MODE>
__device__ __forceinline__ void fused_aggregate_any(bool staged, const float* s_src, float* s_dst, const int32_t* s_ptr, const int32_t* s_off,
                                                    const int32_t* g_ptr, const int32_t* g_idx, int node0, int n, int rows_cap, float* g_out,
                                                    const float* g_mask) {
  if (staged) fused_aggregate<MODE, true>(s_src, s_dst, s_ptr, s_off, g_ptr, g_idx, node0, n, rows_cap, g_out, g_mask);
  else fused_aggregate<MODE, false>(s_src, s_dst, s_ptr, s_off, g_ptr, g_idx, node0, n, rows_cap, g_out, g_mask);
}

// copy rows [node0, node0+n) of a row-major global matrix (ld elements, width fi) into smem with row stride kp,
// zero-filling the padding columns; asynchronous (cp.async), caller commits/waits.
__device__ __forceinline__ void fused_stage_rows(float* __restrict__ s_dst, const float* __restrict__ src, int64_t ld, int fi, int kp,
                                                 int node0, int n, int vec) {
  const int nv = fi / vec;
  for (int e = threadIdx.x; e < n * nv; e += kFusedThreads) {
    const int r = e / nv;
    const int v = e - r * nv;
    const float* s = src + (int64_t)(node0 + r) * ld + v * vec;
    float* d = s_dst + r * kp + v * vec;
    if (vec == 4) cp_async<16>(d, s, true);
    else if (vec == 2) cp_async<8>(d, s, true);
    else cp_async<4>(d, s, true);
  }
  const int pv = kp - nv * vec;
  for (int e = threadIdx.x; e < n * pv; e += kFusedThreads) {
    const int r = e / pv;
    s_dst[r * kp + nv * vec + (e - r * pv)] = 0.f;
  }
}

__global__ void __launch_bounds__(kFusedThreads, 1) k_ginet_fused_fwd(const GinetFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int kp = fused_kp(a.fi);
  // smem carve-up (floats): W1 [32][kp] | W2 [2][32][16] | colsum scratch [8][64] | tile A [cap][max(kp,32)] + zero row |
  //                         tile B [cap+1][32] (row cap = zeros) | row pointers [cap+1] | slot offsets [idx_cap]
  float* sW1 = smem;
  float* sW2 = sW1 + kS1 * kp;
  float* sRed = sW2 + 2 * kF2 * kF1;
  float* sA = sRed + 8 * kS2;                    // x tile, later H1
  const int wa = kp > kS1 ? kp : kS1;
  float* sB = sA + (size_t)a.rows_cap * wa + kS1;  // P, later A2
  int32_t* sPtr = reinterpret_cast<int32_t*>(sB + (size_t)(a.rows_cap + 1) * kS1);
  int32_t* sOff = sPtr + a.rows_cap + 4;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x < kS1) sB[a.rows_cap * kS1 + threadIdx.x] = 0.f;  // tile B's zero row is never overwritten

  // weights once per CTA.  W1 rows permuted for conflict-free float4 reads: logical m = 4*cg + t + 16*jj -> row cg + 4*t + 16*jj
  for (int e = threadIdx.x; e < kS1 * kp; e += kFusedThreads) {
    const int m = e / kp, k = e - m * kp;
    const int srow = ((m >> 2) & 3) + 4 * (m & 3) + (m & ~15);
    sW1[srow * kp + k] = k < a.fi ? __ldg(a.w1s + (size_t)m * a.fi + k) : 0.f;
  }
  for (int e = threadIdx.x; e < kF2 * kF1; e += kFusedThreads) {
    sW2[e] = __ldg(a.w2a + e);
    sW2[kF2 * kF1 + e] = __ldg(a.w2b + e);
  }

  for (int g = blockIdx.x; g < a.num_graphs; g += gridDim.x) {
    const int node0 = __ldg(a.graph_ptr + g);
    const int n = __ldg(a.graph_ptr + g + 1) - node0;
    __syncthreads();  // previous graph finished with the tiles; weights visible
    if (n > a.rows_cap) {  // cannot happen when the host sized the launch from the true maximum; never write out of bounds
      if (threadIdx.x == 0 && a.status != nullptr) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
      continue;
    }
    // ---- stage x rows (async) and the graph's CSR slice
    fused_stage_rows(sA, a.x, a.ldx, a.fi, kp, node0, n, a.x_vec);
    cp_async_commit();
    const bool staged = fused_stage_index(sPtr, sOff, a.rowptr, a.colidx, node0, n, a.idx_cap, a.rows_cap, a.status);
    cp_async_wait<0>();
    __syncthreads();

    // ---- P = x W1s^T : each warp 32 rows x 32 cols (4 rows x 8 cols per lane), as in k_node_linear<8>
    {
      const int cg = lane & 3, rg = lane >> 2;
      for (int tile = warp * 32; tile < n; tile += kFusedWarps * 32) {
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
        const float* a_base = sA + (tile + rg) * kp;
        const float* w_base = sW1 + cg * kp;
        for (int k4 = 0; k4 < kp; k4 += 4) {
          float4 av[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // rows beyond n read stale smem (finite or not): their results are never stored
            const int rr = min(tile + rg + 8 * j, a.rows_cap - 1) - (tile + rg);
            av[j] = *reinterpret_cast<const float4*>(a_base + rr * kp + k4);
          }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 bv = *reinterpret_cast<const float4*>(w_base + (4 * t + 16 * jj) * kp + k4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float s = acc[j][jj * 4 + t];
                s = fmaf(av[j].x, bv.x, s);
                s = fmaf(av[j].y, bv.y, s);
                s = fmaf(av[j].z, bv.z, s);
                s = fmaf(av[j].w, bv.w, s);
                acc[j][jj * 4 + t] = s;
              }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tile + rg + 8 * j;
          if (r < n) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
              *reinterpret_cast<float4*>(sB + r * kS1 + 4 * cg + 16 * jj) = make_float4(acc[j][jj * 4], acc[j][jj * 4 + 1], acc[j][jj * 4 + 2], acc[j][jj * 4 + 3]);
          }
        }
      }
    }
    __syncthreads();
    // ---- H1 = relu(A P)  (tile A is free: x is consumed; re-create its zero row, which the x tile overlapped)
    if (threadIdx.x < kS1) sA[a.rows_cap * kS1 + threadIdx.x] = 0.f;
    fused_aggregate_any<0>(staged, sB, sA, sPtr, sOff, a.rowptr, a.colidx, node0, n, a.rows_cap, a.h1s, nullptr);
    __syncthreads();
    // ---- A2 = A H1
    fused_aggregate_any<1>(staged, sA, sB, sPtr, sOff, a.rowptr, a.colidx, node0, n, a.rows_cap, a.a2s, nullptr);
    __syncthreads();
    // ---- H2 = relu(A2 W2^T) per branch, column sums for the readout.  thread -> column c (0..63), row lane rl (0..7)
    {
      const int c = threadIdx.x & 63;
      const int rl = threadIdx.x >> 6;
      const int branch = c >> 5;
      float w[kF1];
#pragma unroll
      for (int k = 0; k < kF1; ++k) w[k] = sW2[branch * kF2 * kF1 + (c & 31) * kF1 + k];
      float colsum = 0.f;
      for (int r = rl; r < n; r += 32) {  // four independent rows per trip (ILP); rows are still added in ascending order
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = min(r + 8 * u, a.rows_cap - 1);
          const float* arow = sB + rr * kS1 + branch * kF1;
#pragma unroll
          for (int k4 = 0; k4 < kF1; k4 += 4) {
            const float4 v = *reinterpret_cast<const float4*>(arow + k4);
            z[u] = fmaf(v.x, w[k4], z[u]);
            z[u] = fmaf(v.y, w[k4 + 1], z[u]);
            z[u] = fmaf(v.z, w[k4 + 2], z[u]);
            z[u] = fmaf(v.w, w[k4 + 3], z[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (r + 8 * u < n) colsum += z[u] < 0.f ? 0.f : z[u];
      }
      sRed[rl * kS2 + c] = colsum;
    }
    __syncthreads();
    if (threadIdx.x < kS2) {
      float s = 0.f;
#pragma unroll
      for (int rl = 0; rl < 8; ++rl) s += sRed[rl * kS2 + threadIdx.x];
      a.g[(size_t)g * kS2 + threadIdx.x] = s / fmaxf((float)n, 1.f);  // scatter_mean: count clamped to >= 1
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
struct GinetBwdArgs {
  const float* x;
  int64_t ldx;
  int32_t fi;
  const int32_t* graph_ptr;
  const int32_t* colptr;  // CSC (by source): A^T
  const int32_t* rowidx;
  const float* w2a;
  const float* w2b;
  const float* h1s;  // saved [N,32]
  const float* a2s;  // saved [N,32]
  const float* dg;   // [B,64]
  float* partial;    // [grid][kS1*64 + kS2*kF1]: per-CTA dW1s (k padded to 64) then dW2 stacked [64][16]
  int32_t* status;
  int32_t num_graphs;
  int32_t rows_cap;
  int32_t x_vec;
  int32_t idx_cap;
};

constexpr int kBwdBlockRows = 64;          // rows per dZ2 block
constexpr int kDzStride = kS2 + 4;         // 68: consecutive rows start 4 banks apart -> conflict-free float4 reads
constexpr int kWtStride = kF2 + 4;         // 36: same for the rows of W2^T
constexpr int kPartialW1 = kS1 * 64;       // dW1s partial, k padded to 64
constexpr int kPartialW2 = kS2 * kF1;      // dW2 | dW2e partial
constexpr int kPartial = kPartialW1 + kPartialW2;

__global__ void __launch_bounds__(kFusedThreads, 1) k_ginet_fused_bwd(const GinetBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int kp = fused_kp(a.fi);
  // smem (floats): W2 [2][32][16] | W2T [2][16][32] | dZ2 block [64][64] | tile A [cap+1][32] | tile B [cap+1][32] (row cap = zeros)
  //                | x tile [cap][kp] | column pointers [cap+1] | slot offsets [idx_cap]
  float* sW2 = smem;
  float* sW2T = sW2 + 2 * kF2 * kF1;
  float* sDZ = sW2T + 2 * kF1 * kWtStride;
  float* sA = sDZ + kBwdBlockRows * kDzStride;
  float* sB = sA + (size_t)(a.rows_cap + 1) * kS1;
  float* sX = sB + (size_t)(a.rows_cap + 1) * kS1;
  int32_t* sPtr = reinterpret_cast<int32_t*>(sX + (size_t)a.rows_cap * kp);
  int32_t* sOff = sPtr + a.rows_cap + 4;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x < kS1) {
    sA[a.rows_cap * kS1 + threadIdx.x] = 0.f;
    sB[a.rows_cap * kS1 + threadIdx.x] = 0.f;
  }

  for (int e = threadIdx.x; e < kF2 * kF1; e += kFusedThreads) {
    const float va = __ldg(a.w2a + e), vb = __ldg(a.w2b + e);
    sW2[e] = va;
    sW2[kF2 * kF1 + e] = vb;
    const int c = e / kF1, k = e - c * kF1;
    sW2T[k * kWtStride + c] = va;                     // W2^T: [k (16)][c (32)], row stride 36
    sW2T[kF1 * kWtStride + k * kWtStride + c] = vb;
  }

  // register-resident partial weight gradients, accumulated over all graphs of this CTA
  const int c2 = threadIdx.x & 63;   // dW2: column of the stacked conv2 output
  const int rl2 = threadIdx.x >> 6;  // row lane 0..7
  float dw2[kF1];
#pragma unroll
  for (int k = 0; k < kF1; ++k) dw2[k] = 0.f;
  // dW1s: warp -> (k block of 16: wk = warp & 3, row split wn = warp >> 2); lane -> (mg = lane & 7 -> 4 rows m, kg = lane >> 3 -> 4 cols k)
  const int wk = warp & 3, wn = warp >> 2;
  const int mg = lane & 7, kg = lane >> 3;
  float dw1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dw1[i][j] = 0.f;

  for (int g = blockIdx.x; g < a.num_graphs; g += gridDim.x) {
    const int node0 = __ldg(a.graph_ptr + g);
    const int n = __ldg(a.graph_ptr + g + 1) - node0;
    __syncthreads();
    if (n > a.rows_cap) {
      if (threadIdx.x == 0 && a.status != nullptr) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
      continue;
    }
    // ---- stage A2 rows (group 0) and x rows (group 1; only needed by the last phase)
    for (int e = threadIdx.x; e < n * (kS1 / 4); e += kFusedThreads) cp_async<16>(sA + e * 4, a.a2s + (size_t)node0 * kS1 + e * 4, true);
    cp_async_commit();
    fused_stage_rows(sX, a.x, a.ldx, a.fi, kp, node0, n, a.x_vec);
    cp_async_commit();
    const bool staged = fused_stage_index(sPtr, sOff, a.colptr, a.rowidx, node0, n, a.idx_cap, a.rows_cap, a.status);
    cp_async_wait<1>();
    __syncthreads();

    // ---- dZ2 / dW2 / dA2 in blocks of 64 rows
    const float inv_n = 1.f / fmaxf((float)n, 1.f);
    const float dgc = __ldg(a.dg + (size_t)g * kS2 + c2) / fmaxf((float)n, 1.f);  // d mean / d row = dG / max(n,1) (true division)
    (void)inv_n;
    const int branch = c2 >> 5;
    float w[kF1];
#pragma unroll
    for (int k = 0; k < kF1; ++k) w[k] = sW2[branch * kF2 * kF1 + (c2 & 31) * kF1 + k];
    for (int r0 = 0; r0 < n; r0 += kBwdBlockRows) {
      const int rows = min(kBwdBlockRows, n - r0);
      // (a) dZ2[r, c] = (Z2 > 0) ? dG/n : 0, with Z2 recomputed from A2; dW2[c, :] += dZ2[r, c] * A2[r, branch]
      for (int r = rl2; r < rows; r += 8) {
        const float* arow = sA + (r0 + r) * kS1 + branch * kF1;
        float av[kF1];
#pragma unroll
        for (int k4 = 0; k4 < kF1; k4 += 4) {
          const float4 v = *reinterpret_cast<const float4*>(arow + k4);
          av[k4] = v.x; av[k4 + 1] = v.y; av[k4 + 2] = v.z; av[k4 + 3] = v.w;
        }
        float z = 0.f;
#pragma unroll
        for (int k = 0; k < kF1; ++k) z = fmaf(av[k], w[k], z);
        const float dz = z <= 0.f ? 0.f : dgc;  // threshold_backward on relu(z)
        sDZ[r * kDzStride + c2] = dz;
#pragma unroll
        for (int k = 0; k < kF1; ++k) dw2[k] = fmaf(dz, av[k], dw2[k]);
      }
      __syncthreads();
      // (b) dA2[r, br*16 + k] = sum_c dZ2[r, br*32 + c] * W2[c, k]: per branch a [64 x 32] x [32 x 16] product.
      //     warp -> (branch br = warp & 1, 32-row half hv = (warp >> 1) & 1); lane -> (rg = lane >> 2: rows rg + 8j,
      //     cg = lane & 3: outputs 4cg..4cg+3).  Warps 4..15 have nothing to do in this (short) phase.
      if (warp < 4) {
        const int br = warp & 1, hv = warp >> 1;
        const int rg = lane >> 2, cg = lane & 3;
        float o[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int t = 0; t < 4; ++t) o[j][t] = 0.f;
        const float* dzb = sDZ + (hv * 32 + rg) * kDzStride + br * kF2;
        const float* wtb = sW2T + br * kF1 * kWtStride + (4 * cg) * kWtStride;
#pragma unroll
        for (int c4 = 0; c4 < kF2; c4 += 4) {
          float4 dzv[4], wv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) dzv[j] = *reinterpret_cast<const float4*>(dzb + (8 * j) * kDzStride + c4);
#pragma unroll
          for (int t = 0; t < 4; ++t) wv[t] = *reinterpret_cast<const float4*>(wtb + t * kWtStride + c4);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float acc = o[j][t];
              acc = fmaf(dzv[j].x, wv[t].x, acc);
              acc = fmaf(dzv[j].y, wv[t].y, acc);
              acc = fmaf(dzv[j].z, wv[t].z, acc);
              acc = fmaf(dzv[j].w, wv[t].w, acc);
              o[j][t] = acc;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = hv * 32 + rg + 8 * j;
          if (r < rows) *reinterpret_cast<float4*>(sB + (r0 + r) * kS1 + br * kF1 + 4 * cg) = make_float4(o[j][0], o[j][1], o[j][2], o[j][3]);
        }
      }
      __syncthreads();
    }
    // ---- dZ1 = (A^T dA2) * (H1 > 0)   (tile A is free: A2 is consumed)
    fused_aggregate_any<2>(staged, sB, sA, sPtr, sOff, a.colptr, a.rowidx, node0, n, a.rows_cap, nullptr, a.h1s);
    __syncthreads();
    // ---- Q = A^T dZ1
    fused_aggregate_any<1>(staged, sA, sB, sPtr, sOff, a.colptr, a.rowidx, node0, n, a.rows_cap, nullptr, nullptr);
    cp_async_wait<0>();
    __syncthreads();
    // ---- dW1s[m, k] += sum_r Q[r, m] x[r, k]
    {
      const float* qp = sB + mg * 4;
      const float* xp = sX + wk * 16 + kg * 4;
      const bool k_ok = wk * 16 + kg * 4 < kp;  // kp is a multiple of 4: a float4 is all-in or all-out
      if (k_ok) {
        for (int r = wn; r < n; r += 4) {
          const float4 qa = *reinterpret_cast<const float4*>(qp + r * kS1);
          const float4 xb = *reinterpret_cast<const float4*>(xp + r * kp);
          const float qv[4] = {qa.x, qa.y, qa.z, qa.w};
          const float xv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dw1[i][j] = fmaf(qv[i], xv[j], dw1[i][j]);
        }
      }
    }
  }

  // ---- CTA partials -> global (fixed-order combination of the row splits through smem)
  __syncthreads();
  float* red = smem;  // 8*64*16 floats = 32 KB; everything in smem is dead here and the allocation is >= 38 KB for any capacity
  float* out = a.partial + (size_t)blockIdx.x * kPartial;
  // dW2: [rl2][c2][k]
#pragma unroll
  for (int k = 0; k < kF1; ++k) red[(rl2 * kS2 + c2) * kF1 + k] = dw2[k];
  __syncthreads();
  for (int e = threadIdx.x; e < kS2 * kF1; e += kFusedThreads) {
    float s = 0.f;
#pragma unroll
    for (int rl = 0; rl < 8; ++rl) s += red[rl * kS2 * kF1 + e];
    out[kPartialW1 + e] = s;
  }
  __syncthreads();
  // dW1s: [wn][m][k(64)]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(wn * kS1 + mg * 4 + i) * 64 + wk * 16 + kg * 4 + j] = dw1[i][j];
  __syncthreads();
  for (int e = threadIdx.x; e < kS1 * 64; e += kFusedThreads) {
    float s = 0.f;
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) s += red[w4 * kS1 * 64 + e];
    out[e] = s;
  }
}

// sum the per-CTA partials in CTA order: one warp per output element
__global__ void __launch_bounds__(256) k_ginet_fused_bwd_reduce(const float* __restrict__ partial, int n_partials, int fi,
                                                                float* __restrict__ dw1s, float* __restrict__ dw2a, float* __restrict__ dw2b) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = lane_id();
  const int total = kS1 * fi + kPartialW2;
  if (t >= total) return;
  int src;
  float* dst;
  if (t < kS1 * fi) {
    const int m = t / fi, k = t - m * fi;
    src = m * 64 + k;
    dst = dw1s + t;
  } else {
    const int e = t - kS1 * fi;  // [c (64)][k (16)]
    src = kPartialW1 + e;
    dst = e < kF2 * kF1 ? dw2a + e : dw2b + (e - kF2 * kF1);
  }
  float s = 0.f;
  for (int c = lane; c < n_partials; c += 32) s += partial[(size_t)c * kPartial + src];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) *dst = s;
}

static size_t fused_bwd_smem_bytes(int fi, int rows_cap) {  // without the slot-offset buffer
  const int kp = fused_kp(fi);
  return ((size_t)2 * kF2 * kF1 + 2 * kF1 * kWtStride + kBwdBlockRows * kDzStride + (size_t)(rows_cap + 1) * 2 * kS1 + (size_t)rows_cap * kp + rows_cap + 4) * sizeof(float);
}

static size_t fused_fwd_smem_bytes(int fi, int rows_cap) {  // without the slot-offset buffer
  const int kp = fused_kp(fi);
  const int wa = std::max(kp, kS1);
  return ((size_t)kS1 * kp + 2 * kF2 * kF1 + 8 * kS2 + (size_t)rows_cap * wa + kS1 + (size_t)(rows_cap + 1) * kS1 + rows_cap + 4) * sizeof(float);
}

constexpr size_t kSmemBudget = 227 * 1024;

// slots that fit next to the tiles, capped by what the batch needs (max_graph_edges <= 0: unknown -> take what is left)
static int fused_idx_cap(size_t base_bytes, int max_graph_edges) {
  if (base_bytes >= kSmemBudget) return 0;
  int cap = (int)((kSmemBudget - base_bytes) / 4);
  if (max_graph_edges > 0) cap = std::min(cap, (max_graph_edges + 3) / 4 * 4);
  return cap;
}

}  // namespace drk

extern "C" {

int32_t drk_ginet_fused_max_nodes(int32_t fi) {
  using namespace drk;
  if (fi < 1 || fi > 64) return 0;
  const size_t budget = kSmemBudget;
  int cap = 0;
  for (int c = 32; c <= 4096; c += 32) {
    if (fused_fwd_smem_bytes(fi, c) <= budget && fused_bwd_smem_bytes(fi, c) <= budget) cap = c;
    else break;
  }
  return cap;
}

int drk_ginet_fused_fwd(const float* x, int64_t ldx, int32_t fi, const int32_t* graph_ptr, const int32_t* rowptr, const int32_t* colidx,
                        const float* w1s, const float* w2a, const float* w2b, float* h1s, float* a2s, float* g, int32_t num_graphs,
                        int32_t max_graph_nodes, int32_t max_graph_edges, int32_t* status, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_graphs >= 0 && max_graph_nodes >= 0, DRK_EINVAL, "ginet fused fwd: negative size");
  if (num_graphs == 0) return DRK_OK;
  DRK_REQUIRE(x && graph_ptr && rowptr && w1s && w2a && w2b && g, DRK_EINVAL, "ginet fused fwd: null pointer");
  DRK_REQUIRE(fi >= 1 && fi <= 64, DRK_EUNSUPPORTED, "ginet fused fwd: 1 <= F <= 64 node features supported, got %d", fi);
  const int rows_cap = std::max(32, (max_graph_nodes + 31) / 32 * 32);
  const size_t base = fused_fwd_smem_bytes(fi, rows_cap);
  DRK_REQUIRE(base <= kSmemBudget, DRK_EUNSUPPORTED, "ginet fused fwd: a %d-node graph needs %zu bytes of shared memory", max_graph_nodes, base);
  const int idx_cap = fused_idx_cap(base, max_graph_edges);
  const size_t smem = base + (size_t)idx_cap * 4;
  GinetFwdArgs a{x, ldx, fi, graph_ptr, rowptr, colidx, w1s, w2a, w2b, h1s, a2s, g, status, num_graphs, rows_cap, 1, idx_cap};
  if (ldx % 4 == 0 && fi % 4 == 0 && aligned16(x)) a.x_vec = 4;
  else if (ldx % 2 == 0 && fi % 2 == 0 && aligned8(x)) a.x_vec = 2;
  cudaError_t e = cudaFuncSetAttribute(k_ginet_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "ginet fused fwd: smem opt-in: %s", cudaGetErrorString(e));
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int grid = std::min(num_graphs, kNumSM * ctas_per_sm);
  k_ginet_fused_fwd<<<grid, kFusedThreads, smem, as_stream(stream)>>>(a);
  return finish_launch("ginet fused fwd");
}

size_t drk_ginet_fused_bwd_workspace_bytes(void) { return (size_t)drk::kNumSM * drk::kPartial * sizeof(float); }

int drk_ginet_fused_bwd(const float* x, int64_t ldx, int32_t fi, const int32_t* graph_ptr, const int32_t* colptr, const int32_t* rowidx,
                        const float* w2a, const float* w2b, const float* h1s, const float* a2s, const float* dg, float* dw1s, float* dw2a,
                        float* dw2b, int32_t num_graphs, int32_t max_graph_nodes, int32_t max_graph_edges, int32_t* status, void* workspace,
                        size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_graphs >= 0 && max_graph_nodes >= 0, DRK_EINVAL, "ginet fused bwd: negative size");
  DRK_REQUIRE(x && graph_ptr && colptr && w2a && w2b && h1s && a2s && dg && dw1s && dw2a && dw2b, DRK_EINVAL, "ginet fused bwd: null pointer");
  DRK_REQUIRE(fi >= 1 && fi <= 64, DRK_EUNSUPPORTED, "ginet fused bwd: 1 <= F <= 64 node features supported, got %d", fi);
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= drk_ginet_fused_bwd_workspace_bytes(), DRK_EWORKSPACE, "ginet fused bwd: workspace too small");
  const int rows_cap = std::max(32, (max_graph_nodes + 31) / 32 * 32);
  const size_t base = fused_bwd_smem_bytes(fi, rows_cap);
  DRK_REQUIRE(base <= kSmemBudget, DRK_EUNSUPPORTED, "ginet fused bwd: a %d-node graph needs %zu bytes of shared memory", max_graph_nodes, base);
  const int idx_cap = fused_idx_cap(base, max_graph_edges);
  const size_t smem = base + (size_t)idx_cap * 4;
  GinetBwdArgs a{x, ldx, fi, graph_ptr, colptr, rowidx, w2a, w2b, h1s, a2s, dg, static_cast<float*>(workspace), status, num_graphs, rows_cap, 1, idx_cap};
  if (ldx % 4 == 0 && fi % 4 == 0 && aligned16(x)) a.x_vec = 4;
  else if (ldx % 2 == 0 && fi % 2 == 0 && aligned8(x)) a.x_vec = 2;
  cudaError_t e = cudaFuncSetAttribute(k_ginet_fused_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "ginet fused bwd: smem opt-in: %s", cudaGetErrorString(e));
  const int grid = std::max(1, std::min(num_graphs, kNumSM));
  cudaStream_t st = as_stream(stream);
  k_ginet_fused_bwd<<<grid, kFusedThreads, smem, st>>>(a);
  const int total = kS1 * fi + kPartialW2;
  k_ginet_fused_bwd_reduce<<<ceil_div(total * 32, 256), 256, 0, st>>>(static_cast<float*>(workspace), grid, fi, dw1s, dw2a, dw2b);
  return finish_launch("ginet fused bwd", 2);
}

}  // extern "C"
