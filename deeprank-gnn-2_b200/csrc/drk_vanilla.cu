// VanillaConvolutionalLayer (the "NaiveNetwork" convolution; reference deeprank2/neuralnets/gnn/vanilla_gnn.py:10-38) as ONE kernel per
// direction, one CTA per graph of the batch:
//
//   forward    U = x Wa^T + b, V = x Wb^T                      node halves of _edge_mlp  (vanilla_gnn.py:22,30-32)
//              S[i] = sum_{e in row i} relu(U[i] + V[col_e] + C attr_e)           (vanilla_gnn.py:32-35, scatter_sum)
//              out  = relu([x | S] Wn^T + bn)                                     (vanilla_gnn.py:24,36-38)
//   backward   the autograd of the above: dx and, per graph, partial dWe / dbe / dWn / dbn summed in graph order afterwards
//
// Why per graph: a PPI graph (a few hundred nodes, ~20 contacts per node) fits in one SM's shared memory.  The batch-level kernels
// (drk_node_linear -> drk_edge_msg_fwd -> drk_node_linear, drk_sparse.cu / drk_dense.cu) write U|V [N,64], S [N,32] and their gradients
// to HBM and gather V rows through L2; here V stays in shared memory for the whole graph, U and S only ever exist as 64-row tiles, and
// the three projections run on the tensor cores (mma.sync m16n8k8 TF32 with error compensation, fp32-level accuracy: drk_common.cuh)
// from weight fragments that a persistent CTA splits once.  Rows of x stream through a 4-deep ring of 64-row tiles filled by the copy
// engine (cp.async.bulk), prefetched across passes and across graphs.  In the forward edge pass a warp sums two destination rows at a
// time (half a warp each, a lane owns the channel pair (l, l + 16) as one packed fp32x2 operand) and builds the per-edge ReLU mask
// words without a vote in the loop (one 16 x 16 bit transpose per chunk): DESIGN.md 4.5 has the phase clocks that led there.
//
// Data layout: x / out [N, F] row-contiguous, S [N, 32], ReLU masks one 32-bit word per edge in CSR-SLOT order (bit c = channel c),
// edge attributes [E, Fe] in slot order (GraphIndex.attr_in_slot_order) -- all interchangeable with the batch-level kernels, which stay
// the path for graphs that do not fit (drk_vanilla_layer_supported).
#include <algorithm>

#include "drk_common.cuh"

namespace drk {
namespace vanilla {

constexpr int kT = 512;
constexpr int kNW = kT / 32;
constexpr int kRows = 64;    // rows of a node tile: 4 MMA row tiles
constexpr int kStages = 4;   // ring of x tiles
constexpr int kMsg = 32;     // message size fixed by the reference (vanilla_gnn.py:20)
constexpr int kSStride = 36; // row stride of the S tile: A-fragment loads (8 rows x 4 columns) fall in 32 distinct banks
constexpr int kMaxF = 64;
constexpr int kMaxFe = 8;
constexpr size_t kSmemLimit = 227 * 1024 - 256;

struct FwdArgs {
  const float* x;
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* attr;  // [E, fe] in CSR-slot order
  const int32_t* graph_ptr;
  const int32_t* order;  // slot -> graph (data.py:snake_order) or NULL
  const float* we;       // [32, 2F + fe]
  const float* be;       // [32] or NULL
  const float* wn;       // [F, F + 32]
  const float* bn;       // [F] or NULL
  float* out;            // [N, F]
  float* s;              // [N, 32]
  float* cnt;            // [N, 32] active edges per (node, channel), or NULL (inference)
  float* tf;             // [N, fe, 32] sum of the active edges' attributes per (node, feature, channel), or NULL
  uint32_t* mask;        // [E] slot order
  int32_t* status;
  int64_t ld_we, ld_wn;
  int32_t num_graphs, f, fe, rows_cap;
};

// B fragments of Y = A W' for the weight block W'[k][m]: entry ((step * nt_total + nt) * 32 + lane) = (hi.b0, hi.b1, lo.b0, lo.b1) with
// b0 = W'[8 step + t][8 nt + g], b1 = W'[8 step + t + 4][8 nt + g] (g = lane >> 2, t = lane & 3), zero beyond ktot / m.
// kmajor: W'[k][m] = w[(row0 + k) * ld + col0 + m]; otherwise (Y = A W^T of an nn.Linear weight) W'[k][m] = w[(row0 + m) * ld + col0 + k].
__device__ void build_frags(uint4* dst, const float* __restrict__ w, int64_t ld, int row0, int col0, int ktot, int m, int steps, int nt_total, bool kmajor) {
  const int total = steps * nt_total * 32;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * kT) {
    float w0[4], w1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kT;
      w0[u] = w1[u] = 0.f;
      if (e < total) {
        const int ln = e & 31, q = e >> 5, nt = q % nt_total, step = q / nt_total;
        const int kk = step * 8 + (ln & 3), gm = nt * 8 + (ln >> 2);
        if (gm < m) {
          if (kk < ktot) w0[u] = kmajor ? __ldg(w + (int64_t)(row0 + kk) * ld + col0 + gm) : __ldg(w + (int64_t)(row0 + gm) * ld + col0 + kk);
          if (kk + 4 < ktot) w1[u] = kmajor ? __ldg(w + (int64_t)(row0 + kk + 4) * ld + col0 + gm) : __ldg(w + (int64_t)(row0 + gm) * ld + col0 + kk + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kT;
      if (e < total) {
        uint4 q;
        split_tf32(w0[u], q.x, q.z);
        split_tf32(w1[u], q.y, q.w);
        dst[e] = q;
      }
    }
  }
}

// acc[nt] += A[16 rows, ktot] * fragments, one warp.  `a_gt` = &A[g][t] of the warp's row tile (row-major in shared memory, row stride
// lda floats, rows g and g + 8 are read); operands beyond ktot are zero (the rows are not padded).  `frag` = the warp's first column
// tile at k-step 0, + lane; consecutive k-steps are nt_total * 32 entries apart.
template <int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const float* a_gt, int lda, int ktot, int steps, const uint4* frag, int nt_total, int nt_count) {
  const int t = threadIdx.x & 3;
  const float* xa = a_gt;
  const float* xb = a_gt + 8 * lda;
  for (int s = 0; s < steps; ++s) {
    const bool in1 = s * 8 + t < ktot, in2 = s * 8 + t + 4 < ktot;
    uint32_t ahi[4], alo[4];
    split_tf32(in1 ? xa[s * 8] : 0.f, ahi[0], alo[0]);
    split_tf32(in1 ? xb[s * 8] : 0.f, ahi[1], alo[1]);
    split_tf32(in2 ? xa[s * 8 + 4] : 0.f, ahi[2], alo[2]);
    split_tf32(in2 ? xb[s * 8 + 4] : 0.f, ahi[3], alo[3]);
    const uint4* wrow = frag + (size_t)s * nt_total * 32;
    uint4 w[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) w[nt] = nt < nt_count ? wrow[nt * 32] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      if (nt < nt_count) mma_tf32(acc[nt], alo, w[nt].x, w[nt].y);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      if (nt < nt_count) mma_tf32(acc[nt], ahi, w[nt].z, w[nt].w);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      if (nt < nt_count) mma_tf32(acc[nt], ahi, w[nt].x, w[nt].y);
  }
}

// The x tiles a CTA will consume, in order: for every graph of its slots, pass 0 tiles 0..T-1, then pass 1 tiles 0..T-1 (PASSES = 2).
// Warp 0 keeps this cursor and hands the next tile to the copy engine whenever a ring stage is released.
struct TileCursor {
  int slot, pass, tile, n0, n;  // n < 0: past the last graph
  int nx_n0, nx_n;              // the following slot, fetched one graph ahead
};

// node range of the graph in `slot`: n < 0 past the last slot; a graph that does not fit (or a bad id in `order`) has no tiles and is reported
__device__ __forceinline__ bool load_slot(const int32_t* __restrict__ graph_ptr, const int32_t* __restrict__ order, int num_graphs, int rows_cap, int slot, int& n0,
                                          int& n, int* graph = nullptr) {
  n0 = 0;
  n = -1;
  bool invalid = false;
  if (slot < num_graphs) {
    const int g = order != nullptr ? __ldg(order + slot) : slot;
    if (graph != nullptr) *graph = g;
    n = 0;
    if ((unsigned)g < (unsigned)num_graphs) {
      n0 = __ldg(graph_ptr + g);
      n = __ldg(graph_ptr + g + 1) - n0;
    } else {
      invalid = true;
    }
    if (n < 0 || n > rows_cap) {
      invalid = true;
      n = 0;
    }
  }
  return invalid;
}

// ROWS0 / ROWS1 = rows of a tile in pass 0 / pass 1
template <int ROWS0, int ROWS1>
__device__ __forceinline__ void cursor_settle(TileCursor& c, const int32_t* graph_ptr, const int32_t* order, int num_graphs, int rows_cap) {
  while (c.n >= 0 && c.tile * (c.pass == 0 ? ROWS0 : ROWS1) >= c.n) {
    c.tile = 0;
    if (++c.pass == 2 || c.n == 0) {
      c.pass = 0;
      c.slot += gridDim.x;
      c.n0 = c.nx_n0;
      c.n = c.nx_n;
      load_slot(graph_ptr, order, num_graphs, rows_cap, c.slot + gridDim.x, c.nx_n0, c.nx_n);
    }
  }
}

// rows [row0, row0 + rows) of a row-contiguous [*, f] matrix -> stage.  The source need not be 16-byte aligned: the copy starts at the
// aligned address below it (`mis` floats earlier, the consumer skips them) and the last < 16 bytes travel through ordinary loads.
__device__ __forceinline__ uint32_t tile_misalign(const float* src) { return (uint32_t)((reinterpret_cast<uintptr_t>(src) & 15u) >> 2); }

__device__ __forceinline__ uint32_t rows_bulk_bytes(const float* src, int floats) { return ((tile_misalign(src) + (uint32_t)floats) * 4u) & ~15u; }

// lane 0 of the calling warp hands the 16-byte multiples to the copy engine (the barrier must already expect rows_bulk_bytes), lanes 1..3 move the tail
__device__ __forceinline__ void copy_rows(float* dst, const float* src, int floats, void* bar, int lane) {
  const uint32_t mis = tile_misalign(src);
  const float* src_al = src - mis;
  const uint32_t bytes = (mis + (uint32_t)floats) * 4u, bulk = bytes & ~15u;
  if (lane == 0) {
    if (bulk) bulk_copy_g2s(dst, src_al, bulk, bar);
  } else if (lane <= 3) {
    const uint32_t w = (bulk >> 2) + (uint32_t)(lane - 1);
    if (w < (bytes >> 2)) dst[w] = __ldg(src_al + w);
  }
}

// -DDRK_VANILLA_PROBE: thread 0 of every CTA adds up the SM clocks it spends per phase (profiles/vanilla_phase_probe.py reads them)
#ifdef DRK_VANILLA_PROBE
__device__ long long g_vprobe[148][16];  // 0-7 forward, 8-15 backward
#define VPROBE(i)                                   \
  do {                                              \
    if (threadIdx.x == 0) {                         \
      const long long now_ = clock64();             \
      g_vprobe[blockIdx.x][i] += now_ - vprobe_t;   \
      vprobe_t = now_;                              \
    }                                               \
  } while (0)
#else
#define VPROBE(i) \
  do {            \
  } while (0)
#endif

template <int FE>
__global__ void __launch_bounds__(kT, 1) k_vanilla_fwd(const FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kF = FE > 0 ? FE : 1;
  const int f = a.f, ks = (f + 7) >> 3, nto = ks;
  const int stage_floats = (kRows * f + 4 + 3) & ~3;
  uint4* sWU = reinterpret_cast<uint4*>(smem);       // [ks][4][32]  U = x Wa^T
  uint4* sWV = sWU + ks * 4 * 32;                    // [ks][4][32]  V = x Wb^T
  uint4* sWN = sWV + ks * 4 * 32;                    // [ks + 4][nto][32]  out = [x | S] Wn^T
  float* sBe = reinterpret_cast<float*>(sWN + (ks + 4) * nto * 32);
  float* sBn = sBe + kMsg;
  float* sU = sBn + kMaxF;                           // [64][32]
  float* sS = sU + kRows * kMsg;                     // [64][36]
  float* sStage = sS + kRows * kSStride;
  float* sV = sStage + kStages * stage_floats;       // [rows_cap + 1][32], channel-paired; the last row is -inf
  int* sRp = reinterpret_cast<int*>(sV + (size_t)(a.rows_cap + 1) * kMsg);   // [rows_cap + 1] CSR offsets of the graph's rows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) mbar_init(&s_bar[i], 1);
    mbar_init_fence();
  }
  __syncthreads();

  TileCursor cur{0, 0, 0, 0, -1, 0, -1};
  auto produce = [&](int stage) {  // warp 0
    if (cur.n < 0) return;
    const float* src = a.x + (int64_t)(cur.n0 + cur.tile * kRows) * f;
    const int floats = min(kRows, cur.n - cur.tile * kRows) * f;
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&s_bar[stage], rows_bulk_bytes(src, floats));
    }
    copy_rows(sStage + stage * stage_floats, src, floats, &s_bar[stage], lane);
    ++cur.tile;
    cursor_settle<kRows, kRows>(cur, a.graph_ptr, a.order, a.num_graphs, a.rows_cap);
  };
  if (warp == 0) {
    cur.slot = blockIdx.x;
    cur.pass = 0;
    cur.tile = 0;
    load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, cur.slot, cur.n0, cur.n);
    load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, cur.slot + gridDim.x, cur.nx_n0, cur.nx_n);
    cursor_settle<kRows, kRows>(cur, a.graph_ptr, a.order, a.num_graphs, a.rows_cap);
#pragma unroll 1
    for (int i = 0; i < kStages; ++i) produce(i);
  }
  build_frags(sWU, a.we, a.ld_we, 0, 0, f, kMsg, ks, 4, false);
  build_frags(sWV, a.we, a.ld_we, 0, f, f, kMsg, ks, 4, false);
  build_frags(sWN, a.wn, a.ld_wn, 0, 0, f, f, ks, nto, false);
  build_frags(sWN + ks * nto * 32, a.wn, a.ld_wn, 0, f, kMsg, f, 4, nto, false);
  if (tid < kMsg) sBe[tid] = a.be != nullptr ? __ldg(a.be + tid) : 0.f;
  if (tid >= 64 && tid < 64 + kMaxF) sBn[tid - 64] = (a.bn != nullptr && tid - 64 < f) ? __ldg(a.bn + tid - 64) : 0.f;
  float2 cw2[kF];  // C rows of this lane's two message channels (lane mod 16, and + 16)
#pragma unroll
  for (int k = 0; k < kF; ++k)
    cw2[k] = k < FE ? make_float2(__ldg(a.we + (int64_t)(lane & 15) * a.ld_we + 2 * f + k), __ldg(a.we + (int64_t)((lane & 15) + 16) * a.ld_we + 2 * f + k))
                    : make_float2(0.f, 0.f);
  const int dead = a.rows_cap;  // row of -inf behind the graph's V rows: the source of padding edges
  if (tid < kMsg) sV[(size_t)a.rows_cap * kMsg + tid] = -INFINITY;
  __syncthreads();

#ifdef DRK_VANILLA_PROBE
  long long vprobe_t = clock64();
#endif
  bool bad = false, too_big = false;
  uint32_t it = 0;  // tiles consumed so far: stage it % kStages, barrier parity (it / kStages) & 1
  const int mt = warp & 3, ng = warp >> 2;
  int nx_n0, nx_n;  // the next slot's node range, fetched one graph ahead
  too_big |= load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, blockIdx.x, nx_n0, nx_n);
  for (int slot = blockIdx.x; slot < a.num_graphs; slot += gridDim.x) {
    const int n0 = nx_n0, n = nx_n;
    too_big |= load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, slot + gridDim.x, nx_n0, nx_n);
    if (n <= 0) continue;
    for (int j = tid; j <= n; j += kT) sRp[j] = __ldg(a.rowptr + n0 + j);
    const int tiles = (n + kRows - 1) / kRows;
    VPROBE(0);  // graph setup (offsets, rowptr)
    // ---- pass 0: V for every node of the graph
    for (int tile = 0; tile < tiles; ++tile, ++it) {
      const int stage = it & (kStages - 1);
      const float* xt = sStage + stage * stage_floats + tile_misalign(a.x + (int64_t)(n0 + tile * kRows) * f);
      mbar_wait(&s_bar[stage], (it / kStages) & 1u);
      float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      warp_gemm<1>(acc, xt + (mt * 16 + g) * f + t, f, f, ks, sWV + ng * 32 + lane, 4, 1);
      // V and U are kept with channels c and c + 16 next to each other (position 2 (c mod 16) + c / 16): a lane of the edge pass owns such a pair
      const int r = tile * kRows + mt * 16 + g, c = ng * 8 + 2 * t, pc = 2 * (c & 15) + (c >> 4);
      if (r < n) {
        sV[r * kMsg + pc] = acc[0][0];
        sV[r * kMsg + pc + 2] = acc[0][1];
      }
      if (r + 8 < n) {
        sV[(r + 8) * kMsg + pc] = acc[0][2];
        sV[(r + 8) * kMsg + pc + 2] = acc[0][3];
      }
      __syncthreads();
      if (warp == 0) produce(stage);
    }
    // ---- pass 1: per tile U -> edge messages -> node MLP
    VPROBE(1);  // pass 0
    for (int tile = 0; tile < tiles; ++tile, ++it) {
      const int stage = it & (kStages - 1);
      const float* xt = sStage + stage * stage_floats + tile_misalign(a.x + (int64_t)(n0 + tile * kRows) * f);
      // Edge pass layout: a warp sums TWO destination rows at a time -- lanes 0-15 the first, lanes 16-31 the second, lane l the channels
      // l and l + 16 as one packed pair (FADD2 / FFMA2) -- so a warp instruction advances two edges.  Rows r = warp + 16 k, k = 0..3:
      // pairs (k = 0, 1) and (k = 2, 3).  The first 16 edges of every row are fetched while U is multiplied.
      const int hl = lane & 15, half = lane >> 4;
      int e_beg[2], e_len[2], e_src[2][2];  // [pair][chunk of 16 edges]: 32 edges per row are in flight before they are needed
      float e_att[2][2][kF];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int li = tile * kRows + warp + kNW * (2 * p + half);
        e_beg[p] = 0;
        e_len[p] = 0;
        if (li < n) {
          e_beg[p] = sRp[li];
          e_len[p] = sRp[li + 1] - e_beg[p];
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          e_src[p][c] = dead;
#pragma unroll
          for (int q = 0; q < kF; ++q) e_att[p][c][q] = 0.f;
          if (16 * c + hl < e_len[p]) {
            e_src[p][c] = ld_stream_i32(a.colidx + e_beg[p] + 16 * c + hl) - n0;
#pragma unroll
            for (int q = 0; q < FE; ++q) e_att[p][c][q] = ld_stream_f32(a.attr + (int64_t)(e_beg[p] + 16 * c + hl) * FE + q);
          }
        }
      }
      mbar_wait(&s_bar[stage], (it / kStages) & 1u);
      VPROBE(2);  // stage wait
      {
        float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        warp_gemm<1>(acc, xt + (mt * 16 + g) * f + t, f, f, ks, sWU + ng * 32 + lane, 4, 1);
        const int r = mt * 16 + g, c = ng * 8 + 2 * t, pc = 2 * (c & 15) + (c >> 4);
        const float b0 = sBe[c], b1 = sBe[c + 1];
        sU[r * kMsg + pc] = acc[0][0] + b0;
        sU[r * kMsg + pc + 2] = acc[0][1] + b1;
        sU[(r + 8) * kMsg + pc] = acc[0][2] + b0;
        sU[(r + 8) * kMsg + pc + 2] = acc[0][3] + b1;
      }
      __syncthreads();
      VPROBE(3);  // U GEMM + barrier
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int r = warp + kNW * (2 * p + half), li = tile * kRows + r;
        const float2 u = *reinterpret_cast<const float2*>(sU + r * kMsg + 2 * hl);
        const int beg = e_beg[p], len = e_len[p];
        const int len_max = max(len, __shfl_xor_sync(kFull, len, 16));  // the two rows of the pair advance together
        float2 sum = make_float2(0.f, 0.f), active = make_float2(0.f, 0.f);
        float2 tf[kF];
#pragma unroll
        for (int q = 0; q < kF; ++q) tf[q] = make_float2(0.f, 0.f);
        for (int off = 0; off < len_max; off += 16) {
          int src = off == 0 ? e_src[p][0] : e_src[p][1];
          float att[kF];
#pragma unroll
          for (int q = 0; q < kF; ++q) att[q] = off == 0 ? e_att[p][0][q] : e_att[p][1][q];
          if (off >= 32) {  // rows with more than 32 edges: further chunks are fetched in place
            src = dead;
#pragma unroll
            for (int q = 0; q < kF; ++q) att[q] = 0.f;
            if (off + hl < len) {
              src = ld_stream_i32(a.colidx + beg + off + hl) - n0;
#pragma unroll
              for (int q = 0; q < FE; ++q) att[q] = ld_stream_f32(a.attr + (int64_t)(beg + off + hl) * FE + q);
            }
          }
          if ((unsigned)src >= (unsigned)n && src != dead) {  // an edge that leaves the graph: reported, the message is dropped
            bad = true;
            src = dead;
          }
          const int cnt = min(16, len_max - off), mine = len - off;  // `mine` of the chunk's edges belong to this half's row
          // No ballot inside the loop: a vote is a convergent operation the compiler may not move, and with one per edge the chain
          // shuffle -> LDS -> add -> fma -> compare -> vote of edge j had to retire before edge j + 1 could start (230 cycles per trip,
          // profiles/vanilla_phase_probe.py).  Every lane collects its own two channels' bits over the chunk (bit j = edge j); one
          // 16 x 16 bit transpose per chunk (4 shuffle stages) then turns them into one mask word per edge.
          // The loop body is written for the two math pipes (each takes one warp instruction every other cycle per scheduler): packed
          // adds / fmas on the FMA pipe, the 0/1 indicator as a float (FSET) so that the running sums need no select, the mask bits
          // shifted in with an integer multiply-add (bit order reversed, undone once per chunk).
          unsigned bits_lo = 0u, bits_hi = 0u;
#pragma unroll 4
          for (int j = 0; j < cnt; ++j) {
            const int sj = __shfl_sync(kFull, src, j, 16);  // padding edges (the shorter row of the pair) read the -inf row: inactive
            const float2 v = *reinterpret_cast<const float2*>(sV + sj * kMsg + 2 * hl);
            float2 m = __fadd2_rn(u, v);
            float aj[kF];
#pragma unroll
            for (int q = 0; q < FE; ++q) {
              aj[q] = __shfl_sync(kFull, att[q], j, 16);
              m = __ffma2_rn(cw2[q], make_float2(aj[q], aj[q]), m);
            }
            const float2 on = make_float2(m.x > 0.f ? 1.f : 0.f, m.y > 0.f ? 1.f : 0.f);
            sum = __fadd2_rn(sum, make_float2(fmaxf(m.x, 0.f), fmaxf(m.y, 0.f)));
            active = __fadd2_rn(active, on);
#pragma unroll
            for (int q = 0; q < FE; ++q) tf[q] = __ffma2_rn(on, make_float2(aj[q], aj[q]), tf[q]);
            bits_lo = bits_lo * 2u + (__float_as_uint(on.x) >> 29 & 1u);  // 1.0f = 0x3f800000
            bits_hi = bits_hi * 2u + (__float_as_uint(on.y) >> 29 & 1u);
          }
          // edge j sits at bit cnt - 1 - j: reverse
          unsigned bits = (__brev(bits_lo) >> (32 - cnt)) | ((__brev(bits_hi) >> (32 - cnt)) << 16);  // bit j: channel hl of edge j; bit 16 + j: channel 16 + hl
          // transpose inside each half warp: lane l ends up with bit l' of its word = bit l of lane l' -> the mask word of edge l
          // (channels 0-15 in the low half, 16-31 in the high half: bit c = channel c)
#pragma unroll
          for (int st = 1; st < 16; st <<= 1) {
            const unsigned keep = st == 1 ? 0x55555555u : st == 2 ? 0x33333333u : st == 4 ? 0x0f0f0f0fu : 0x00ff00ffu;
            const unsigned other = __shfl_xor_sync(kFull, bits, st, 16);
            bits = (hl & st) ? (((other >> st) & keep) | (bits & ~keep)) : ((bits & keep) | ((other & keep) << st));
          }
          if (hl < mine) a.mask[beg + off + hl] = bits;
        }
        if (li < n) {
          sS[r * kSStride + hl] = sum.x;
          sS[r * kSStride + 16 + hl] = sum.y;
          float* srow = a.s + (int64_t)(n0 + li) * kMsg;
          srow[hl] = sum.x;
          srow[16 + hl] = sum.y;
          if (a.cnt != nullptr) {
            float* crow = a.cnt + (int64_t)(n0 + li) * kMsg;
            crow[hl] = active.x;
            crow[16 + hl] = active.y;
          }
          if (a.tf != nullptr) {
#pragma unroll
            for (int q = 0; q < FE; ++q) {
              float* trow = a.tf + ((int64_t)(n0 + li) * FE + q) * kMsg;
              trow[hl] = tf[q].x;
              trow[16 + hl] = tf[q].y;
            }
          }
        } else {
          sS[r * kSStride + hl] = 0.f;
          sS[r * kSStride + 16 + hl] = 0.f;
        }
      }
      VPROBE(4);  // this warp's edge pass
      __syncthreads();
      VPROBE(5);  // barrier after the edge pass (waiting for the slowest warp)
      {
        const int per = (nto + 3) >> 2, nt0 = ng * per, cnt = min(per, nto - nt0);  // column tiles of this warp (<= 2 for F <= 64)
        if (cnt > 0) {
          float acc[2][4];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
          warp_gemm<2>(acc, xt + (mt * 16 + g) * f + t, f, f, ks, sWN + nt0 * 32 + lane, nto, cnt);
          warp_gemm<2>(acc, sS + (mt * 16 + g) * kSStride + t, kSStride, kMsg, 4, sWN + (ks * nto + nt0) * 32 + lane, nto, cnt);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q >= cnt) continue;
            const int c = (nt0 + q) * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int li = tile * kRows + mt * 16 + g + 8 * h;
              if (li >= n) continue;
              float v0 = acc[q][2 * h] + sBn[c], v1 = acc[q][2 * h + 1] + sBn[c + 1];
              v0 = v0 < 0.f ? 0.f : v0;
              v1 = v1 < 0.f ? 0.f : v1;
              float* dst = a.out + (int64_t)(n0 + li) * f + c;
              if (c + 1 < f && (f & 1) == 0) {
                *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
              } else {
                if (c < f) dst[0] = v0;
                if (c + 1 < f) dst[1] = v1;
              }
            }
          }
        }
      }
      __syncthreads();
      if (warp == 0) produce(stage);
      VPROBE(6);  // out GEMM + closing barrier
    }
  }
  if (a.status != nullptr) {
    if (bad) atomicOr(a.status, DRK_STATUS_CROSS_GRAPH);
    if (too_big && tid == 0) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
  }
}

// =====================================================================================================================================
// backward.  Per graph:   pass 0 (128-row tiles of dout):  dZ = dout * (out > 0),  dS = dZ Wn[:, F:]  -> shared memory, whole graph
//                         pass 1 (64-row tiles of dout | x | S):
//                            dU[i] = dS[i] * cnt[i],  dV[j] = sum_{e: col_e = j} mask_e * dS[row_e]   (CSC walk, dS rows from shared memory)
//                            dx    = [dZ | dU | dV] [Wn[:, :F] ; Wa ; Wb]
//                            dWn|dbn += dZ^T [x | S | 1],   dWa;dWb|dbe += [dU | dV]^T [x | 1],   dC += dS * tf
// The weight gradients accumulate in registers (MMA accumulators) over the tiles of a graph and are written as one partial per graph;
// k_vanilla_reduce adds the partials in graph order -- no floating-point atomics, the result does not depend on the CTA schedule.
struct BwdArgs {
  const float* x;
  const float* s;
  const float* out;
  const float* dout;
  const float* cnt;
  const float* tf;
  const uint32_t* mask;
  const int32_t* colptr;
  const int32_t* rowidx;
  const int32_t* slot_map;  // CSC slot -> CSR slot
  const int32_t* graph_ptr;
  const int32_t* order;
  const float* we;
  const float* wn;
  float* dx;       // [N, F] or NULL
  float* partial;  // [num_graphs][partial_stride]
  int32_t* status;
  int64_t ld_we, ld_wn;
  int32_t num_graphs, f, fe, rows_cap, partial_stride;
};

constexpr int kRowsA = 128;   // rows of a pass-0 tile
constexpr int kStagesB = 2;
constexpr int kDuvStride = 68;

// acc[j] += A^T B over the 64 rows of a tile: A^T[m][k] = am[k * lda + m] (m = the warp's 16 columns from m0, zero beyond m_total), B column tiles
// given per lane as (pointer to the lane's column at row 0, row stride) -- both operands are split on the fly.
template <int NT>
__device__ __forceinline__ void warp_gemm_tn(float (&acc)[NT][4], const float* am, int lda, int m0, int m_total, const float* const (&bp)[NT], const int (&bs)[NT],
                                             int nt_count) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool m_lo = m0 + g < m_total, m_hi = m0 + g + 8 < m_total;
  const float* a0 = am + t * lda + m0 + g;
#pragma unroll 2
  for (int s = 0; s < kRows / 8; ++s) {
    const float* ar = a0 + s * 8 * lda;
    uint32_t ahi[4], alo[4];
    split_tf32(m_lo ? ar[0] : 0.f, ahi[0], alo[0]);
    split_tf32(m_hi ? ar[8] : 0.f, ahi[1], alo[1]);
    split_tf32(m_lo ? ar[4 * lda] : 0.f, ahi[2], alo[2]);
    split_tf32(m_hi ? ar[4 * lda + 8] : 0.f, ahi[3], alo[3]);
    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      if (j < nt_count) {
        split_tf32(bp[j][(s * 8 + t) * bs[j]], bh[j][0], bl[j][0]);
        split_tf32(bp[j][(s * 8 + t + 4) * bs[j]], bh[j][1], bl[j][1]);
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j < nt_count) mma_tf32(acc[j], alo, bh[j][0], bh[j][1]);
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j < nt_count) mma_tf32(acc[j], ahi, bl[j][0], bl[j][1]);
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j < nt_count) mma_tf32(acc[j], ahi, bh[j][0], bh[j][1]);
  }
}

constexpr int kNtN = 4;   // column tiles of dWn|dbn per warp: ceil((F + 33) / 8) <= 13 tiles over 4 warp groups
constexpr int kNtAB = 3;  // column tiles of dWab|dbe per warp: ceil((F + 1) / 8) <= 9

template <int FE>
__global__ void __launch_bounds__(kT, 1) k_vanilla_bwd(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar[kStagesB];
  __shared__ float s_const[2];  // {0, 1}: the padding and the bias column of the transposed products
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kF = FE > 0 ? FE : 1;
  const int f = a.f, ks = (f + 7) >> 3, nto = ks;
  const int sub_floats = (kRows * f + 4 + 3) & ~3;           // dout / x sub-tile of a stage
  const int stage_floats = 2 * sub_floats + kRows * kMsg + 4;  // + S sub-tile
  uint4* sWD = reinterpret_cast<uint4*>(smem);               // [ks][4][32]       dS = dZ Wn[:, F:]
  uint4* sWX = sWD + ks * 4 * 32;                            // [ks + 8][nto][32] dx = [dZ | dU | dV] [Wn[:, :F] ; Wa ; Wb]
  float* sDuv = reinterpret_cast<float*>(sWX + (ks + 8) * nto * 32);  // [64][68]
  float* sStage = sDuv + kRows * kDuvStride;
  float* sDs = sStage + kStagesB * stage_floats;             // [rows_cap][32]
  int* sCp = reinterpret_cast<int*>(sDs + (size_t)a.rows_cap * kMsg);  // [rows_cap + 1] CSC offsets of the graph's nodes

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStagesB; ++i) mbar_init(&s_bar[i], 1);
    mbar_init_fence();
    s_const[0] = 0.f;
    s_const[1] = 1.f;
  }
  __syncthreads();

  TileCursor cur{0, 0, 0, 0, -1, 0, -1};
  auto produce = [&](int stage) {  // warp 0
    if (cur.n < 0) return;
    float* dst = sStage + stage * stage_floats;
    if (cur.pass == 0) {
      const float* src = a.dout + (int64_t)(cur.n0 + cur.tile * kRowsA) * f;
      const int floats = min(kRowsA, cur.n - cur.tile * kRowsA) * f;
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&s_bar[stage], rows_bulk_bytes(src, floats));
      }
      copy_rows(dst, src, floats, &s_bar[stage], lane);
    } else {
      const int64_t row0 = cur.n0 + cur.tile * kRows;
      const int rows = min(kRows, cur.n - cur.tile * kRows);
      const float* s0 = a.dout + row0 * f;
      const float* s1 = a.x + row0 * f;
      const float* s2 = a.s + row0 * kMsg;
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&s_bar[stage], rows_bulk_bytes(s0, rows * f) + rows_bulk_bytes(s1, rows * f) + rows_bulk_bytes(s2, rows * kMsg));
      }
      copy_rows(dst, s0, rows * f, &s_bar[stage], lane);
      copy_rows(dst + sub_floats, s1, rows * f, &s_bar[stage], lane);
      copy_rows(dst + 2 * sub_floats, s2, rows * kMsg, &s_bar[stage], lane);
    }
    ++cur.tile;
    cursor_settle<kRowsA, kRows>(cur, a.graph_ptr, a.order, a.num_graphs, a.rows_cap);
  };
  if (warp == 0) {
    cur.slot = blockIdx.x;
    cur.pass = 0;
    cur.tile = 0;
    load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, cur.slot, cur.n0, cur.n);
    load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, cur.slot + gridDim.x, cur.nx_n0, cur.nx_n);
    cursor_settle<kRowsA, kRows>(cur, a.graph_ptr, a.order, a.num_graphs, a.rows_cap);
#pragma unroll 1
    for (int i = 0; i < kStagesB; ++i) produce(i);
  }
  build_frags(sWD, a.wn, a.ld_wn, 0, f, f, kMsg, ks, 4, true);
  build_frags(sWX, a.wn, a.ld_wn, 0, 0, f, f, ks, nto, true);
  build_frags(sWX + ks * nto * 32, a.we, a.ld_we, 0, 0, kMsg, f, 4, nto, true);
  build_frags(sWX + (ks + 4) * nto * 32, a.we, a.ld_we, 0, f, kMsg, f, 4, nto, true);
  __syncthreads();

  // persistent accumulators of the weight gradients: warp = (row tile mt, column group cg); column tiles cg, cg + 4, ...
  const int mt = warp & 3, cg = warp >> 2;
  const int ntn = (f + kMsg + 1 + 7) >> 3, ntab = (f + 1 + 7) >> 3;
  float accN[kNtN][4], accAB[kNtAB][4], accC[kF];
#pragma unroll
  for (int j = 0; j < kNtN; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) accN[j][i] = 0.f;
#pragma unroll
  for (int j = 0; j < kNtAB; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) accAB[j][i] = 0.f;
#pragma unroll
  for (int q = 0; q < kF; ++q) accC[q] = 0.f;
  const int cntN = cg < ntn ? min(kNtN, (ntn - cg + 3) >> 2) : 0, cntAB = cg < ntab ? min(kNtAB, (ntab - cg + 3) >> 2) : 0;

#ifdef DRK_VANILLA_PROBE
  long long vprobe_t = clock64();
#endif
  bool bad = false, too_big = false;
  uint32_t it = 0;
  const int ldn = f + kMsg + 1, ldab = f + 1;
  int nx_n0, nx_n, nx_g = -1;
  too_big |= load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, blockIdx.x, nx_n0, nx_n, &nx_g);
  for (int slot = blockIdx.x; slot < a.num_graphs; slot += gridDim.x) {
    const int n0 = nx_n0, n = nx_n, gr = nx_g;
    too_big |= load_slot(a.graph_ptr, a.order, a.num_graphs, a.rows_cap, slot + gridDim.x, nx_n0, nx_n, &nx_g);
    if (n <= 0) {  // an empty (or rejected) graph contributes a zero partial
      if ((unsigned)gr < (unsigned)a.num_graphs)
        for (int i = tid; i < a.partial_stride; i += kT) a.partial[(size_t)gr * a.partial_stride + i] = 0.f;
      continue;
    }
    for (int j = tid; j <= n; j += kT) sCp[j] = __ldg(a.colptr + n0 + j);
    VPROBE(9);  // graph setup
    // ---- pass 0: dS for every node
    const int tiles_a = (n + kRowsA - 1) / kRowsA;
    for (int tile = 0; tile < tiles_a; ++tile, ++it) {
      const int stage = it & (kStagesB - 1);
      const int64_t row0 = n0 + tile * kRowsA;
      const int rows = min(kRowsA, n - tile * kRowsA);
      float* dz = sStage + stage * stage_floats + tile_misalign(a.dout + row0 * f);
      constexpr int kPer = (kRowsA * kMaxF + kT - 1) / kT;  // 16
      float o[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int idx = tid + j * kT;
        o[j] = idx < rows * f ? ld_stream_f32(a.out + row0 * f + idx) : 1.f;
      }
      mbar_wait(&s_bar[stage], (it / kStagesB) & 1u);
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int idx = tid + j * kT;
        if (idx < rows * f && o[j] <= 0.f) dz[idx] = 0.f;
      }
      __syncthreads();
      {
        const int mta = warp & 7, nt0 = (warp >> 3) * 2;
        if (mta * 16 < rows) {
          float acc[2][4];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
          warp_gemm<2>(acc, dz + (mta * 16 + g) * f + t, f, f, ks, sWD + nt0 * 32 + lane, 4, 2);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int r = tile * kRowsA + mta * 16 + g, c = (nt0 + q) * 8 + 2 * t;
            if (r < n) *reinterpret_cast<float2*>(sDs + r * kMsg + c) = make_float2(acc[q][0], acc[q][1]);
            if (r + 8 < n) *reinterpret_cast<float2*>(sDs + (r + 8) * kMsg + c) = make_float2(acc[q][2], acc[q][3]);
          }
        }
      }
      __syncthreads();
      if (warp == 0) produce(stage);
    }
    VPROBE(10);  // pass 0: dS product + barriers
    // ---- pass 1
    const int tiles = (n + kRows - 1) / kRows;
    for (int tile = 0; tile < tiles; ++tile, ++it) {
      const int stage = it & (kStagesB - 1);
      const int64_t row0 = n0 + tile * kRows;
      const int rows = min(kRows, n - tile * kRows);
      float* base = sStage + stage * stage_floats;
      float* dz = base + tile_misalign(a.dout + row0 * f);
      float* xt = base + sub_floats + tile_misalign(a.x + row0 * f);
      float* st = base + 2 * sub_floats + tile_misalign(a.s + row0 * kMsg);
      constexpr int kPer = (kRows * kMaxF + kT - 1) / kT;  // 8
      float o[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int idx = tid + j * kT;
        o[j] = idx < rows * f ? ld_stream_f32(a.out + row0 * f + idx) : 1.f;
      }
      // the first 32 incoming-edge records of this warp's 4 source nodes, and the per-node counts
      int c_len[4], c_beg[4], c_row[4];
      uint32_t c_msk[4];
      float c_cnt[4], c_tf[4][kF];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int li = tile * kRows + warp + kNW * k;
        c_beg[k] = 0;
        c_len[k] = 0;
        c_cnt[k] = 0.f;
#pragma unroll
        for (int q = 0; q < kF; ++q) c_tf[k][q] = 0.f;
        if (li < n) {
          c_beg[k] = sCp[li];
          c_len[k] = sCp[li + 1] - c_beg[k];
          c_cnt[k] = ld_stream_f32(a.cnt + (int64_t)(n0 + li) * kMsg + lane);
#pragma unroll
          for (int q = 0; q < FE; ++q) c_tf[k][q] = ld_stream_f32(a.tf + ((int64_t)(n0 + li) * FE + q) * kMsg + lane);  // consumed after the wait below
        }
        c_row[k] = 0;
        c_msk[k] = 0u;
      }
      int c_slot[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        c_slot[k] = -1;
        if (lane < c_len[k]) {
          c_row[k] = ld_stream_i32(a.rowidx + c_beg[k] + lane) - n0;
          c_slot[k] = ld_stream_i32(a.slot_map + c_beg[k] + lane);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c_slot[k] >= 0) c_msk[k] = __ldg(a.mask + c_slot[k]);
      mbar_wait(&s_bar[stage], (it / kStagesB) & 1u);
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int idx = tid + j * kT;
        if (idx < rows * f && o[j] <= 0.f) dz[idx] = 0.f;
      }
      if (rows < kRows) {  // the rows of a ragged tile that the copies did not write are the K dimension of the transposed products
        for (int idx = rows * f + tid; idx < kRows * f; idx += kT) {
          dz[idx] = 0.f;
          xt[idx] = 0.f;
        }
        for (int idx = rows * kMsg + tid; idx < kRows * kMsg; idx += kT) st[idx] = 0.f;
      }
      VPROBE(11);  // pass 1: prefetch, stage wait, dZ in place
      // walks: rows warp, warp + 16, ... of the tile; lane = message channel
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = warp + kNW * k, li = tile * kRows + r;
        float du = 0.f, dv = 0.f;
        if (li < n) {
          const float ds = sDs[li * kMsg + lane];
          du = ds * c_cnt[k];
          if (FE > 0) {
#pragma unroll
            for (int q = 0; q < FE; ++q) accC[q] = fmaf(ds, c_tf[k][q], accC[q]);
          }
          for (int off = 0; off < c_len[k]; off += 32) {
            int rr = c_row[k];
            uint32_t mm = c_msk[k];
            if (off > 0) {
              rr = 0;
              mm = 0u;
              if (off + lane < c_len[k]) {
                rr = ld_stream_i32(a.rowidx + c_beg[k] + off + lane) - n0;
                mm = __ldg(a.mask + ld_stream_i32(a.slot_map + c_beg[k] + off + lane));
              }
            }
            if ((unsigned)rr >= (unsigned)n) {
              bad = true;
              rr = 0;
              mm = 0u;
            }
            const int cnt = min(32, c_len[k] - off);
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
              const int rj = __shfl_sync(kFull, rr, j);
              const uint32_t wj = __shfl_sync(kFull, mm, j);
              const float v = sDs[rj * kMsg + lane];
              dv += ((wj >> lane) & 1u) ? v : 0.f;
            }
          }
        }
        sDuv[r * kDuvStride + lane] = du;
        sDuv[r * kDuvStride + kMsg + lane] = dv;
      }
      VPROBE(12);  // this warp's walks (dU, dV)
      __syncthreads();
      VPROBE(13);  // barrier after the walks
      // dx tile
      if (a.dx != nullptr) {
        const int per = (nto + 3) >> 2, nt0 = cg * per, cnt = min(per, nto - nt0);
        if (cnt > 0 && mt * 16 < rows) {
          float acc[2][4];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
          warp_gemm<2>(acc, dz + (mt * 16 + g) * f + t, f, f, ks, sWX + nt0 * 32 + lane, nto, cnt);
          warp_gemm<2>(acc, sDuv + (mt * 16 + g) * kDuvStride + t, kDuvStride, 2 * kMsg, 8, sWX + (ks * nto + nt0) * 32 + lane, nto, cnt);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q >= cnt) continue;
            const int c = (nt0 + q) * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int lr = mt * 16 + g + 8 * h;
              if (lr >= rows) continue;
              float* dst = a.dx + (row0 + lr) * f + c;
              if (c + 1 < f && (f & 1) == 0) {
                *reinterpret_cast<float2*>(dst) = make_float2(acc[q][2 * h], acc[q][2 * h + 1]);
              } else {
                if (c < f) dst[0] = acc[q][2 * h];
                if (c + 1 < f) dst[1] = acc[q][2 * h + 1];
              }
            }
          }
        }
      }
      VPROBE(14);  // dx product
      // weight gradients: dWn | dbn += dZ^T [x | S | 1]
      {
        const float* bp[kNtN];
        int bs[kNtN];
#pragma unroll
        for (int j = 0; j < kNtN; ++j) {
          const int col = (cg + 4 * j) * 8 + g;
          bp[j] = &s_const[0];
          bs[j] = 0;
          if (col < f) {
            bp[j] = xt + col;
            bs[j] = f;
          } else if (col < f + kMsg) {
            bp[j] = st + (col - f);
            bs[j] = kMsg;
          } else if (col == f + kMsg) {
            bp[j] = &s_const[1];
          }
        }
        if (mt * 16 < f) warp_gemm_tn<kNtN>(accN, dz, f, mt * 16, f, bp, bs, cntN);
      }
      // dWa ; dWb | dbe += [dU | dV]^T [x | 1]
      {
        const float* bp[kNtAB];
        int bs[kNtAB];
#pragma unroll
        for (int j = 0; j < kNtAB; ++j) {
          const int col = (cg + 4 * j) * 8 + g;
          bp[j] = &s_const[0];
          bs[j] = 0;
          if (col < f) {
            bp[j] = xt + col;
            bs[j] = f;
          } else if (col == f) {
            bp[j] = &s_const[1];
          }
        }
        warp_gemm_tn<kNtAB>(accAB, sDuv, kDuvStride, mt * 16, 2 * kMsg, bp, bs, cntAB);
      }
      __syncthreads();
      if (warp == 0) produce(stage);
      VPROBE(15);  // dWn / dWab products + closing barrier
    }
    // ---- this graph's partial: [F][F + 33] (dWn | dbn), [64][F + 1] (dWa ; dWb | dbe), [32][8] (dC).  One partial per GRAPH, added up in
    // graph order by k_vanilla_reduce: the gradients do not depend on which CTA ran which graph (the issue order is a scheduling hint)
    {
      float* part = a.partial + (size_t)gr * a.partial_stride;
#pragma unroll
      for (int j = 0; j < kNtN; ++j) {
        if (j >= cntN) continue;
        const int c = (cg + 4 * j) * 8 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mt * 16 + g + 8 * h;
          if (m < f) {
            if (c < ldn) part[m * ldn + c] = accN[j][2 * h];
            if (c + 1 < ldn) part[m * ldn + c + 1] = accN[j][2 * h + 1];
          }
          accN[j][2 * h] = 0.f;
          accN[j][2 * h + 1] = 0.f;
        }
      }
      float* part_ab = part + f * ldn;
#pragma unroll
      for (int j = 0; j < kNtAB; ++j) {
        if (j >= cntAB) continue;
        const int c = (cg + 4 * j) * 8 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mt * 16 + g + 8 * h;
          if (c < ldab) part_ab[m * ldab + c] = accAB[j][2 * h];
          if (c + 1 < ldab) part_ab[m * ldab + c + 1] = accAB[j][2 * h + 1];
          accAB[j][2 * h] = 0.f;
          accAB[j][2 * h + 1] = 0.f;
        }
      }
      // dC: the warps' channel sums, folded in warp order (sDuv is free: the last tile ended with a barrier)
      float* red = sDuv;  // [warps][32][kF]
#pragma unroll
      for (int q = 0; q < kF; ++q) {
        red[(warp * kMsg + lane) * kF + q] = accC[q];
        accC[q] = 0.f;
      }
      __syncthreads();
      if (warp == 0) {
        float* part_c = part_ab + 2 * kMsg * ldab;
#pragma unroll
        for (int q = 0; q < kMaxFe; ++q) {
          float sum = 0.f;
          if (q < FE) {
            for (int w = 0; w < kNW; ++w) sum += red[(w * kMsg + lane) * kF + q];
          }
          part_c[lane * kMaxFe + q] = sum;
        }
      }
    }
    VPROBE(8);  // this graph's gradient partial
  }
  if (a.status != nullptr) {
    if (bad) atomicOr(a.status, DRK_STATUS_CROSS_GRAPH);
    if (too_big && tid == 0) atomicOr(a.status, DRK_STATUS_INDEX_RANGE);
  }
}

// dwe [32, 2F + fe] | dbe [32] | dwn [F, F + 32] | dbn [F] = sum of the graphs' partials.  Block = 32 outputs x 8 contiguous ranges of graphs;
// a range is summed in graph order (4 interleaved running sums), the 8 range sums in range order: a fixed association.
__global__ void __launch_bounds__(256) k_vanilla_reduce(const float* __restrict__ partial, int parts, int stride, int f, int fe, float* __restrict__ dwe,
                                                       int64_t ld_dwe, float* __restrict__ dbe, float* __restrict__ dwn, int64_t ld_dwn, float* __restrict__ dbn) {
  __shared__ float s_sum[8][32];
  const int ldn = f + kMsg + 1, ldab = f + 1;
  const int n_we = kMsg * (2 * f + fe), n_be = kMsg, n_wn = f * (f + kMsg), n_bn = f;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  int src = 0;
  float* dst = nullptr;
  if (i < n_we) {
    const int c = i / (2 * f + fe), k = i % (2 * f + fe);
    src = k < f ? f * ldn + c * ldab + k : (k < 2 * f ? f * ldn + (kMsg + c) * ldab + (k - f) : f * ldn + 2 * kMsg * ldab + c * kMaxFe + (k - 2 * f));
    dst = dwe + (int64_t)c * ld_dwe + k;
  } else if (i < n_we + n_be) {
    const int c = i - n_we;
    src = f * ldn + c * ldab + f;
    dst = dbe != nullptr ? dbe + c : nullptr;
  } else if (i < n_we + n_be + n_wn) {
    const int j = i - n_we - n_be, m = j / (f + kMsg), k = j % (f + kMsg);
    src = m * ldn + k;
    dst = dwn + (int64_t)m * ld_dwn + k;
  } else if (i < n_we + n_be + n_wn + n_bn) {
    const int m = i - n_we - n_be - n_wn;
    src = m * ldn + f + kMsg;
    dst = dbn != nullptr ? dbn + m : nullptr;
  }
  const int chunk = (parts + 7) / 8, p0 = ty * chunk, p1 = min(parts, p0 + chunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (dst != nullptr) {
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += partial[(size_t)(p + u) * stride + src];
    }
    for (int u = 0; p < p1; ++p, ++u) acc[u] += partial[(size_t)p * stride + src];
  }
  s_sum[ty][tx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (ty == 0 && dst != nullptr) {
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) sum += s_sum[r][tx];
    *dst = sum;
  }
}

static int partial_stride_for(int f) { return (f * (f + kMsg + 1) + 2 * kMsg * (f + 1) + kMsg * kMaxFe + 3) & ~3; }

static size_t bwd_smem_bytes(int f, int rows_cap) {
  const int ks = (f + 7) / 8, nto = ks;
  const size_t sub_floats = (size_t)((kRows * f + 4 + 3) & ~3);
  const size_t stage_floats = 2 * sub_floats + kRows * kMsg + 4;
  size_t b = (size_t)(ks * 4 * 32 + (ks + 8) * nto * 32) * sizeof(uint4);
  b += (size_t)kRows * kDuvStride * 4;
  b += kStagesB * stage_floats * 4;
  b += (size_t)rows_cap * kMsg * 4;
  b += (size_t)(rows_cap + 4) * 4;
  return b;
}

static size_t fwd_smem_bytes(int f, int rows_cap) {
  const int ks = (f + 7) / 8, nto = ks;
  const size_t stage_floats = (size_t)((kRows * f + 4 + 3) & ~3);
  size_t b = (size_t)(ks * 4 * 32 * 2 + (ks + 4) * nto * 32) * sizeof(uint4);
  b += (size_t)(kMsg + kMaxF + kRows * kMsg + kRows * kSStride) * 4;
  b += kStages * stage_floats * 4;
  b += (size_t)(rows_cap + 1) * kMsg * 4;
  b += (size_t)(rows_cap + 4) * 4;
  return b;
}

static int rows_cap_for(int max_graph_nodes) { return std::max(kRows, (max_graph_nodes + kRows - 1) / kRows * kRows); }

}  // namespace vanilla
}  // namespace drk

extern "C" {

int drk_vanilla_layer_supported(int32_t f, int32_t fe, int32_t max_graph_nodes) {
  using namespace drk::vanilla;
  if (f < 1 || f > kMaxF || fe < 0 || fe > kMaxFe || max_graph_nodes < 1) return 0;
  const int cap = rows_cap_for(max_graph_nodes);
  return (fwd_smem_bytes(f, cap) <= kSmemLimit && bwd_smem_bytes(f, cap) <= kSmemLimit) ? 1 : 0;
}

int drk_vanilla_layer_fwd(const float* x, int32_t f, const int32_t* rowptr, const int32_t* colidx, const float* attr_slots, int32_t fe,
                          const int32_t* graph_ptr, const int32_t* order, int32_t num_graphs, int32_t max_graph_nodes, const float* we, int64_t ld_we,
                          const float* be, const float* wn, int64_t ld_wn, const float* bn, float* out, float* s, float* cnt, float* tf, uint32_t* mask,
                          int32_t* status, void* stream) {
  using namespace drk;
  using namespace drk::vanilla;
  DRK_REQUIRE(num_graphs >= 0, DRK_EINVAL, "vanilla layer: negative graph count");
  if (num_graphs == 0) return DRK_OK;
  DRK_REQUIRE(drk_vanilla_layer_supported(f, fe, max_graph_nodes), DRK_EUNSUPPORTED, "vanilla layer: F=%d Fe=%d max graph nodes=%d do not fit one SM (use the batch-level kernels)",
              f, fe, max_graph_nodes);
  DRK_REQUIRE(x && rowptr && colidx && graph_ptr && we && wn && out && s && mask && (fe == 0 || attr_slots), DRK_EINVAL, "vanilla layer: null pointer");
  DRK_REQUIRE(aligned16(x) && aligned8(out), DRK_EINVAL, "vanilla layer: x must be 16-byte aligned, out 8-byte aligned");
  DRK_REQUIRE(ld_we >= 2 * f + fe && ld_wn >= f + kMsg, DRK_EINVAL, "vanilla layer: weight row stride too small");
  FwdArgs a{x, rowptr, colidx, attr_slots, graph_ptr, order, we, be, wn, bn, out, s, cnt, tf, mask, status, ld_we, ld_wn, num_graphs, f, fe, rows_cap_for(max_graph_nodes)};
  const size_t smem = fwd_smem_bytes(f, a.rows_cap);
  const int grid = std::min(num_graphs, kNumSM);
  cudaStream_t st = as_stream(stream);
#define DRK_VANILLA_FWD(FE)                                                                                                     \
  case FE: {                                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(k_vanilla_fwd<FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "vanilla layer: smem opt-in: %s", cudaGetErrorString(e));                          \
    k_vanilla_fwd<FE><<<grid, kT, smem, st>>>(a);                                                                               \
  } break;
  switch (fe) {
    DRK_VANILLA_FWD(0)
    DRK_VANILLA_FWD(1)
    DRK_VANILLA_FWD(2)
    DRK_VANILLA_FWD(3)
    DRK_VANILLA_FWD(4)
    DRK_VANILLA_FWD(5)
    DRK_VANILLA_FWD(6)
    DRK_VANILLA_FWD(7)
    DRK_VANILLA_FWD(8)
  }
#undef DRK_VANILLA_FWD
  return finish_launch("vanilla layer forward");
}

size_t drk_vanilla_layer_bwd_workspace_bytes(int32_t f, int32_t num_graphs) {
  using namespace drk;
  if (f < 1 || num_graphs < 1) return 16;
  return (size_t)num_graphs * (size_t)vanilla::partial_stride_for(f) * sizeof(float);
}

int drk_vanilla_layer_bwd(const float* x, const float* s, const float* out, const float* dout, const float* cnt, const float* tf, int32_t f, int32_t fe,
                          const uint32_t* mask, const int32_t* colptr, const int32_t* rowidx, const int32_t* slot_map, const int32_t* graph_ptr,
                          const int32_t* order, int32_t num_graphs, int32_t max_graph_nodes, const float* we, int64_t ld_we, const float* wn, int64_t ld_wn,
                          float* dx, float* dwe, int64_t ld_dwe, float* dbe, float* dwn, int64_t ld_dwn, float* dbn, int32_t* status, void* workspace,
                          size_t workspace_bytes, void* stream) {
  using namespace drk;
  using namespace drk::vanilla;
  DRK_REQUIRE(num_graphs >= 0, DRK_EINVAL, "vanilla layer backward: negative graph count");
  DRK_REQUIRE(dwe && dwn, DRK_EINVAL, "vanilla layer backward: null gradient pointer");
  DRK_REQUIRE(drk_vanilla_layer_supported(f, fe, std::max(max_graph_nodes, 1)), DRK_EUNSUPPORTED,
              "vanilla layer backward: F=%d Fe=%d max graph nodes=%d do not fit one SM (use the batch-level kernels)", f, fe, max_graph_nodes);
  cudaStream_t st = as_stream(stream);
  const int grid = std::max(1, std::min(num_graphs, kNumSM));
  const int stride = partial_stride_for(f);
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)std::max(num_graphs, 1) * stride * sizeof(float), DRK_EWORKSPACE,
              "vanilla layer backward: workspace too small");
  const int n_out = kMsg * (2 * f + fe) + kMsg + f * (f + kMsg) + f;
  int launches = 1;
  if (num_graphs > 0) {
    DRK_REQUIRE(x && s && out && dout && cnt && mask && colptr && rowidx && slot_map && graph_ptr && we && wn && (fe == 0 || tf), DRK_EINVAL,
                "vanilla layer backward: null pointer");
    DRK_REQUIRE(aligned16(x) && aligned16(s) && aligned16(dout) && (dx == nullptr || aligned8(dx)), DRK_EINVAL,
                "vanilla layer backward: x, s, dout must be 16-byte aligned, dx 8-byte aligned");
    BwdArgs a{x, s, out, dout, cnt, tf, mask, colptr, rowidx, slot_map, graph_ptr, order, we, wn, dx, static_cast<float*>(workspace), status, ld_we, ld_wn,
              num_graphs, f, fe, rows_cap_for(max_graph_nodes), stride};
    const size_t smem = bwd_smem_bytes(f, a.rows_cap);
#define DRK_VANILLA_BWD(FE)                                                                                                     \
  case FE: {                                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(k_vanilla_bwd<FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "vanilla layer backward: smem opt-in: %s", cudaGetErrorString(e));                 \
    k_vanilla_bwd<FE><<<grid, kT, smem, st>>>(a);                                                                               \
  } break;
    switch (fe) {
      DRK_VANILLA_BWD(0)
      DRK_VANILLA_BWD(1)
      DRK_VANILLA_BWD(2)
      DRK_VANILLA_BWD(3)
      DRK_VANILLA_BWD(4)
      DRK_VANILLA_BWD(5)
      DRK_VANILLA_BWD(6)
      DRK_VANILLA_BWD(7)
      DRK_VANILLA_BWD(8)
    }
#undef DRK_VANILLA_BWD
    launches = 2;
  }
  k_vanilla_reduce<<<ceil_div(n_out, 32), 256, 0, st>>>(static_cast<const float*>(workspace), num_graphs, stride, f, fe, dwe, ld_dwe, dbe, dwn,
                                                         ld_dwn, dbn);
  return finish_launch("vanilla layer backward", launches);
}

#ifdef DRK_VANILLA_PROBE
// probe builds only (not part of the ABI): copy the per-CTA phase clocks of the forward kernel to the host and clear them
__attribute__((visibility("default"))) int drk_vanilla_probe_read(long long* host_out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_out, drk::vanilla::g_vprobe, sizeof(long long) * 148 * 16);
  static long long zeros[148 * 16];
  cudaMemcpyToSymbol(drk::vanilla::g_vprobe, zeros, sizeof(zeros));
  return 0;
}
#endif

}  // extern "C"
