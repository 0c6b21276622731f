// Dense node-level projections of the message-passing path, fp32 FFMA from shared memory.
//
//   drk_node_linear   C = act(A op(B) + bias) [* relu-mask]      self.fc(...) ginet.py:45-46,
//                     torch.mm(x, wc|wn) foutnet.py:50-51, _edge_mlp/_node_mlp vanilla_gnn.py:22-24,
//                     and dX = dY W in the backward pass
//   drk_weight_grad   dW = dY^T X (+ dbias)                       autograd of the above
//
// Why not tensor cores: the parity bar is rtol 1e-5 against the reference's fp32 CPU result;
// TF32 (10-bit mantissa) is ~1e-3.  The contractions are tiny (K <= ~100, M <= 64: 0.25 GFLOP
// for the whole 77k-node batch) and sit at ~10 FLOP/B, right at the fp32-SIMT ridge of a B200,
// so the kernels are organised to be FFMA-issue-bound rather than shared-memory-bound:
// every 128-bit shared-memory load feeds 16 FMAs per thread and the tile strides are chosen
// so that the 8 (resp. 4) distinct float4 a warp reads per instruction fall in distinct banks.
#include <algorithm>
#include <cstdlib>

#include "drk_common.cuh"

namespace drk {

// smem row stride for a K-chunk of kc floats: multiple of 4 (float4 rows) with stride/4 odd, so that
// r * stride mod 32 hits 8 distinct 4-bank groups for 8 consecutive rows r.
static inline int padded_stride(int kc) {
  int kp = (kc + 3) / 4 * 4;
  if (((kp / 4) & 1) == 0) kp += 4;
  return kp;
}

constexpr int kLinThreads = 128;  // 4 warps
constexpr int kLinTileN = 128;    // rows per CTA: 4 warps x 8 row-groups x 4 rows
constexpr int kLinKChunk = 64;    // K is processed in chunks of <= 64

struct LinearArgs {
  const float* a;
  int64_t lda;
  const float* b;
  int64_t ldb;
  int32_t trans_b;
  const float* bias;
  const float* mask;
  int64_t ld_mask;
  float* c;
  int64_t ldc;
  int64_t n;
  int32_t k;
  int32_t m;
  int32_t act;
  int32_t a_vec;  // 4 / 2 / 1: widest aligned vector for loading A rows
  int32_t c_vec;  // 4 or 1: vector width for storing C (and loading mask)
  // optional second operand pair: C = act(A op(B) + A2 op(B2) + bias) -- a concat-free Linear on [A | A2]
  const float* a2;
  int64_t lda2;
  const float* b2;
  int64_t ldb2;
  int32_t k2;
  int32_t a2_vec;
};

// ---------------------------------------------------------------- tensor-core variant (row-contiguous operands)
// Same contract as k_node_linear, products on mma.sync m16n8k8 TF32 with error compensation (3xTF32, drk_common.cuh: fp32-level
// accuracy, profiles/tf32_emulation.py).  A CTA owns a 64-column slab of the output: the weights of that slab are split into hi/lo
// B fragments ONCE per CTA (shared memory, [k-step][column tile][lane]); the CTA then walks 64-row tiles of A (4 warps x 16 rows).
// The rows of A must be contiguous in global memory (lda == k): a 64-row tile is then ONE bulk copy (cp.async.bulk, the copy engine)
// into a double-buffered stage, so the next tile streams in while this one is multiplied -- thread-issued cp.async was measured at
// ~10 B/clk/SM on B200, the bulk copy at ~50.  The tile is used in place (row stride k floats, no padding).
constexpr int kTcThreads = 128;
constexpr int kTcRows = 64;
constexpr int kTcCols = 64;

__global__ void __launch_bounds__(kTcThreads) k_node_linear_tc(const LinearArgs p, int num_tiles) {
  extern __shared__ __align__(16) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long tc_bar[2];
  const int ks1 = (p.k + 7) / 8, ks2 = (p.k2 + 7) / 8, ks = ks1 + ks2;
  const int a1_floats = kTcRows * p.k, a2_floats = kTcRows * p.k2;
  const int stage_floats = (a1_floats + a2_floats + 3) & ~3;
  uint4* sW = reinterpret_cast<uint4*>(tc_smem);                 // [ks][8][32]: (hi.b0, hi.b1, lo.b0, lo.b1)
  float* sStage = reinterpret_cast<float*>(sW + ks * 8 * 32);    // 2 x [A tile | A2 tile], 4 spare floats behind the last one
  const int lane = lane_id(), warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * kTcCols;
  const int nt_count = min(8, (p.m - m0 + 7) / 8);
  if (threadIdx.x == 0) {
    mbar_init(&tc_bar[0], 1);
    mbar_init(&tc_bar[1], 1);
    mbar_init_fence();
  }
  __syncthreads();
  // one thread hands a tile's rows to the copy engine; the last 0..12 bytes of a ragged last tile travel through ordinary loads
  auto issue = [&](int tile, int stage) {
    const int64_t row0 = (int64_t)tile * kTcRows;
    const int rows = (int)min((int64_t)kTcRows, p.n - row0);
    float* dst = sStage + stage * stage_floats;
    const uint32_t bytes1 = (uint32_t)rows * p.k * 4u, bytes2 = (uint32_t)rows * p.k2 * 4u;
    const uint32_t bulk1 = bytes1 & ~15u, bulk2 = bytes2 & ~15u;
    if (threadIdx.x == 0) {
      fence_proxy_async();
      mbar_expect_tx(&tc_bar[stage], bulk1 + bulk2);
      if (bulk1) bulk_copy_g2s(dst, p.a + row0 * p.k, bulk1, &tc_bar[stage]);
      if (bulk2) bulk_copy_g2s(dst + a1_floats, p.a2 + row0 * p.k2, bulk2, &tc_bar[stage]);
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + (int)((bytes1 - bulk1) >> 2)) {
      const int w = (int)(bulk1 >> 2) + (threadIdx.x - 32);
      dst[w] = __ldg(p.a + row0 * p.k + w);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + (int)((bytes2 - bulk2) >> 2)) {
      const int w = (int)(bulk2 >> 2) + (threadIdx.x - 64);
      dst[a1_floats + w] = __ldg(p.a2 + row0 * p.k2 + w);
    }
  };
  if ((int)blockIdx.x < num_tiles) issue(blockIdx.x, 0);
  // B fragments: b0 = W[m0 + 8 nt + g][8 s + t], b1 = W[m0 + 8 nt + g][8 s + t + 4]  (W = op(B) as [M, K]; zero beyond M / K).
  // All of a thread's weight loads are issued before the first split (4 entries = 8 loads in flight per trip): the build used to be a
  // chain of ~14 dependent L2 round trips per CTA, as long as the two row tiles a CTA multiplies afterwards.
  for (int e0 = threadIdx.x; e0 < ks * 8 * 32; e0 += 4 * kTcThreads) {
    float w0[4], w1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kTcThreads;
      w0[u] = w1[u] = 0.f;
      if (e < ks * 8 * 32) {
        const int ln = e & 31, nt = (e >> 5) & 7, step = e >> 8;
        const bool second = step >= ks1;
        const int kk = (second ? step - ks1 : step) * 8 + (ln & 3);
        const int ktot = second ? p.k2 : p.k;
        const float* __restrict__ pb = second ? p.b2 : p.b;
        const int64_t ldb = second ? p.ldb2 : p.ldb;
        const int gm = m0 + nt * 8 + (ln >> 2);
        if (gm < p.m) {
          if (kk < ktot) w0[u] = p.trans_b ? __ldg(pb + (int64_t)gm * ldb + kk) : __ldg(pb + (int64_t)kk * ldb + gm);
          if (kk + 4 < ktot) w1[u] = p.trans_b ? __ldg(pb + (int64_t)gm * ldb + kk + 4) : __ldg(pb + (int64_t)(kk + 4) * ldb + gm);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kTcThreads;
      if (e < ks * 8 * 32) {
        uint4 q;
        split_tf32(w0[u], q.x, q.z);
        split_tf32(w1[u], q.y, q.w);
        sW[e] = q;
      }
    }
  }
  __syncthreads();
  uint32_t phase[2] = {0u, 0u};
  int stage = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, stage ^= 1) {
    const int64_t row0 = (int64_t)tile * kTcRows;
    if (tile + (int)gridDim.x < num_tiles) issue(tile + gridDim.x, stage ^ 1);  // that stage was released by the barrier ending the previous trip
    mbar_wait(&tc_bar[stage], phase[stage]);
    phase[stage] ^= 1u;
    __syncthreads();  // the ragged-tail floats written by ordinary stores are visible too
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    for (int op = 0; op < (p.k2 > 0 ? 2 : 1); ++op) {
      const int ktot = op ? p.k2 : p.k;
      const int steps = op ? ks2 : ks1, step0 = op ? ks1 : 0;
      const float* xa = sStage + stage * stage_floats + (op ? a1_floats : 0) + (warp * 16 + g) * ktot + t;
      const float* xb = xa + 8 * ktot;
      for (int s = 0; s < steps; ++s) {
        const bool in1 = s * 8 + t < ktot, in2 = s * 8 + t + 4 < ktot;  // the last k-step reaches past the (unpadded) row: those operands are zero
        uint32_t ahi[4], alo[4];
        split_tf32(in1 ? xa[s * 8] : 0.f, ahi[0], alo[0]);
        split_tf32(in1 ? xb[s * 8] : 0.f, ahi[1], alo[1]);
        split_tf32(in2 ? xa[s * 8 + 4] : 0.f, ahi[2], alo[2]);
        split_tf32(in2 ? xb[s * 8 + 4] : 0.f, ahi[3], alo[3]);
        const uint4* wrow = sW + (size_t)(step0 + s) * 8 * 32 + lane;
        uint4 w[8];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) w[nt] = nt < nt_count ? wrow[nt * 32] : make_uint4(0u, 0u, 0u, 0u);
        // the three MMAs of a compensated product column tile by column tile: dependent MMAs are never back to back
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          if (nt < nt_count) mma_tf32(acc[nt], alo, w[nt].x, w[nt].y);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          if (nt < nt_count) mma_tf32(acc[nt], ahi, w[nt].z, w[nt].w);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          if (nt < nt_count) mma_tf32(acc[nt], ahi, w[nt].x, w[nt].y);
      }
    }
    // epilogue: bias, activation, relu-mask, store.  c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt >= nt_count) continue;
      const int col = m0 + nt * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t gr = row0 + warp * 16 + g + 8 * h;
        if (gr >= p.n) continue;
        float v0 = acc[nt][2 * h], v1 = acc[nt][2 * h + 1];
        if (p.bias != nullptr) {
          if (col < p.m) v0 += __ldg(p.bias + col);
          if (col + 1 < p.m) v1 += __ldg(p.bias + col + 1);
        }
        if (p.act == DRK_ACT_RELU) {
          v0 = v0 < 0.f ? 0.f : v0;
          v1 = v1 < 0.f ? 0.f : v1;
        }
        if (p.mask != nullptr) {
          if (col < p.m) v0 = p.mask[gr * p.ld_mask + col] <= 0.f ? 0.f : v0;
          if (col + 1 < p.m) v1 = p.mask[gr * p.ld_mask + col + 1] <= 0.f ? 0.f : v1;
        }
        float* dst = p.c + gr * p.ldc + col;
        if (col + 1 < p.m && (p.ldc & 1) == 0 && p.c_vec >= 2) {
          *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
        } else {
          if (col < p.m) dst[0] = v0;
          if (col + 1 < p.m) dst[1] = v1;
        }
      }
    }
    fence_proxy_async();
    __syncthreads();  // every warp is done with this stage: the next trip may hand it to the copy engine
  }
}

// CT = output columns per thread (4, 8 or 16) -> CTA column tile of 4*CT = 16 / 32 / 64.
template <int CT>
__global__ void __launch_bounds__(kLinThreads) k_node_linear(const LinearArgs p) {
  constexpr int kTileM = 4 * CT;
  extern __shared__ __align__(16) float smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int cg = lane & 3;   // column group
  const int rg = lane >> 2;  // row group
  const int64_t row0 = (int64_t)blockIdx.x * kLinTileN;
  const int m0 = blockIdx.y * kTileM;

  float acc[4][CT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[j][c] = 0.f;

  const int chunks1 = (p.k + kLinKChunk - 1) / kLinKChunk;
  const int chunks2 = (p.k2 + kLinKChunk - 1) / kLinKChunk;
  for (int chunk = 0; chunk < chunks1 + chunks2; ++chunk) {
    const bool second = chunk >= chunks1;
    const int k0 = (second ? chunk - chunks1 : chunk) * kLinKChunk;
    const int ktot = second ? p.k2 : p.k;
    const float* __restrict__ pa = second ? p.a2 : p.a;
    const float* __restrict__ pb = second ? p.b2 : p.b;
    const int64_t lda = second ? p.lda2 : p.lda;
    const int64_t ldb = second ? p.ldb2 : p.ldb;
    const int avec = second ? p.a2_vec : p.a_vec;
    const int kc = min(kLinKChunk, ktot - k0);
    const int kp = ((kc + 3) / 4 * 4) + ((((kc + 3) / 4) & 1) == 0 ? 4 : 0);  // == padded_stride(kc)
    float* sA = smem;                   // [kLinTileN][kp]
    float* sW = smem + kLinTileN * kp;  // [kTileM][kp], rows permuted
    if (chunk > 0) __syncthreads();

    // ---- A tile: rows row0.., columns k0..k0+kc, zero padded to kp.  cp.async: every thread fires all of its
    // copies back to back (~25 in flight) instead of waiting ~1 us of HBM latency per row.
    {
      const int vec = avec;
      const int nv = kc / vec;            // full vectors per row (kc % vec == 0 whenever vec > 1, see dispatcher)
      const int pv = (kp - nv * vec);     // zero-padding floats per row
      for (int e = threadIdx.x; e < kLinTileN * nv; e += kLinThreads) {
        const int r = e / nv;
        const int v = e - r * nv;
        const int64_t gr = row0 + r;
        const bool ok = gr < p.n;
        const float* src = pa + (ok ? gr : 0) * lda + k0 + v * vec;
        float* dst = sA + r * kp + v * vec;
        if (vec == 4) cp_async<16>(dst, src, ok);
        else if (vec == 2) cp_async<8>(dst, src, ok);
        else cp_async<4>(dst, src, ok);
      }
      for (int e = threadIdx.x; e < kLinTileN * pv; e += kLinThreads) {
        const int r = e / pv;
        sA[r * kp + nv * vec + (e - r * pv)] = 0.f;
      }
      cp_async_commit();
    }
    // ---- W tile: logical row m = 4*cg + t + 16*jj is stored at smem row cg + 4*t + 16*jj
    for (int e = threadIdx.x; e < kTileM * kp; e += kLinThreads) {
      const int ml = e / kp;  // logical column of C within the tile
      const int kk = e - ml * kp;
      const int gm = m0 + ml;
      float v = 0.f;
      if (gm < p.m && kk < kc) v = p.trans_b ? __ldg(pb + (int64_t)gm * ldb + (k0 + kk)) : __ldg(pb + (int64_t)(k0 + kk) * ldb + gm);
      const int srow = ((ml >> 2) & 3) + 4 * (ml & 3) + (ml & ~15);
      sW[srow * kp + kk] = v;
    }
    cp_async_wait<0>();
    __syncthreads();

    const float* a_base = sA + (warp * 32 + rg) * kp;
    const float* w_base = sW + cg * kp;
    for (int k4 = 0; k4 < kp; k4 += 4) {
      float4 av[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const float4*>(a_base + (8 * j) * kp + k4);
#pragma unroll
      for (int jj = 0; jj < CT / 4; ++jj) {
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
    }
  }

  // ---- epilogue: bias, activation, relu-mask, store (float4 per (row, 4-column group))
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t gr = row0 + warp * 32 + rg + 8 * j;
    if (gr >= p.n) continue;
#pragma unroll
    for (int jj = 0; jj < CT / 4; ++jj) {
      const int gm = m0 + 4 * cg + 16 * jj;
      if (gm >= p.m) continue;
      float o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float v = acc[j][jj * 4 + t];
        if (p.bias != nullptr && gm + t < p.m) v += __ldg(p.bias + gm + t);
        if (p.act == DRK_ACT_RELU) v = v < 0.f ? 0.f : v;
        o[t] = v;
      }
      float* dst = p.c + gr * p.ldc + gm;
      if (p.c_vec == 4 && gm + 3 < p.m) {
        if (p.mask != nullptr) {
          const float4 mk = ld_stream_f4(p.mask + gr * p.ld_mask + gm);
          o[0] = mk.x <= 0.f ? 0.f : o[0];
          o[1] = mk.y <= 0.f ? 0.f : o[1];
          o[2] = mk.z <= 0.f ? 0.f : o[2];
          o[3] = mk.w <= 0.f ? 0.f : o[3];
        }
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (gm + t < p.m) {
            float v = o[t];
            if (p.mask != nullptr) v = p.mask[gr * p.ld_mask + gm + t] <= 0.f ? 0.f : v;
            dst[t] = v;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- weight gradient
// dW[m,k] = sum_n dY[n,m] X[n,k].  Persistent CTAs stride over 64-row tiles of (dY, X) staged in
// shared memory as loaded (no transpose: the reduction index n is the smem row, so the float4 a warp
// reads per instruction lie in ONE row -> conflict free for any stride).  Each warp owns a
// 32(m) x 16(k) block of dW in registers (4x4 per lane); WN warps split the rows of a tile and are
// combined through shared memory in a fixed order; CTA partials go to the workspace and a second
// kernel adds them in CTA order.  No atomics -> deterministic.
constexpr int kWgThreads = 256;
constexpr int kWgTileN = 64;

struct WgradArgs {
  const float* dy;
  int64_t ld_dy;
  const float* x;
  int64_t ldx;
  int64_t n;
  int32_t k;
  int32_t m;
  float* partial;  // [gridDim.x][m_pad][k_pad] (+ [gridDim.x][m_pad] bias partials after it)
  float* partial_bias;
  int32_t m_pad;
  int32_t k_pad;
  int32_t dy_vec;  // 4 / 2 / 1: widest aligned vector for the dY resp. X tile copies
  int32_t x_vec;
};

template <int TILE_W>
__device__ __forceinline__ void wg_issue_tile(float* dst, const float* src, int64_t ld, int64_t row0, int64_t n, int col0, int width,
                                              int vec) {
  // dst[r][c] = src[row0 + r][col0 + c] for r < kWgTileN, c < TILE_W; zero-filled outside the matrix
  const int nv = TILE_W / vec;
  for (int e = threadIdx.x; e < kWgTileN * nv; e += kWgThreads) {
    const int r = e / nv;
    const int c = (e - r * nv) * vec;
    const int64_t gr = row0 + r;
    const bool ok = gr < n && col0 + c < width;  // width % vec == 0 and col0 % vec == 0: a vector is all-in or all-out
    const float* s = src + (ok ? gr * ld + col0 + c : 0);
    float* d = dst + r * TILE_W + c;
    if (vec == 4) cp_async<16>(d, s, ok);
    else if (vec == 2) cp_async<8>(d, s, ok);
    else cp_async<4>(d, s, ok);
  }
}

template <int WM, int WK>
__global__ void __launch_bounds__(kWgThreads) k_weight_grad(const WgradArgs p) {
  constexpr int WN = 8 / (WM * WK);
  constexpr int kTileM = 32 * WM;
  constexpr int kTileK = 16 * WK;
  constexpr int kStage = kWgTileN * (kTileM + kTileK);  // floats per pipeline stage
  extern __shared__ __align__(16) float smem[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int wn = warp % WN;
  const int wk = (warp / WN) % WK;
  const int wm = warp / (WN * WK);
  const int mg = lane & 7;
  const int kg = lane >> 3;
  const int m0 = blockIdx.z * kTileM;
  const int k0 = blockIdx.y * kTileK;

  float acc[4][4];
  float accb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const bool bias_lane = (wk == 0 && kg == 0 && blockIdx.y == 0);

  const int64_t n_tiles = (p.n + kWgTileN - 1) / kWgTileN;
  int stage = 0;
  if ((int64_t)blockIdx.x < n_tiles) {
    wg_issue_tile<kTileM>(smem, p.dy, p.ld_dy, (int64_t)blockIdx.x * kWgTileN, p.n, m0, p.m, p.dy_vec);
    wg_issue_tile<kTileK>(smem + kWgTileN * kTileM, p.x, p.ldx, (int64_t)blockIdx.x * kWgTileN, p.n, k0, p.k, p.x_vec);
  }
  cp_async_commit();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {  // prefetch the next tile into the other stage while this one is consumed
      float* nb = smem + (stage ^ 1) * kStage;
      wg_issue_tile<kTileM>(nb, p.dy, p.ld_dy, next * kWgTileN, p.n, m0, p.m, p.dy_vec);
      wg_issue_tile<kTileK>(nb + kWgTileN * kTileM, p.x, p.ldx, next * kWgTileN, p.n, k0, p.k, p.x_vec);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* sDY = smem + stage * kStage;
    const float* sX = sDY + kWgTileN * kTileM;
    const float* dyp = sDY + wm * 32 + mg * 4;
    const float* xp = sX + wk * 16 + kg * 4;
#pragma unroll 4
    for (int r = wn; r < kWgTileN; r += WN) {
      const float4 a = *reinterpret_cast<const float4*>(dyp + r * kTileM);
      const float4 b = *reinterpret_cast<const float4*>(xp + r * kTileK);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        if (bias_lane) accb[i] += av[i];
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  cp_async_wait<0>();

  // combine the WN row-splits in fixed order through smem, then write this CTA's partial
  __syncthreads();
  float* red = smem;  // [WN][kTileM][kTileK]; the launcher sizes smem as max(2 stages, this scratch)
  float* redb = smem + WN * kTileM * kTileK;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ml = wm * 32 + mg * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(wn * kTileM + ml) * kTileK + wk * 16 + kg * 4 + j] = acc[i][j];
    if (bias_lane) redb[wn * kTileM + ml] = accb[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kTileM * kTileK; e += kWgThreads) {
    const int ml = e / kTileK;
    const int kl = e - ml * kTileK;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WN; ++w) s += red[(w * kTileM + ml) * kTileK + kl];
    const int gm = m0 + ml, gk = k0 + kl;
    if (gm < p.m_pad && gk < p.k_pad) p.partial[((int64_t)blockIdx.x * p.m_pad + gm) * p.k_pad + gk] = s;
  }
  if (p.partial_bias != nullptr && blockIdx.y == 0) {
    for (int ml = threadIdx.x; ml < kTileM; ml += kWgThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < WN; ++w) s += redb[w * kTileM + ml];
      if (m0 + ml < p.m_pad) p.partial_bias[(int64_t)blockIdx.x * p.m_pad + m0 + ml] = s;
    }
  }
}

// One warp per output element: lane l adds partials l, l+32, ... in order, then a fixed xor tree.
__global__ void __launch_bounds__(256) k_weight_grad_reduce(const float* __restrict__ partial, const float* __restrict__ partial_bias,
                                                            int32_t n_partials, int32_t m, int32_t k, int32_t m_pad, int32_t k_pad,
                                                            float* __restrict__ dw, int64_t ld_dw, float* __restrict__ dbias,
                                                            int32_t accumulate) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = lane_id();
  const int total = m * k + (dbias != nullptr ? m : 0);
  if (t >= total) return;
  float s = 0.f;
  float* dst;
  if (t < m * k) {
    const int mm = t / k, kk = t - mm * k;
    for (int c = lane; c < n_partials; c += 32) s += partial[((int64_t)c * m_pad + mm) * k_pad + kk];
    dst = dw + (int64_t)mm * ld_dw + kk;
  } else {
    const int mm = t - m * k;
    for (int c = lane; c < n_partials; c += 32) s += partial_bias[(int64_t)c * m_pad + mm];
    dst = dbias + mm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) *dst = accumulate ? *dst + s : s;
}

// ---------------------------------------------------------------- tensor-core variant (row-contiguous operands, M <= 64, K <= 64)
// dW = dY^T X as transposed 3xTF32 products over 64-row tiles that arrive by bulk copy (double-buffered).  The 8 rows of a k-step are
// one warp's: warp = (k-step s = warp & 7, column half h = warp >> 3) multiplies rows 8 s .. 8 s + 7 of every tile by ALL row tiles of
// dW and its half of the column tiles, so every operand element is split by one warp (dY) or two (X) -- the SIMT kernel above runs at
// 36 us for [77 k, 64]^T [77 k, 50], this one is bound by the 35 MB it reads.  The warps' accumulators are folded in k-step order once,
// at the end, and every CTA writes one partial in the layout k_weight_grad_reduce expects.
constexpr int kWtThreads = 512;
constexpr int kWtRows = 64;

struct WgradTcArgs {
  const float* dy;
  const float* x;
  int64_t n;
  int32_t k, m;
  float* partial;
  float* partial_bias;  // may be NULL
  int32_t m_pad, k_pad, num_tiles;
};

__global__ void __launch_bounds__(kWtThreads, 1) k_weight_grad_tc(const WgradTcArgs p) {
  extern __shared__ __align__(16) unsigned char wt_smem[];
  __shared__ __align__(8) unsigned long long wt_bar[2];
  const int m = p.m, k = p.k;
  const int dy_floats = kWtRows * m, x_floats = kWtRows * k;
  const int stage_floats = (dy_floats + x_floats + 7) & ~3;
  float* sStage = reinterpret_cast<float*>(wt_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int s = warp & 7, h = warp >> 3;
  const int mt_count = (m + 15) >> 4, nt_total = (k + 7) >> 3, nth = (nt_total + 1) >> 1;
  const int nt0 = h * nth, nt_count = max(0, min(nth, nt_total - nt0));
  if (tid == 0) {
    mbar_init(&wt_bar[0], 1);
    mbar_init(&wt_bar[1], 1);
    mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int tile, int stage) {
    const int64_t row0 = (int64_t)tile * kWtRows;
    const int rows = (int)min((int64_t)kWtRows, p.n - row0);
    float* dst = sStage + stage * stage_floats;
    const uint32_t b1 = (uint32_t)rows * m * 4u, b2 = (uint32_t)rows * k * 4u, bulk1 = b1 & ~15u, bulk2 = b2 & ~15u;
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&wt_bar[stage], bulk1 + bulk2);
      if (bulk1) bulk_copy_g2s(dst, p.dy + row0 * m, bulk1, &wt_bar[stage]);
      if (bulk2) bulk_copy_g2s(dst + ((dy_floats + 3) & ~3), p.x + row0 * k, bulk2, &wt_bar[stage]);
    }
    if (tid >= 32 && tid < 32 + (int)((b1 - bulk1) >> 2)) {
      const int w = (int)(bulk1 >> 2) + (tid - 32);
      dst[w] = __ldg(p.dy + row0 * m + w);
    }
    if (tid >= 64 && tid < 64 + (int)((b2 - bulk2) >> 2)) {
      const int w = (int)(bulk2 >> 2) + (tid - 64);
      dst[((dy_floats + 3) & ~3) + w] = __ldg(p.x + row0 * k + w);
    }
  };
  if ((int)blockIdx.x < p.num_tiles) issue(blockIdx.x, 0);
  float acc[4][4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  float bias_acc = 0.f;
  uint32_t phase[2] = {0u, 0u};
  int stage = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, stage ^= 1) {
    const int rows = (int)min((int64_t)kWtRows, p.n - (int64_t)tile * kWtRows);
    if (tile + (int)gridDim.x < p.num_tiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&wt_bar[stage], phase[stage]);
    phase[stage] ^= 1u;
    __syncthreads();  // the ragged-tail floats written by ordinary stores are visible too
    const float* sdy = sStage + stage * stage_floats;
    const float* sx = sdy + ((dy_floats + 3) & ~3);
    const int r0 = 8 * s + t, r1 = r0 + 4;
    const bool v0 = r0 < rows, v1 = r1 < rows;  // rows of a ragged last tile beyond the matrix hold stale data
    if (8 * s < rows) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = (nt0 + j) * 8 + g;
        const bool in = j < nt_count && col < k;
        split_tf32((in && v0) ? sx[r0 * k + col] : 0.f, bh[j][0], bl[j][0]);
        split_tf32((in && v1) ? sx[r1 * k + col] : 0.f, bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a >= mt_count) break;
        const int c0 = a * 16 + g, c1 = c0 + 8;
        uint32_t ahi[4], alo[4];
        split_tf32((v0 && c0 < m) ? sdy[r0 * m + c0] : 0.f, ahi[0], alo[0]);
        split_tf32((v0 && c1 < m) ? sdy[r0 * m + c1] : 0.f, ahi[1], alo[1]);
        split_tf32((v1 && c0 < m) ? sdy[r1 * m + c0] : 0.f, ahi[2], alo[2]);
        split_tf32((v1 && c1 < m) ? sdy[r1 * m + c1] : 0.f, ahi[3], alo[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nt_count) mma_tf32(acc[a][j], alo, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nt_count) mma_tf32(acc[a][j], ahi, bl[j][0], bl[j][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nt_count) mma_tf32(acc[a][j], ahi, bh[j][0], bh[j][1]);
      }
    }
    if (p.partial_bias != nullptr && tid < m) {  // column sums of dY in row order
      for (int r = 0; r < rows; ++r) bias_acc += sdy[r * m + tid];
    }
    fence_proxy_async();
    __syncthreads();  // every warp is done with this stage: the next trip may hand it to the copy engine
  }
  // fold the k-step warps' accumulators in k-step order (s = 1..7 into s = 0), then one partial per CTA
  float* scratch = sStage;  // [2 halves][16 units][32 lanes][4]
  for (int r = 1; r < 8; ++r) {
    if (s == r) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(scratch + (((h * 16 + a * 4 + j) * 32 + lane) << 2)) = make_float4(acc[a][j][0], acc[a][j][1], acc[a][j][2], acc[a][j][3]);
    }
    __syncthreads();
    if (s == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(scratch + (((h * 16 + a * 4 + j) * 32 + lane) << 2));
          acc[a][j][0] += v.x;
          acc[a][j][1] += v.y;
          acc[a][j][2] += v.z;
          acc[a][j][3] += v.w;
        }
    }
    __syncthreads();
  }
  if (s == 0) {
    float* part = p.partial + (size_t)blockIdx.x * p.m_pad * p.k_pad;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (a >= mt_count) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= nt_count) continue;
        const int col = (nt0 + j) * 8 + 2 * t;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = a * 16 + g + 8 * hh;
          if (row < p.m_pad) {
            if (col < p.k_pad) part[row * p.k_pad + col] = acc[a][j][2 * hh];
            if (col + 1 < p.k_pad) part[row * p.k_pad + col + 1] = acc[a][j][2 * hh + 1];
          }
        }
      }
    }
  }
  if (p.partial_bias != nullptr && tid < p.m_pad) p.partial_bias[(size_t)blockIdx.x * p.m_pad + tid] = tid < m ? bias_acc : 0.f;
}

struct WgradPlan {
  int wm, wk;
  int grid_x, grid_y, grid_z;
  int m_pad, k_pad;
  size_t bytes;
};

static WgradPlan plan_wgrad(int32_t k, int32_t m) {
  WgradPlan pl{};
  pl.wm = m > 32 ? 2 : 1;
  const int k_tiles16 = (k + 15) / 16;
  const int max_wk = 8 / pl.wm > 4 ? 4 : 8 / pl.wm;
  pl.wk = k_tiles16 >= 4 ? 4 : (k_tiles16 >= 2 ? 2 : 1);
  if (pl.wk > max_wk) pl.wk = max_wk;
  const int tile_m = 32 * pl.wm, tile_k = 16 * pl.wk;
  pl.grid_z = (m + tile_m - 1) / tile_m;
  pl.grid_y = (k + tile_k - 1) / tile_k;
  pl.m_pad = pl.grid_z * tile_m;
  pl.k_pad = pl.grid_y * tile_k;
  pl.grid_x = std::max(1, (kNumSM * 2) / (pl.grid_y * pl.grid_z));
  pl.bytes = ((size_t)pl.grid_x * pl.m_pad * pl.k_pad + (size_t)pl.grid_x * pl.m_pad) * sizeof(float);
  return pl;
}

template <int WM, int WK>
static void launch_wgrad(const WgradArgs& a, const WgradPlan& pl, cudaStream_t st) {
  constexpr int WN = 8 / (WM * WK);
  const size_t tile = (size_t)2 * kWgTileN * (32 * WM + 16 * WK);  // two pipeline stages
  const size_t scratch = (size_t)WN * (32 * WM) * (16 * WK) + (size_t)WN * 32 * WM;  // cross-warp reduction reuses the tile smem
  const size_t smem = std::max(tile, scratch) * sizeof(float);
  dim3 grid(pl.grid_x, pl.grid_y, pl.grid_z);
  static bool opted_in = false;  // > 48 KB of dynamic shared memory needs a one-time opt-in per kernel
  if (!opted_in && smem > 48 * 1024) {
    cudaFuncSetAttribute(k_weight_grad<WM, WK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    opted_in = true;
  }
  k_weight_grad<WM, WK><<<grid, kWgThreads, smem, st>>>(a);
}

}  // namespace drk

extern "C" {

static int a_vector_width(const float* a, int64_t lda, int32_t k) {
  if (lda % 4 == 0 && drk::aligned16(a) && (k % 4 == 0)) return 4;
  if (lda % 2 == 0 && drk::aligned8(a) && (k % 2 == 0)) return 2;
  return 1;
}

int drk_node_linear2(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t k, const float* a2, int64_t lda2, const float* b2,
                     int64_t ldb2, int32_t k2, int32_t trans_b, const float* bias, const float* mask, int64_t ld_mask, float* c,
                     int64_t ldc, int64_t n, int32_t m, int32_t act, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0 && k >= 0 && k2 >= 0 && m >= 0, DRK_EINVAL, "node linear: negative size");
  if (n == 0 || m == 0) return DRK_OK;
  DRK_REQUIRE(a && b && c, DRK_EINVAL, "node linear: null pointer");
  DRK_REQUIRE(k >= 1, DRK_EINVAL, "node linear: k must be >= 1");
  DRK_REQUIRE(k2 == 0 || (a2 && b2), DRK_EINVAL, "node linear: null second operand");
  DRK_REQUIRE(act == DRK_ACT_NONE || act == DRK_ACT_RELU, DRK_EINVAL, "node linear: unknown activation %d", act);
  LinearArgs p{a, lda, b, ldb, trans_b, bias, mask, ld_mask, c, ldc, n, k, m, act, 1, 1, a2, lda2, b2, ldb2, k2, 1};
  // widest vector that keeps every row start and every K-chunk start aligned
  p.a_vec = a_vector_width(a, lda, k);
  if (k2 > 0) p.a2_vec = a_vector_width(a2, lda2, k2);
  if (ldc % 4 == 0 && aligned16(c) && (mask == nullptr || (ld_mask % 4 == 0 && aligned16(mask)))) p.c_vec = 4;
  const int kc = std::min(std::max(k, k2), kLinKChunk);
  const int kp = padded_stride(kc);
  cudaStream_t st = as_stream(stream);
  const unsigned gx = (unsigned)ceil_div<int64_t>(n, kLinTileN);
  auto launch = [&](auto kernel, int tile_m) -> int {
    const size_t smem = (size_t)(kLinTileN + tile_m) * kp * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "node linear: smem opt-in: %s", cudaGetErrorString(e));
    }
    dim3 grid(gx, (unsigned)ceil_div(m, tile_m));
    kernel<<<grid, kLinThreads, smem, st>>>(p);
    return DRK_OK;
  };
  // Row-contiguous operands of large row counts: tensor cores (3xTF32) fed by bulk copies, see k_node_linear_tc.  DRK_LINEAR_TC=0
  // forces the SIMT kernel (also used for column-sliced views, whose rows are not contiguous).
  const char* tc_env = std::getenv("DRK_LINEAR_TC");
  const bool tc_off = tc_env != nullptr && tc_env[0] == '0';
  const bool contiguous = lda == k && aligned16(a) && (k2 == 0 || (lda2 == k2 && aligned16(a2)));
  if (!tc_off && contiguous && m >= 8 && n >= 2048 && k <= 128 && k2 <= 128 && (int64_t)n * std::max(k, k2) < ((int64_t)1 << 31)) {
    const int ks = (k + 7) / 8 + (k2 + 7) / 8;
    const size_t stage = (size_t)((kTcRows * k + kTcRows * k2 + 3) & ~3) * sizeof(float);
    const size_t smem = (size_t)ks * 8 * 32 * sizeof(uint4) + 2 * stage + 16;
    cudaError_t e = cudaFuncSetAttribute(k_node_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "node linear: smem opt-in: %s", cudaGetErrorString(e));
    const int num_tiles = (int)ceil_div<int64_t>(n, kTcRows);
    const int per_sm = std::max(1, std::min(4, (int)((220 * 1024) / (smem + 1024))));
    dim3 grid((unsigned)std::min(num_tiles, kNumSM * per_sm), (unsigned)ceil_div(m, kTcCols));
    if (ldc % 2 == 0 && aligned8(c)) p.c_vec = std::max(p.c_vec, 2);
    k_node_linear_tc<<<grid, kTcThreads, smem, st>>>(p, num_tiles);
    return finish_launch("node linear (tensor cores)");
  }
  int rc;
  if (m <= 16) rc = launch(k_node_linear<4>, 16);
  else if (m <= 32 || (m > 64 && m % 64 != 0 && m % 32 == 0)) rc = launch(k_node_linear<8>, 32);
  else rc = launch(k_node_linear<16>, 64);
  if (rc != DRK_OK) return rc;
  return finish_launch("node linear");
}

int drk_node_linear(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t trans_b, const float* bias, const float* mask,
                    int64_t ld_mask, float* c, int64_t ldc, int64_t n, int32_t k, int32_t m, int32_t act, void* stream) {
  return drk_node_linear2(a, lda, b, ldb, k, nullptr, 0, nullptr, 0, 0, trans_b, bias, mask, ld_mask, c, ldc, n, m, act, stream);
}

static size_t wgrad_tc_bytes(int32_t k, int32_t m) {
  const size_t m_pad = (size_t)(m + 15) / 16 * 16, k_pad = (size_t)(k + 7) / 8 * 8;
  return (size_t)drk::kNumSM * (m_pad * k_pad + m_pad) * sizeof(float);
}

size_t drk_weight_grad_workspace_bytes(int32_t k, int32_t m) {
  if (k <= 0 || m <= 0) return 0;
  return std::max(drk::plan_wgrad(k, m).bytes, wgrad_tc_bytes(k, m));
}

int drk_weight_grad(const float* dy, int64_t ld_dy, const float* x, int64_t ldx, int64_t n, int32_t k, int32_t m, float* dw,
                    int64_t ld_dw, float* dbias, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0 && k >= 1 && m >= 1, DRK_EINVAL, "weight grad: bad size");
  DRK_REQUIRE(dy && x && dw, DRK_EINVAL, "weight grad: null pointer");
  const WgradPlan pl = plan_wgrad(k, m);
  DRK_REQUIRE(workspace != nullptr && workspace_bytes >= pl.bytes, DRK_EWORKSPACE, "weight grad: workspace %zu < %zu bytes",
              workspace_bytes, pl.bytes);
  // Row-contiguous operands of many rows: tensor cores (3xTF32) fed by bulk copies, see k_weight_grad_tc.  DRK_WGRAD_TC=0 disables.
  static const bool wt_off = [] {
    const char* e = std::getenv("DRK_WGRAD_TC");
    return e != nullptr && e[0] == '0';
  }();
  if (!wt_off && ld_dy == m && ldx == k && aligned16(dy) && aligned16(x) && m <= 64 && k <= 64 && n >= 4096 && workspace_bytes >= wgrad_tc_bytes(k, m) &&
      n * std::max(k, m) < ((int64_t)1 << 31)) {
    const int m_pad = (m + 15) / 16 * 16, k_pad = (k + 7) / 8 * 8;
    const int num_tiles = (int)ceil_div<int64_t>(n, kWtRows);
    const int grid = std::min(num_tiles, kNumSM);
    float* part = static_cast<float*>(workspace);
    float* part_bias = part + (size_t)grid * m_pad * k_pad;
    WgradTcArgs t{dy, x, n, k, m, part, dbias != nullptr ? part_bias : nullptr, m_pad, k_pad, num_tiles};
    const size_t stage = (size_t)((kWtRows * m + 3) / 4 * 4 + kWtRows * k + 8) * sizeof(float);
    const size_t smem = std::max<size_t>(2 * stage, (size_t)2 * 16 * 32 * 4 * sizeof(float)) + 64;
    cudaError_t e = cudaFuncSetAttribute(k_weight_grad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DRK_REQUIRE(e == cudaSuccess, DRK_ECUDA, "weight grad: smem opt-in: %s", cudaGetErrorString(e));
    cudaStream_t st_tc = as_stream(stream);
    k_weight_grad_tc<<<grid, kWtThreads, smem, st_tc>>>(t);
    const int total_tc = m * k + (dbias != nullptr ? m : 0);
    k_weight_grad_reduce<<<ceil_div(total_tc * 32, 256), 256, 0, st_tc>>>(part, part_bias, grid, m, k, m_pad, k_pad, dw, ld_dw, dbias, accumulate);
    return finish_launch("weight grad (tensor cores)", 2);
  }
  float* partial = static_cast<float*>(workspace);
  float* partial_bias = partial + (size_t)pl.grid_x * pl.m_pad * pl.k_pad;
  WgradArgs a{dy, ld_dy, x, ldx, n, k, m, partial, dbias != nullptr ? partial_bias : nullptr, pl.m_pad, pl.k_pad, 1, 1};
  if (ld_dy % 4 == 0 && m % 4 == 0 && aligned16(dy)) a.dy_vec = 4;
  else if (ld_dy % 2 == 0 && m % 2 == 0 && aligned8(dy)) a.dy_vec = 2;
  if (ldx % 4 == 0 && k % 4 == 0 && aligned16(x)) a.x_vec = 4;
  else if (ldx % 2 == 0 && k % 2 == 0 && aligned8(x)) a.x_vec = 2;
  cudaStream_t st = as_stream(stream);
  if (pl.wm == 1 && pl.wk == 4) launch_wgrad<1, 4>(a, pl, st);
  else if (pl.wm == 1 && pl.wk == 2) launch_wgrad<1, 2>(a, pl, st);
  else if (pl.wm == 1 && pl.wk == 1) launch_wgrad<1, 1>(a, pl, st);
  else if (pl.wm == 2 && pl.wk == 4) launch_wgrad<2, 4>(a, pl, st);
  else if (pl.wm == 2 && pl.wk == 2) launch_wgrad<2, 2>(a, pl, st);
  else launch_wgrad<2, 1>(a, pl, st);
  const int total = m * k + (dbias != nullptr ? m : 0);
  k_weight_grad_reduce<<<ceil_div(total * 32, 256), 256, 0, st>>>(partial, partial_bias, pl.grid_x, m, k, pl.m_pad, pl.k_pad, dw, ld_dw, dbias,
                                                             accumulate);
  return finish_launch("weight grad", 2);
}

}  // extern "C"
