// GINet attention with a segment softmax per destination node -- the operator BASELINE.json's north_star describes
// ("edge-attention logits over [x_i, x_j, edge_attr], segment softmax per destination node, scatter-sum aggregation").
//
// The reference builds the logit exactly like this (ginet.py:45-52) but then normalises it with softmax(alpha, dim=1) over the
// singleton axis of a [E,1] tensor (ginet.py:54), which makes every coefficient 1; `attention="reference"` reproduces that.
// `attention="segment_softmax"` is the opt-in intended operator, restated in oracle/restate.py
// (ginet_conv_segment_softmax):
//
//   P = x W^T                                   (drk_node_linear)
//   q_e     = a_r.P[row_e] + a_c.P[col_e] + a_e.(We attr_e)  =  s_r[row_e] + s_c[col_e] + u.attr_e,   u = We^T a_e
//   alpha_e = softmax over {e : row_e = i} of leaky_relu(q_e)
//   z[i]    = sum_{e : row_e = i} alpha_e P[col_e]
//
// Nothing of size [E, F] is ever materialised: the logit needs two scalars per node (s = P [a_r a_c]^T, one 8-byte gather per
// edge) and the F_e edge attributes.  One sub-warp of LPR lanes owns a destination (same layout and lock-step control flow as
// drk_spmm): the lanes split the segment's edges for the max and the sum of the softmax (fixed-order combine), then every lane
// keeps a float4 of the output row while the segment's source rows are gathered in CSR order (= the order of the reference's
// scatter_add_).  No atomics, bit-reproducible.  The backward is two kernels of the same shape: one over destinations (CSR:
// softmax/leaky-relu backward, the logit gradient per edge), one over sources (CSC: the gradient of the projected rows).
//
// Per-edge state lives in CSR-slot order so that the destination kernels stream it: edge_attr is permuted once per batch
// (drk_gather_rows), and one float2 per slot carries (alpha with the sign of the logit, dq).  The source kernel reaches it through
// a CSC-slot -> CSR-slot map built once per batch (drk_attn_slot_map): one 8-byte gather per edge next to the 64-byte row gather.
#include <algorithm>
#include <cmath>

#include "drk_common.cuh"

namespace drk {

constexpr int kAttnThreads = 256;
constexpr int kMaxEdgeFeat = 32;
constexpr unsigned kFullMask = 0xffffffffu;

struct AttnFwdArgs {
  const int32_t* ptr;   // CSR rowptr [n+1]
  const int32_t* idx;   // source node of every CSR slot
  const float* p;       // [n, width] projected rows
  const float* s;       // [n, 2]: (a_r.P[i], a_c.P[i])
  const float* attr;    // [E, fe] in CSR-slot order
  const float* u;       // [fe]
  float* z;             // [n, width]
  float2* adq;          // [E] CSR-slot order: .x = alpha with the sign bit set where the logit is <= 0 (written here), .y = dq (backward)
  float* lg;            // [E] scratch: activated logits in CSR-slot order
  uint32_t ldp, ldz, ld_attr;
  int32_t n, width, fe, act, rows_per_block;
  float slope;
};

// logit of the edge in CSR slot `slot` before the leaky ReLU
__device__ __forceinline__ float edge_logit(const AttnFwdArgs& a, const float* su, int slot, float sr) {
  const int col = ld_stream_i32(a.idx + slot);
  float q = sr + __ldg(a.s + 2 * (size_t)col + 1);
  const float* at = a.attr + (size_t)slot * a.ld_attr;
  for (int k = 0; k < a.fe; ++k) q = fmaf(su[k], ld_stream_f32(at + k), q);
  return q;
}

template <int LPR>
__global__ void __launch_bounds__(kAttnThreads, 3) k_attn_fwd(const AttnFwdArgs a) {
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kRowsPerPass = (kAttnThreads / 32) * kRowsPerWarp;
  constexpr int kChunks = LPR >= 8 ? 1 : 8 / LPR;
  constexpr int kTrip = kChunks * LPR;
  constexpr int kBatch = 8;
  __shared__ float su[kMaxEdgeFeat];
  if ((int)threadIdx.x < a.fe) su[threadIdx.x] = a.u[threadIdx.x];
  __syncthreads();
  const int lane = lane_id();
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const int group_base = sub * LPR;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n);
  const int c = sl * 4;
  const bool col_ok = c < a.width;
  const float slope = a.slope;

  for (int rw = row0 + warp * kRowsPerWarp; rw < row_end; rw += kRowsPerPass) {  // warp-uniform
    const int r = rw + sub;
    const bool row_ok = r < row_end;
    int beg = 0, len = 0;
    float sr = 0.f;
    if (row_ok) {
      beg = __ldg(a.ptr + r);
      len = __ldg(a.ptr + r + 1) - beg;
      sr = __ldg(a.s + 2 * (size_t)r);
    }
    int max_len = len;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) max_len = max(max_len, __shfl_xor_sync(kFullMask, max_len, o));

    // softmax statistics of the segment in ONE sweep: the lanes of the sub-warp split the edges and keep a running
    // (max, sum of exp) each, combined in a fixed order; the activated logits are parked in CSR-slot order (contiguous per
    // segment) so the aggregation sweep does not gather s / edge_attr again.  Every slot is written and read by the same lane.
    float m = -INFINITY, l = 0.f;
    for (int off = 0; off < max_len; off += LPR) {
      const int e = off + sl;
      if (e < len) {
        const float q = edge_logit(a, su, beg + e, sr);
        const float lg = q > 0.f ? q : q * slope;
        a.lg[beg + e] = lg;
        if (lg > m) {
          l = fmaf(l, expf(m - lg), 1.f);
          m = lg;
        } else {
          l += expf(lg - m);
        }
      }
    }
    float mx = m;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    l = m == -INFINITY ? 0.f : l * expf(m - mx);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) l += __shfl_xor_sync(kFullMask, l, o);
    m = mx;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int off = 0; off < max_len; off += kTrip) {
      int my_col[kChunks];
      float my_w[kChunks];
#pragma unroll
      for (int qd = 0; qd < kChunks; ++qd) {
        const int e = off + qd * LPR + sl;
        my_col[qd] = -1;
        my_w[qd] = 0.f;
        if (e < len) {
          const float lg = a.lg[beg + e];  // plain load: written above by this lane
          const float al = expf(lg - m) / l;
          my_col[qd] = ld_stream_i32(a.idx + beg + e);
          my_w[qd] = al;
          a.adq[beg + e].x = lg > 0.f ? al : __uint_as_float(__float_as_uint(al) | 0x80000000u);  // leaky_relu keeps the sign of q
        }
      }
#pragma unroll
      for (int u0 = 0; u0 < kTrip; u0 += kBatch) {
        float4 t[kBatch];
        int srow[kBatch];
        float tw[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          srow[u] = __shfl_sync(kFullMask, my_col[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
          tw[u] = __shfl_sync(kFullMask, my_w[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (srow[u] >= 0 && col_ok) t[u] = ld_gather_f4(a.p + (size_t)(uint32_t)srow[u] * a.ldp + c);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (srow[u] >= 0 && col_ok) {  // sequential accumulation in CSR order
            acc[0] = fmaf(tw[u], t[u].x, acc[0]);
            acc[1] = fmaf(tw[u], t[u].y, acc[1]);
            acc[2] = fmaf(tw[u], t[u].z, acc[2]);
            acc[3] = fmaf(tw[u], t[u].w, acc[3]);
          }
        }
      }
    }
    if (!row_ok || !col_ok) continue;
    if (a.act == DRK_ACT_RELU) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = acc[k] < 0.f ? 0.f : acc[k];
    }
    *reinterpret_cast<float4*>(a.z + (size_t)r * a.ldz + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------- backward over destinations (CSR)
// dz = dy (* (y > 0) if the ReLU was fused);  c_i = dz[i].z[i] = sum_e alpha_e dalpha_e  (y may stand in for z: where they
// differ dz is 0);  dalpha_e = dz[i].P[col_e];  dq_e = alpha_e lrelu'(q_e) (dalpha_e - c_i);  ds_r[i] = sum_e dq_e.
struct AttnBwdDstArgs {
  const int32_t* ptr;
  const int32_t* idx;
  const float* p;
  const float* dy;
  const float* y;
  float2* adq;  // [E] CSR-slot order: .x read, .y = dq written
  float slope;
  float* ds;  // [n, 2]: column 0 written here
  float* dz;  // [n, width] or NULL (only needed when the ReLU was fused)
  uint32_t ldp, ld_dy, ld_y, ld_dz;
  int32_t n, width, act, rows_per_block;
};

template <int LPR>
__global__ void __launch_bounds__(kAttnThreads, 3) k_attn_bwd_dst(const AttnBwdDstArgs a) {
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kRowsPerPass = (kAttnThreads / 32) * kRowsPerWarp;
  constexpr int kBatch = LPR < 8 ? LPR : 8;
  const int lane = lane_id();
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const int group_base = sub * LPR;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n);
  const int c = sl * 4;
  const bool col_ok = c < a.width;

  for (int rw = row0 + warp * kRowsPerWarp; rw < row_end; rw += kRowsPerPass) {
    const int r = rw + sub;
    const bool row_ok = r < row_end;
    int beg = 0, len = 0;
    float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
    float cdot = 0.f;
    if (row_ok) {
      beg = __ldg(a.ptr + r);
      len = __ldg(a.ptr + r + 1) - beg;
      if (col_ok) {
        dz = ld_stream_f4(a.dy + (size_t)r * a.ld_dy + c);
        const float4 yv = ld_stream_f4(a.y + (size_t)r * a.ld_y + c);
        if (a.act == DRK_ACT_RELU) {
          dz.x = yv.x <= 0.f ? 0.f : dz.x;
          dz.y = yv.y <= 0.f ? 0.f : dz.y;
          dz.z = yv.z <= 0.f ? 0.f : dz.z;
          dz.w = yv.w <= 0.f ? 0.f : dz.w;
        }
        if (a.dz != nullptr) *reinterpret_cast<float4*>(a.dz + (size_t)r * a.ld_dz + c) = dz;
        cdot = fmaf(dz.w, yv.w, fmaf(dz.z, yv.z, fmaf(dz.y, yv.y, dz.x * yv.x)));
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) cdot += __shfl_xor_sync(kFullMask, cdot, o);
    int max_len = len;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) max_len = max(max_len, __shfl_xor_sync(kFullMask, max_len, o));

    float dsr = 0.f;
    for (int off = 0; off < max_len; off += LPR) {
      const int e = off + sl;
      int col = -1;
      float sal = 0.f;
      if (e < len) {
        col = ld_stream_i32(a.idx + beg + e);
        const float v = a.adq[beg + e].x;
        sal = v < 0.f || __float_as_uint(v) == 0x80000000u ? -v * a.slope : v;
      }
      float mine = 0.f;
#pragma unroll
      for (int u0 = 0; u0 < LPR; u0 += kBatch) {
        float4 t[kBatch];
        int srow[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) srow[u] = __shfl_sync(kFullMask, col, group_base + u0 + u);
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (srow[u] >= 0 && col_ok) t[u] = ld_gather_f4(a.p + (size_t)(uint32_t)srow[u] * a.ldp + c);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          float d = fmaf(dz.w, t[u].w, fmaf(dz.z, t[u].z, fmaf(dz.y, t[u].y, dz.x * t[u].x)));
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFullMask, d, o);
          if (sl == u0 + u) mine = d;
        }
      }
      if (e < len) {
        const float g = sal * (mine - cdot);
        a.adq[beg + e].y = g;
        dsr += g;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) dsr += __shfl_xor_sync(kFullMask, dsr, o);
    if (row_ok && sl == 0) a.ds[2 * (size_t)r] = dsr;
  }
}

// ---------------------------------------------------------------- backward over sources (CSC)
// dP[j] = sum_{e : col_e = j} alpha_e dz[row_e]  +  ds_r[j] a_r  +  ds_c[j] a_c,   ds_c[j] = sum_{e : col_e = j} dq_e
struct AttnBwdSrcArgs {
  const int32_t* ptr;  // CSC colptr
  const int32_t* idx;  // destination node of every CSC slot
  const int32_t* map;  // CSR slot of every CSC slot
  const float* dz;
  const float2* adq;  // [E] CSR-slot order: (signed alpha, dq)
  const float* att;   // [2*width]: a_r then a_c
  float* ds;         // [n,2]: column 0 read, column 1 written
  float* dp;
  uint32_t ld_dz, ld_dp;
  int32_t n, width, rows_per_block;
};

template <int LPR>
__global__ void __launch_bounds__(kAttnThreads, 3) k_attn_bwd_src(const AttnBwdSrcArgs a) {
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kRowsPerPass = (kAttnThreads / 32) * kRowsPerWarp;
  constexpr int kChunks = LPR >= 8 ? 1 : 8 / LPR;
  constexpr int kTrip = kChunks * LPR;
  constexpr int kBatch = 8;
  const int lane = lane_id();
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const int group_base = sub * LPR;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * a.rows_per_block;
  const int row_end = min(row0 + a.rows_per_block, a.n);
  const int c = sl * 4;
  const bool col_ok = c < a.width;
  float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), ac = ar;
  if (col_ok) {
    ar = ld_gather_f4(a.att + c);
    ac = ld_gather_f4(a.att + a.width + c);
  }

  for (int rw = row0 + warp * kRowsPerWarp; rw < row_end; rw += kRowsPerPass) {
    const int r = rw + sub;
    const bool row_ok = r < row_end;
    int beg = 0, len = 0;
    if (row_ok) {
      beg = __ldg(a.ptr + r);
      len = __ldg(a.ptr + r + 1) - beg;
    }
    int max_len = len;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) max_len = max(max_len, __shfl_xor_sync(kFullMask, max_len, o));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float dsc = 0.f;
    for (int off = 0; off < max_len; off += kTrip) {
      int my_row[kChunks];
      float my_w[kChunks];
#pragma unroll
      for (int qd = 0; qd < kChunks; ++qd) {
        const int e = off + qd * LPR + sl;
        my_row[qd] = -1;
        my_w[qd] = 0.f;
        if (e < len) {
          my_row[qd] = ld_stream_i32(a.idx + beg + e);
          const float2 v = __ldg(a.adq + ld_stream_i32(a.map + beg + e));
          my_w[qd] = fabsf(v.x);
          dsc += v.y;
        }
      }
#pragma unroll
      for (int u0 = 0; u0 < kTrip; u0 += kBatch) {
        float4 t[kBatch];
        int srow[kBatch];
        float tw[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          srow[u] = __shfl_sync(kFullMask, my_row[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
          tw[u] = __shfl_sync(kFullMask, my_w[(u0 + u) / LPR], group_base + ((u0 + u) % LPR));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (srow[u] >= 0 && col_ok) t[u] = ld_gather_f4(a.dz + (size_t)(uint32_t)srow[u] * a.ld_dz + c);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (srow[u] >= 0 && col_ok) {
            acc[0] = fmaf(tw[u], t[u].x, acc[0]);
            acc[1] = fmaf(tw[u], t[u].y, acc[1]);
            acc[2] = fmaf(tw[u], t[u].z, acc[2]);
            acc[3] = fmaf(tw[u], t[u].w, acc[3]);
          }
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) dsc += __shfl_xor_sync(kFullMask, dsc, o);
    if (!row_ok) continue;
    if (sl == 0) a.ds[2 * (size_t)r + 1] = dsc;
    if (!col_ok) continue;
    const float dsr = a.ds[2 * (size_t)r];
    acc[0] += fmaf(dsr, ar.x, dsc * ac.x);
    acc[1] += fmaf(dsr, ar.y, dsc * ac.y);
    acc[2] += fmaf(dsr, ar.z, dsc * ac.z);
    acc[3] += fmaf(dsr, ar.w, dsc * ac.w);
    *reinterpret_cast<float4*>(a.dp + (size_t)r * a.ld_dp + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------- g[k] = sum_e dq[e] attr[e,k]  (two-stage, fixed order)
constexpr int kEdgeGradBlocks = 2 * kNumSM;
constexpr int kEdgeGradThreads = 256;

__global__ void __launch_bounds__(kEdgeGradThreads) k_attn_edge_grad_partial(const float2* __restrict__ adq, const float* __restrict__ attr,
                                                                             uint32_t ld_attr, int64_t num_edges, int32_t fe,
                                                                             float* __restrict__ partial) {
  __shared__ float warp_sum[kEdgeGradThreads / 32];
  const int k = blockIdx.y;
  const int64_t per = (num_edges + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per;
  const int64_t end = beg + per < num_edges ? beg + per : num_edges;
  float acc = 0.f;
  for (int64_t e = beg + threadIdx.x; e < end; e += kEdgeGradThreads) acc = fmaf(adq[e].y, ld_stream_f32(attr + e * ld_attr + k), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  if (lane_id() == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kEdgeGradThreads / 32; ++w) t += warp_sum[w];
    partial[(size_t)blockIdx.x * fe + k] = t;
  }
}

__global__ void k_attn_edge_grad_final(const float* __restrict__ partial, int32_t blocks, int32_t fe, float* __restrict__ g) {
  const int k = blockIdx.x;  // one warp per edge feature: lane-strided partial sums, fixed-order combine
  float t = 0.f;
  for (int b = threadIdx.x; b < blocks; b += 32) t += partial[(size_t)b * fe + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFullMask, t, o);
  if (threadIdx.x == 0) g[k] = t;
}

// CSC slot -> CSR slot: inv[perm[s]] = s, then map[t] = inv[permT[t]]
__global__ void k_attn_invert_perm(const int32_t* __restrict__ perm, int64_t num_edges, int32_t* __restrict__ inv) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < num_edges) inv[ld_stream_i32(perm + s)] = (int32_t)s;
}
__global__ void k_attn_compose_map(const int32_t* __restrict__ permT, const int32_t* __restrict__ inv, int64_t num_edges, int32_t* __restrict__ map) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < num_edges) map[t] = __ldg(inv + ld_stream_i32(permT + t));
}

static int attn_lpr(int width) {
  int lpr = 4;
  while (lpr < 32 && lpr * 4 < width) lpr <<= 1;
  return lpr;
}

static int attn_rows_per_block(int n, int lpr) {
  const int rows_per_pass = (kAttnThreads / 32) * (32 / lpr);
  int passes = (int)ceil_div<int64_t>(n, (int64_t)kNumSM * 8 * rows_per_pass);
  passes = std::max(1, std::min(passes, 8));
  return rows_per_pass * passes;
}

static bool attn_ld_ok(int64_t ld) { return ld >= 0 && ld % 4 == 0 && ld < ((int64_t)1 << 30); }

#define DRK_ATTN_DISPATCH(KERNEL, ARGS, BLOCKS, ST)                    \
  do {                                                                 \
    switch (lpr) {                                                     \
      case 4: KERNEL<4><<<BLOCKS, kAttnThreads, 0, ST>>>(ARGS); break;   \
      case 8: KERNEL<8><<<BLOCKS, kAttnThreads, 0, ST>>>(ARGS); break;   \
      case 16: KERNEL<16><<<BLOCKS, kAttnThreads, 0, ST>>>(ARGS); break; \
      default: KERNEL<32><<<BLOCKS, kAttnThreads, 0, ST>>>(ARGS); break; \
    }                                                                  \
  } while (0)

}  // namespace drk

extern "C" {

int drk_attn_supported(int32_t width, int32_t fe) { return width >= 4 && width % 4 == 0 && width <= 128 && fe >= 0 && fe <= drk::kMaxEdgeFeat; }

int drk_attn_fwd(const int32_t* rowptr, const int32_t* colidx, const float* p, int64_t ldp, const float* s, const float* attr_csr,
                 int64_t ld_attr, int32_t fe, const float* u, float slope, float* z, int64_t ldz, float* adq, float* logit_scratch, int32_t n,
                 int32_t width, int32_t act, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0, DRK_EINVAL, "attention fwd: negative size");
  DRK_REQUIRE(drk_attn_supported(width, fe), DRK_EUNSUPPORTED, "attention fwd: width %d (multiple of 4, <= 128) / %d edge features (<= %d)", width,
              fe, kMaxEdgeFeat);
  DRK_REQUIRE(act == DRK_ACT_NONE || act == DRK_ACT_RELU, DRK_EINVAL, "attention fwd: unknown activation %d", act);
  if (n == 0) return DRK_OK;
  DRK_REQUIRE(rowptr && colidx && p && s && z && adq && logit_scratch && (fe == 0 || (attr_csr && u)), DRK_EINVAL, "attention fwd: null pointer");
  DRK_REQUIRE(attn_ld_ok(ldp) && attn_ld_ok(ldz) && aligned16(p) && aligned16(z) && aligned8(adq) && ld_attr >= 0 && ld_attr < ((int64_t)1 << 30),
              DRK_EUNSUPPORTED, "attention fwd: rows must be 16-byte aligned");
  const int lpr = attn_lpr(width);
  AttnFwdArgs a{rowptr, colidx, p, s, attr_csr, u, z, reinterpret_cast<float2*>(adq), logit_scratch, (uint32_t)ldp, (uint32_t)ldz, (uint32_t)ld_attr,
                n, width, fe, act, attn_rows_per_block(n, lpr), slope};
  const int blocks = ceil_div(n, a.rows_per_block);
  cudaStream_t st = as_stream(stream);
  DRK_ATTN_DISPATCH(k_attn_fwd, a, blocks, st);
  return finish_launch("attention fwd");
}

int drk_attn_bwd_dst(const int32_t* rowptr, const int32_t* colidx, const float* p, int64_t ldp, const float* dy, int64_t ld_dy, const float* y,
                     int64_t ld_y, float* adq, float slope, float* ds, float* dz, int64_t ld_dz, int32_t n, int32_t width, int32_t act,
                     void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0, DRK_EINVAL, "attention bwd: negative size");
  DRK_REQUIRE(drk_attn_supported(width, 0), DRK_EUNSUPPORTED, "attention bwd: width %d (multiple of 4, <= 128)", width);
  if (n == 0) return DRK_OK;
  DRK_REQUIRE(rowptr && colidx && p && dy && y && adq && ds, DRK_EINVAL, "attention bwd: null pointer");
  DRK_REQUIRE(act == DRK_ACT_NONE || dz != nullptr, DRK_EINVAL, "attention bwd: dz is required when the ReLU was fused");
  DRK_REQUIRE(attn_ld_ok(ldp) && attn_ld_ok(ld_dy) && attn_ld_ok(ld_y) && attn_ld_ok(ld_dz) && aligned16(p) && aligned16(dy) && aligned16(y) &&
                  aligned16(dz) && aligned8(adq),
              DRK_EUNSUPPORTED, "attention bwd: rows must be 16-byte aligned");
  const int lpr = attn_lpr(width);
  AttnBwdDstArgs a{rowptr, colidx, p, dy, y, reinterpret_cast<float2*>(adq), slope, ds, dz, (uint32_t)ldp, (uint32_t)ld_dy, (uint32_t)ld_y,
                   (uint32_t)ld_dz, n, width, act, attn_rows_per_block(n, lpr)};
  const int blocks = ceil_div(n, a.rows_per_block);
  cudaStream_t st = as_stream(stream);
  DRK_ATTN_DISPATCH(k_attn_bwd_dst, a, blocks, st);
  return finish_launch("attention bwd (destinations)");
}

int drk_attn_bwd_src(const int32_t* colptr, const int32_t* rowidx, const int32_t* slot_map, const float* dz, int64_t ld_dz, const float* adq,
                     float* ds, const float* att, float* dp, int64_t ld_dp, int32_t n, int32_t width, void* stream) {
  using namespace drk;
  DRK_REQUIRE(n >= 0, DRK_EINVAL, "attention bwd: negative size");
  DRK_REQUIRE(drk_attn_supported(width, 0), DRK_EUNSUPPORTED, "attention bwd: width %d (multiple of 4, <= 128)", width);
  if (n == 0) return DRK_OK;
  DRK_REQUIRE(colptr && rowidx && slot_map && dz && adq && ds && att && dp, DRK_EINVAL, "attention bwd: null pointer");
  DRK_REQUIRE(attn_ld_ok(ld_dz) && attn_ld_ok(ld_dp) && aligned16(dz) && aligned16(dp) && aligned16(att) && aligned8(adq), DRK_EUNSUPPORTED,
              "attention bwd: rows must be 16-byte aligned");
  const int lpr = attn_lpr(width);
  AttnBwdSrcArgs a{colptr, rowidx, slot_map, dz, reinterpret_cast<const float2*>(adq), att, ds, dp, (uint32_t)ld_dz, (uint32_t)ld_dp,
                   n, width, attn_rows_per_block(n, lpr)};
  const int blocks = ceil_div(n, a.rows_per_block);
  cudaStream_t st = as_stream(stream);
  DRK_ATTN_DISPATCH(k_attn_bwd_src, a, blocks, st);
  return finish_launch("attention bwd (sources)");
}

int drk_attn_slot_map(const int32_t* perm, const int32_t* permT, int64_t num_edges, int32_t* inverse_scratch, int32_t* slot_map, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_edges >= 0, DRK_EINVAL, "attention slot map: negative size");
  if (num_edges == 0) return DRK_OK;
  DRK_REQUIRE(perm && permT && inverse_scratch && slot_map, DRK_EINVAL, "attention slot map: null pointer");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = (unsigned)ceil_div<int64_t>(num_edges, 256);
  k_attn_invert_perm<<<blocks, 256, 0, st>>>(perm, num_edges, inverse_scratch);
  k_attn_compose_map<<<blocks, 256, 0, st>>>(permT, inverse_scratch, num_edges, slot_map);
  return finish_launch("attention slot map", 2);
}

size_t drk_attn_edge_grad_workspace_bytes(int32_t fe) { return fe > 0 ? (size_t)drk::kEdgeGradBlocks * fe * sizeof(float) : 0; }

int drk_attn_edge_grad(const float* adq, const float* attr_csr, int64_t ld_attr, int64_t num_edges, int32_t fe, float* g, void* workspace,
                       size_t workspace_bytes, void* stream) {
  using namespace drk;
  DRK_REQUIRE(num_edges >= 0 && fe >= 0 && fe <= kMaxEdgeFeat, DRK_EINVAL, "attention edge grad: bad size");
  if (fe == 0) return DRK_OK;
  DRK_REQUIRE(g && workspace && (num_edges == 0 || (adq && attr_csr)), DRK_EINVAL, "attention edge grad: null pointer");
  DRK_REQUIRE(workspace_bytes >= drk_attn_edge_grad_workspace_bytes(fe), DRK_EWORKSPACE, "attention edge grad: workspace %zu < %zu bytes",
              workspace_bytes, drk_attn_edge_grad_workspace_bytes(fe));
  DRK_REQUIRE(ld_attr >= fe && ld_attr < ((int64_t)1 << 30), DRK_EINVAL, "attention edge grad: bad leading dimension");
  cudaStream_t st = as_stream(stream);
  float* partial = static_cast<float*>(workspace);
  k_attn_edge_grad_partial<<<dim3(kEdgeGradBlocks, fe), kEdgeGradThreads, 0, st>>>(reinterpret_cast<const float2*>(adq), attr_csr, (uint32_t)ld_attr, num_edges, fe,
                                                                                   partial);
  k_attn_edge_grad_final<<<fe, 32, 0, st>>>(partial, kEdgeGradBlocks, fe, g);
  return finish_launch("attention edge grad", 2);
}

}  // extern "C"
