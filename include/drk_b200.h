/*
 * drk_b200.h -- C ABI of the B200-native DeepRank2 message-passing path.
 *
 * One shared library (libdrk_b200.so, built for sm_100a from deeprank-gnn-2_b200/csrc/).
 * The reference (DeepRank2 v3.1.0) has NO native boundary of its own: its GNN path is
 * Python calling torch / torch_scatter / torch_geometric.  Each entry point below
 * therefore cites the reference *call site* (file:line under /root/reference) whose
 * arithmetic it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless marked [host];
 *   - row-major tensors with an explicit leading dimension (elements, not bytes);
 *   - float tensors are fp32, graph indices are int32 after drk_graph_index_build
 *     (the reference's int64 edge_index / batch are the inputs of that call);
 *   - asynchronous on `stream` (a cudaStream_t passed as void*), no allocation, no
 *     ownership transfer, no global state: the caller owns outputs and workspaces;
 *   - return 0 on success, a negative DRK_E* code otherwise; drk_last_error() gives
 *     the thread-local message.  Data-dependent faults (an out-of-range index) cannot
 *     be returned synchronously: they are OR-ed into the int32 `status` word the caller
 *     passes (device memory, may be NULL) -- see DRK_STATUS_*;
 *   - deterministic: no floating-point atomics anywhere; integer atomics only where
 *     the result does not depend on their order.
 */
#ifndef DRK_B200_H
#define DRK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRK_ABI_VERSION 1

#if defined(__GNUC__)
#define DRK_API __attribute__((visibility("default")))
#else
#define DRK_API
#endif

/* return codes */
#define DRK_OK 0
#define DRK_EINVAL (-1)    /* bad argument (null pointer, negative size, unsupported width) */
#define DRK_EWORKSPACE (-2) /* workspace too small */
#define DRK_ECUDA (-3)     /* a CUDA runtime call failed (message has the cudaError string) */
#define DRK_EUNSUPPORTED (-4)

/* bits of the device-side status word */
#define DRK_STATUS_INDEX_RANGE 1 /* an edge endpoint / segment id outside [0, n) */
#define DRK_STATUS_CROSS_GRAPH 2 /* an edge joins two different graphs of the batch */
#define DRK_STATUS_UNSORTED 4    /* `batch` is not non-decreasing */

/* activations / epilogues */
#define DRK_ACT_NONE 0
#define DRK_ACT_RELU 1

/* segment reductions (drk_spmm `reduce`) */
#define DRK_REDUCE_SUM 0        /* torch_scatter.scatter_sum                       (ginet.py:58, vanilla_gnn.py:35) */
#define DRK_REDUCE_MEAN_CLAMP 1 /* torch_scatter.scatter_mean: sum / max(count,1)  (sgat.py:72)                     */
#define DRK_REDUCE_MEAN_NAN 2   /* torch.mean of the gathered rows: 0/0 = NaN      (foutnet.py:56-58)               */

DRK_API int drk_abi_version(void);
DRK_API const char* drk_last_error(void);
/* number of kernels this library has launched from the calling process (all threads); bench.py's gpu_launches */
DRK_API int64_t drk_launch_count(void);
/* clears (and returns) a non-sticky CUDA error another user of the runtime left pending in this process, so that it is not
 * reported against this library's first launch; the host layer calls it once after loading the library on a CUDA box */
DRK_API int drk_runtime_init(void);

/* ------------------------------------------------------------------ graph index
 * Replaces what the reference leaves implicit in `row, col = edge_index` + torch_scatter's
 * index broadcasting (ginet.py:41,58; vanilla_gnn.py:28,35; foutnet.py:57): a destination-sorted
 * CSR and a source-sorted CSC of the batch's edge list, both STABLE (edge ids ascending inside
 * a segment, i.e. torch.sort(stable=True) order == the visiting order of CPU scatter_add_).
 *
 *   edge_index  int64 [2,E] row-major (row 0 = destination "row", row 1 = gathered source "col")
 *   rowptr      int32 [N+1]   colidx  int32 [E] = col[perm]    perm  int32 [E]  (edge ids by destination)
 *   colptr      int32 [N+1]   rowidx  int32 [E] = row[permT]   permT int32 [E]  (edge ids by source)
 *   any of the CSC outputs may be NULL as a group (forward-only use).
 */
DRK_API size_t drk_graph_index_workspace_bytes(int64_t num_edges, int32_t num_nodes);
DRK_API int drk_graph_index_build(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes,
                          int32_t* rowptr, int32_t* colidx, int32_t* perm,
                          int32_t* colptr, int32_t* rowidx, int32_t* permT,
                          int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Generic stable counting sort of one int64 key vector (the `index` argument of
 * torch_scatter.scatter_* when it is NOT a graph edge list: cluster ids, community_pooling.py:209,216).
 *   ptr int32 [num_segments+1], perm int32 [n]. */
DRK_API size_t drk_segment_index_workspace_bytes(int64_t n, int32_t num_segments);
DRK_API int drk_segment_index_build(const int64_t* index, int64_t n, int32_t num_segments,
                            int32_t* ptr, int32_t* perm, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Batch offsets: `ptr` of PyG's Batch.from_data_list from the int64 `batch` vector
 * (trainer.py:541-557 collate; consumed by scatter_mean(x, batch), ginet_nocluster.py:103).
 *   batch int64 [N] non-decreasing, graph_ptr int32 [B+1], batch32 int32 [N] (may be NULL). */
DRK_API int drk_batch_offsets(const int64_t* batch, int32_t num_nodes, int32_t num_graphs,
                      int32_t* graph_ptr, int32_t* batch32, int32_t* status, void* stream);

/* Gather rows of a per-edge tensor into CSR order: out[s,:] = src[perm[s],:]  (edge_attr for vanilla_gnn.py:31). */
DRK_API int drk_gather_rows(const float* src, int64_t ld_src, const int32_t* perm, int64_t n, int32_t width,
                    float* out, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------ dense node projections
 * C[n,:] = act( A[n,:] * op(B) + bias ),  A [N,K], C [N,M];
 *   trans_b = 1: B is [M,K] (nn.Linear weight: C = A B^T)   -- self.fc(x[col]) ginet.py:45, _edge_mlp/_node_mlp vanilla_gnn.py:22-24
 *   trans_b = 0: B is [K,M]                                 -- torch.mm(x, self.wc) foutnet.py:50-51, and dX = dY W in backward
 * If `mask` != NULL the result is multiplied by (mask[n,m] > 0) AFTER the activation (ReLU backward
 * fused into the producer of the gradient).  bias may be NULL.  fp32 FFMA, no tensor cores: the
 * parity bar is rtol 1e-5 (TF32 would be ~1e-3). */
DRK_API int drk_node_linear(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t trans_b,
                    const float* bias, const float* mask, int64_t ld_mask,
                    float* c, int64_t ldc, int64_t n, int32_t k, int32_t m, int32_t act, void* stream);

/* Concat-free Linear on [A | A2]:  C = act(A op(B) + A2 op(B2) + bias) [* mask]; k2 = 0 disables the second pair.
 * node_input = cat([node_features, message_sums]) -> _node_mlp   (vanilla_gnn.py:37-38) without materialising the cat,
 * and dX = dZ Wx + dUV Wab in its backward. */
DRK_API int drk_node_linear2(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t k,
                     const float* a2, int64_t lda2, const float* b2, int64_t ldb2, int32_t k2, int32_t trans_b,
                     const float* bias, const float* mask, int64_t ld_mask,
                     float* c, int64_t ldc, int64_t n, int32_t m, int32_t act, void* stream);

/* dW[M,K] = sum_n dY[n,:]^T X[n,:]   (+ dbias[M] = sum_n dY[n,:] if dbias != NULL).
 * Autograd of the nn.Linear calls above.  Two-stage, fixed-order reduction (deterministic).
 * If `accumulate` != 0 the result is added to the existing contents of dw/dbias. */
DRK_API size_t drk_weight_grad_workspace_bytes(int32_t k, int32_t m);
DRK_API int drk_weight_grad(const float* dy, int64_t ld_dy, const float* x, int64_t ldx, int64_t n, int32_t k, int32_t m,
                    float* dw, int64_t ld_dw, float* dbias, int32_t accumulate,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ segmented gather-reduce ("SpMM")
 * out[i,:] = epilogue( reduce_{s in [ptr[i],ptr[i+1])}  w[s] * src[idx[s],:] ),  i in [0, n_out)
 *   - replaces x[col] -> ... -> scatter_sum(h, row, out=zeros)    ginet.py:45,58 (alpha == 1: w = NULL)
 *   - with reduce = MEAN_NAN: the per-node Python loop of FoutLayer  foutnet.py:56-58
 *   - with w = edge weight in CSR order: edge_attr * (...) of SGAT   sgat.py:68-72
 *   - idx == NULL means idx[s] = s (segments of consecutive rows: scatter_mean(x, batch), ginet_nocluster.py:103)
 * epilogue: act (DRK_ACT_*), then `* (mask[i,:] > 0)` if mask != NULL.  `addend` (may be NULL) is added
 * before the activation (FoutLayer: x Wc + mean(...) + b is formed as addend + mean).
 * One sub-warp per output row, edges visited in CSR order, no atomics. */
DRK_API int drk_spmm(const int32_t* ptr, const int32_t* idx, const float* w,
             const float* src, int64_t ld_src, const float* addend, int64_t ld_addend,
             const float* mask, int64_t ld_mask,
             float* out, int64_t ld_out, int32_t n_out, int32_t width, int32_t reduce, int32_t act, void* stream);

/* Per-graph mean readout  g[b,:] = sum_{i in graph b} x[i,:] / max(n_b, 1)   (one CTA per graph)
 * scatter_mean(data.x, data.batch, dim=0): ginet_nocluster.py:103-104, ginet.py:117-118, vanilla_gnn.py:62, foutnet.py:114. */
/* The same operation for a BLOCK-DIAGONAL adjacency (a collated batch: rows [graph_ptr[g], graph_ptr[g+1]) gather from the same range)
 * whose graphs have up to 3584 nodes: every 16-column slice of a graph's source rows is staged in shared memory once and gathered from
 * there (atom-level graphs).  idx is required; width % 16 == 0; operands 16-byte aligned.  Bit-identical to drk_spmm. */
DRK_API int drk_spmm_tiled_supported(int32_t max_graph_nodes, int32_t width);
DRK_API int drk_spmm_tiled(const int32_t* ptr, const int32_t* idx, const float* w, const float* src, int64_t ld_src, const float* addend, int64_t ld_addend,
                   const float* mask, int64_t ld_mask, float* out, int64_t ld_out, const int32_t* graph_ptr, int32_t num_graphs,
                   int32_t max_graph_nodes, int32_t width, int32_t reduce, int32_t act, void* stream);
DRK_API int drk_segment_mean(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int32_t width,
                     float* out, int64_t ld_out, void* stream);
/* The same with the total number of rows (graph_ptr[num_graphs], known to the caller on the host) as a hint: graphs of >= 512 rows on
 * average are summed by a thread-block cluster of 2 / 4 / 8 CTAs each (partial sums meet through distributed shared memory, rank order). */
DRK_API int drk_segment_mean_rows(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t num_graphs, int64_t num_rows_hint, int32_t width,
                          float* out, int64_t ld_out, void* stream);
/* its backward: dx[i,:] = dg[batch[i],:] / max(n_b,1), optionally * (mask[i,:] > 0) (the ReLU that fed the readout). */
DRK_API int drk_segment_mean_bwd(const float* dg, int64_t ld_dg, const int32_t* graph_ptr, const int32_t* batch32,
                         const float* mask, int64_t ld_mask, int32_t num_nodes, int32_t width,
                         float* dx, int64_t ld_dx, void* stream);

/* ------------------------------------------------------------------ per-edge ReLU messages (VanillaConvolutionalLayer)
 * vanilla_gnn.py:26-35:  messages = relu(_edge_mlp(cat[x_i, x_j, e]));  S = scatter_sum(messages, node0)
 * with the edge MLP split as  U = x Wa^T + b (destination half), V = x Wb^T (source half), C = edge-feature block:
 *   m_e = relu(U[row_e] + V[col_e] + C attr_e),  S[i] = sum_{e in row i} m_e            (message size fixed at 32)
 *   uv [N,64] = U | V;  edge_attr [E,Fe] in ORIGINAL edge order (Fe <= 8);  cmat [32, ld_c]
 *   outputs: s [N,32];  cnt [N,32] = active edges per (node, channel) (dU = dS * cnt);  mask uint32 [E] per edge id.
 *   perm == NULL: edge_attr and mask are indexed by CSR SLOT instead of by original edge id (attributes permuted once per batch, masks
 *   written and read as streams; drk_edge_msg_bwd_src then takes the CSC-slot -> CSR-slot map of drk_attn_slot_map as its `permT`).
 * backward:  drk_edge_msg_bwd_src  dV[j] = sum_{e: col_e = j} dS[row_e] * mask_e   (over the CSC half)
 *            drk_edge_msg_bwd_c    dC[c,k] = sum_e dS[row_e,c] mask_e[c] attr[e,k]  (two-stage fixed-order reduction) */
DRK_API int drk_edge_msg_fwd(const int32_t* rowptr, const int32_t* colidx, const int32_t* perm,
                     const float* uv, int64_t ld_uv, const float* edge_attr, int64_t ld_attr, int32_t num_edge_features,
                     const float* cmat, int64_t ld_c, float* s, int64_t ld_s, float* cnt, uint32_t* mask,
                     int32_t num_nodes, void* stream);
DRK_API int drk_edge_msg_bwd_src(const int32_t* colptr, const int32_t* rowidx, const int32_t* permT,
                         const float* ds, int64_t ld_ds, const uint32_t* mask, float* dv, int64_t ld_dv,
                         int32_t num_nodes, void* stream);
DRK_API size_t drk_edge_msg_bwd_c_workspace_bytes(void);
DRK_API int drk_edge_msg_bwd_c(const int32_t* rowptr, const int32_t* perm, const float* ds, int64_t ld_ds, const uint32_t* mask,
                       const float* edge_attr, int64_t ld_attr, int32_t num_edge_features, float* dc, int64_t ld_dc,
                       int32_t num_nodes, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ VanillaConvolutionalLayer, one kernel per direction
 * Replaces the whole forward / backward of vanilla_gnn.py:26-38 (cat[x_i, x_j, e] -> _edge_mlp -> ReLU -> scatter_sum -> cat[x, S] ->
 * _node_mlp -> ReLU) for a collated batch whose graphs each fit one SM's shared memory: one CTA per graph, V rows and dS rows stay in
 * shared memory, the projections run on the tensor cores with error-compensated TF32 (fp32-level accuracy).  Same tensors as the
 * batch-level route drk_node_linear -> drk_edge_msg_fwd -> drk_node_linear (and interchangeable with it):
 *   x, out [N,F] row-contiguous (F <= 64); s, cnt [N,32]; tf [N,Fe,32] = sum of the ACTIVE edges' attributes per (node, feature, channel)
 *   (dC = sum_i dS[i] * tf[i]); mask uint32 [E] per CSR slot; attr_slots [E,Fe] = edge_attr[perm]; we [32, 2F+Fe], wn [F, F+32] (nn.Linear
 *   weights, row strides ld_*); graph_ptr int32 [G+1] node offsets of the graphs; order int32 [G] = slot -> graph issue order or NULL.
 *   x, s, dout must be 16-byte aligned.  cnt / tf may be NULL in the forward call (inference).
 * Edges that leave their graph raise DRK_STATUS_CROSS_GRAPH, graphs of more than max_graph_nodes nodes DRK_STATUS_INDEX_RANGE.
 * drk_vanilla_layer_supported == 0 -> DRK_EUNSUPPORTED: use the batch-level kernels.
 * backward: dx may be NULL (input without gradient); dbe / dbn may be NULL; the weight gradients are summed per graph and then in graph
 * order (no atomics: bit-reproducible and independent of `order`); workspace = drk_vanilla_layer_bwd_workspace_bytes (one partial per graph). */
DRK_API int drk_vanilla_layer_supported(int32_t num_features, int32_t num_edge_features, int32_t max_graph_nodes);
DRK_API int drk_vanilla_layer_fwd(const float* x, int32_t num_features, const int32_t* rowptr, const int32_t* colidx, const float* attr_slots,
                          int32_t num_edge_features, const int32_t* graph_ptr, const int32_t* order, int32_t num_graphs, int32_t max_graph_nodes,
                          const float* we, int64_t ld_we, const float* be, const float* wn, int64_t ld_wn, const float* bn, float* out, float* s,
                          float* cnt, float* tf, uint32_t* mask, int32_t* status, void* stream);
DRK_API size_t drk_vanilla_layer_bwd_workspace_bytes(int32_t num_features, int32_t num_graphs);
DRK_API int drk_vanilla_layer_bwd(const float* x, const float* s, const float* out, const float* dout, const float* cnt, const float* tf,
                          int32_t num_features, int32_t num_edge_features, const uint32_t* mask, const int32_t* colptr, const int32_t* rowidx,
                          const int32_t* slot_map, const int32_t* graph_ptr, const int32_t* order, int32_t num_graphs, int32_t max_graph_nodes,
                          const float* we, int64_t ld_we, const float* wn, int64_t ld_wn, float* dx, float* dwe, int64_t ld_dwe, float* dbe,
                          float* dwn, int64_t ld_dwn, float* dbn, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ community pooling
 * drk_segment_max: torch_scatter.scatter_max(x, cluster, dim=0) (community_pooling.py:209) and PyG max_pool_x
 *   (ginet.py:103): out[c,:] = max over the segment, arg = element id of the FIRST maximum, empty segment -> (0, n_src).
 *   Segments come from drk_segment_index_build (ptr, perm); perm == NULL means consecutive elements.
 * drk_segment_max_bwd: routes dout[c,f] to dsrc[arg[c,f], f] (dsrc zero-filled by the caller).
 * drk_cluster_offsets: get_preloaded_cluster (community_pooling.py:23-27) without its per-graph host loop:
 *   cluster[i] += sum_{h < batch[i]} (max(cluster in graph h) + 1), in place on the int64 vector;
 *   total_ids (device int64, may be NULL) receives the number of ids after offsetting. */
DRK_API int drk_segment_max(const int32_t* ptr, const int32_t* perm, const float* src, int64_t ld_src, int32_t n_src,
                    int32_t n_seg, int32_t width, float* out, int64_t ld_out, int32_t* arg, void* stream);
DRK_API int drk_segment_max_bwd(const float* dout, int64_t ld_dout, const int32_t* arg, int32_t n_src, int32_t n_seg,
                        int32_t width, float* dsrc, int64_t ld_dsrc, void* stream);
DRK_API size_t drk_cluster_offsets_workspace_bytes(int32_t num_graphs);
DRK_API int drk_cluster_offsets(int64_t* cluster, const int32_t* graph_ptr, const int32_t* batch32, int32_t num_nodes,
                        int32_t num_graphs, int64_t* total_ids, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ fused per-graph GINet (SURVEY 8a rows B, B', C, E)
 * The whole two-branch convolution stack of ginet_nocluster.GINet.forward (ginet_nocluster.py:88-106), one graph per CTA,
 * intermediates in shared memory:
 *   forward : P = x [W1;W1e]^T -> H1 = relu(A P) -> A2 = A H1 -> H2 = relu([A2a W2^T | A2b W2e^T]) -> g[graph] = mean_i H2[i]
 *   backward: dW1s [32,F] (rows 0-15 conv1.fc.weight, 16-31 conv1_ext.fc.weight), dW2a, dW2b [32,16] from dg [B,64]
 * w1s is the row-stack of the two conv1 weights; h1s / a2s [N,32] are the activations saved between the two calls.
 * F <= 64; graphs of more than drk_ginet_fused_max_nodes(F) nodes -> DRK_EUNSUPPORTED (use the unfused kernels).
 * max_graph_edges (<= 0: unknown) sizes the shared-memory staging of each graph's CSR slice; graphs with more edges stream
 * their indices from global memory instead.
 * Edges must stay inside their graph (true for any collated batch); a violation sets DRK_STATUS_CROSS_GRAPH. */
DRK_API int32_t drk_ginet_fused_max_nodes(int32_t num_node_features);
DRK_API int drk_ginet_fused_fwd(const float* x, int64_t ldx, int32_t num_node_features, const int32_t* graph_ptr,
                        const int32_t* rowptr, const int32_t* colidx, const float* w1s, const float* w2a, const float* w2b,
                        float* h1s, float* a2s, float* g, int32_t num_graphs, int32_t max_graph_nodes, int32_t max_graph_edges,
                        int32_t* status, void* stream);
DRK_API size_t drk_ginet_fused_bwd_workspace_bytes(void);
DRK_API int drk_ginet_fused_bwd(const float* x, int64_t ldx, int32_t num_node_features, const int32_t* graph_ptr,
                        const int32_t* colptr, const int32_t* rowidx, const float* w2a, const float* w2b,
                        const float* h1s, const float* a2s, const float* dg, float* dw1s, float* dw2a, float* dw2b,
                        int32_t num_graphs, int32_t max_graph_nodes, int32_t max_graph_edges, int32_t* status,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ per-graph index build for collated batches
 * PyG's collate (torch_geometric Batch.from_data_list; dataset.py:944-948 per graph) concatenates the graphs' edge lists, so
 * the edges of graph g are the contiguous slice [edge_ptr[g], edge_ptr[g+1]) of edge_index.  One CTA per graph builds the
 * same bit-exact, stable CSR / CSC as drk_graph_index_build inside shared memory (per-warp histograms, ordered placement;
 * no atomics, no global sort).  drk_edge_ptr derives edge_ptr from edge_index + graph_ptr when collate did not keep it.
 * An edge whose endpoints leave its graph sets DRK_STATUS_CROSS_GRAPH (the caller falls back to drk_graph_index_build).
 * colptr/rowidx/permT may be NULL together (no CSC); without them perm may be NULL too (a caller that only aggregates -- inference
 * without edge attributes -- skips half of the placement sweep's scattered stores). */
DRK_API int drk_edge_ptr(const int64_t* edge_index, int64_t num_edges, const int32_t* graph_ptr, int32_t num_graphs,
                 int32_t* edge_ptr, void* stream);
DRK_API int drk_graph_index_blocked_supported(int32_t max_graph_nodes, int32_t max_graph_edges);
DRK_API int drk_graph_index_build_blocked(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes,
                                  const int32_t* graph_ptr, const int32_t* edge_ptr, int32_t num_graphs,
                                  int32_t max_graph_nodes, int32_t max_graph_edges, int32_t* rowptr, int32_t* colidx,
                                  int32_t* perm, int32_t* colptr, int32_t* rowidx, int32_t* permT, int32_t* status,
                                  void* stream);

/* ------------------------------------------------------------------ whole GINet step per graph (SURVEY 8a rows B, B', C, D, E, K)
 * One pass of Trainer._epoch's loop body (trainer.py:682-694) for ginet_nocluster.GINet (ginet_nocluster.py:72-111) on a
 * collated batch, as ONE kernel with one CTA per graph + one small finalize kernel:
 *   graph index (CSR + CSC of the graph's edge slice, in shared memory) -> conv1/conv1_ext -> ReLU -> conv2/conv2_ext ->
 *   ReLU -> scatter_mean readout -> fc1 -> ReLU -> dropout -> fc2 -> loss term -> full backward -> per-graph gradient
 *   contributions, summed in graph order by the finalize kernel (bit-reproducible; no floating-point atomics).
 * train == 0: forward only (pred is written, nothing else).
 * loss_kind: DRK_LOSS_MSE            target = float [B, out] (MSELoss on pred.reshape(-1), trainer.py:807-835 regress)
 *            DRK_LOSS_CROSS_ENTROPY  target = int64 [B] class indices (CrossEntropyLoss without class weights)
 * inv_loss_count = 1 / (number of loss elements of the GLOBAL mini-batch): B_global*out for MSE, B_global for CE; with
 *   data-parallel ranks the summed gradients of all ranks then equal the single-process gradient.
 * dropout: keep-mask from Philox4x32-10(seed, state[0], graph, unit); state (device int64[2], zero-initialised by the caller and
 *   private to the calls: [0] = steps done, [1] = scratch of the finalize kernel) is advanced by the finalize kernel so
 *   CUDA-graph replays draw fresh masks.  dropout_p == 0 disables it.  With `peers` the buffer is int64[4] ([2] = exchange epoch).
 * peers (may be NULL = single rank): the per-rank gradients are summed over the ranks inside the finalize kernel (DrkPeers above);
 *   the loss and every gradient output then hold the global values on every rank.
 * adam (may be NULL = gradients only): torch.optim.Adam step (L2 weight decay, no amsgrad) applied by the finalize kernel right
 *   behind the gradient reduction -- same arithmetic as torch's fused CUDA Adam, on torch's own state tensors:
 *   live[0..7] = the tensors of the gradient outputs below, in that order; dead[0..num_dead) = parameters whose gradient is
 *   identically zero (fc_edge_attr / fc_attention: weight decay still moves them); step = the float32 device scalar of each.
 * order (may be NULL = identity): slot -> graph id; CTA b of the G = drk_ginet_step_ctas(B) CTAs processes slots b, b+G, b+2G, ...
 *   A longest-processing-time-first layout (largest graphs first, each to the least loaded CTA; CTAs numbered by decreasing
 *   graph count) gives every CTA about the same total work; the result does not depend on the order (per-graph contributions are summed in graph order).
 * Outputs: pred [B,out]; loss [1]; gradients of conv1.fc.weight / conv1_ext.fc.weight [16,F], conv2.fc.weight /
 *   conv2_ext.fc.weight [32,16], fc1.{weight [128,64], bias [128]}, fc2.{weight [out,128], bias [out]}.
 * The gradients of fc_edge_attr / fc_attention are identically zero in the reference (softmax over a singleton axis,
 * ginet_nocluster.py:48-51) and are not produced here.
 * drk_ginet_step_supported: 1 if graphs of that size fit the shared-memory plan (else use the layer kernels). */
/* edge_layout: what edge_index [2, num_edges] and edge_ptr describe.
 *   DRK_EDGES_DIRECTED          the reference's list of directed edges (dataset.py:944-948 stores every contact twice)
 *   DRK_EDGES_UNDIRECTED_PAIRS  each contact ONCE, as the HDF5 files hold it (`edge_features/_index`, utils/graph.py:210-264);
 *                               the kernel treats pair p of a graph with P pairs as directed edges p = (i,j) and P + p = (j,i),
 *                               i.e. exactly the doubled list, without it ever crossing PCIe or HBM.  max_graph_edges still
 *                               counts DIRECTED edges (2P).
 *   DRK_EDGES_LOCAL_PAIRS16     the same contacts, one 32-bit word each: (i - node0) | (j - node0) << 16 with node0 the first node of
 *                               the pair's graph (a graph of the step kernel has < 65536 nodes).  `edge_index` then points at
 *                               num_edges such words (pass it as the pointer; only its first num_edges*4 bytes are read).  This is
 *                               the form the host collate ships over PCIe: 4 bytes per contact instead of 16. */
#define DRK_EDGES_DIRECTED 0
#define DRK_EDGES_UNDIRECTED_PAIRS 1
#define DRK_EDGES_LOCAL_PAIRS16 2
#define DRK_LOSS_MSE 0
#define DRK_LOSS_CROSS_ENTROPY 1
typedef struct DrkAdamTensor {
  float* param;
  float* exp_avg;
  float* exp_avg_sq;
  float* step; /* device float32 scalar (torch.optim.Adam state["step"] with capturable/fused) */
  int64_t numel;
} DrkAdamTensor;
/* One-shot gradient all-reduce over NVLink peer memory, fused into the finalize kernel (data-parallel ranks, one process per GPU).
 * grad_buf[q]: rank q's symmetric buffer as mapped in THIS process (e.g. torch.distributed._symmetric_memory), 8-byte aligned,
 * 4 * world * drk_ginet_step_exchange_floats() floats (two epochs x one slot array per sending rank x (value, epoch) words), zero
 * before the first call.  Every rank stores its partial sums together with the step's epoch straight into the receivers' memory
 * (one 8-byte store per value) and polls its own.  All ranks must call drk_ginet_step the same number of times; the sums are
 * formed in rank order, so every rank gets bit-identical gradients.  `flags` / `flag_capacity` are reserved (unused). */
typedef struct DrkPeers {
  int32_t world, rank;
  int64_t capacity;      /* floats in each grad_buf */
  int64_t flag_capacity; /* int32 in each flags array */
  float* grad_buf[8];
  int32_t* flags[8];
} DrkPeers;
typedef struct DrkAdam {
  float lr, beta1, beta2, eps, weight_decay;
  int32_t num_dead;
  DrkAdamTensor live[8];
  DrkAdamTensor dead[8];
} DrkAdam;
DRK_API int32_t drk_ginet_step_ctas(int32_t num_graphs);
/* Profiling hook (process-wide, not for production): when `clocks` is non-null every later drk_ginet_step launch of at most
 * `num_slots` graphs writes the SM clock (clock64) at its phase boundaries to clocks[slot * 16 + mark]; null switches it off.
 * Marks: 0 graph start, 1 index built, 2 x staged, 3 projected, 4 H1, 5 A2, 6 conv2+readout, 7 head+loss, 8 dA2, 9 dZ1, 10 Q,
 * 11 dW1 partials, 12 graph done. */
DRK_API int drk_ginet_step_set_phase_clocks(int64_t* clocks, int32_t num_slots);
DRK_API int32_t drk_ginet_step_exchange_floats(int32_t num_node_features, int32_t out_dim);
DRK_API int drk_ginet_step_supported(int32_t num_node_features, int32_t out_dim, int32_t max_graph_nodes, int32_t max_graph_edges);
DRK_API size_t drk_ginet_step_workspace_bytes(int32_t num_node_features, int32_t out_dim, int32_t num_graphs,
                                      int32_t max_graph_nodes, int32_t max_graph_edges);
DRK_API int drk_ginet_step(const float* x, int64_t ldx, int32_t num_node_features, const int64_t* edge_index, int64_t num_edges, int32_t edge_layout,
                   const int32_t* graph_ptr, const int32_t* edge_ptr, const int32_t* order, int32_t outputs_by_slot,
                   int32_t num_graphs, int32_t max_graph_nodes, int32_t max_graph_edges,
                   const float* w1a, const float* w1b, const float* w2a, const float* w2b,
                   const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b, int32_t out_dim,
                   int32_t loss_kind, const void* target, float inv_loss_count,
                   float dropout_p, uint64_t seed, int64_t* state, int32_t train,
                   float* pred, float* loss,
                   float* dw1a, float* dw1b, float* dw2a, float* dw2b,
                   float* dfc1_w, float* dfc1_b, float* dfc2_w, float* dfc2_b,
                   const DrkAdam* adam, const DrkPeers* peers, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ community pooling structure
 * Replaces PyG's consecutive_cluster (torch.unique + scatter_) and pool_edge (relabel, remove_self_loops, coalesce) as called from
 * deeprank2/utils/community_pooling.py:206-219 and torch_geometric's max_pool_x (ginet.py:103,114; foutnet.py:111).  Sizes that depend
 * on the data (distinct clusters, distinct pooled edges) are passed in as CAPACITIES known to the host collate (exact counts or upper
 * bounds); the counts found are written to device scalars, DRK_STATUS_INDEX_RANGE is raised when a capacity is exceeded, and nothing
 * is read back, so the chain can be captured into a CUDA graph.
 *
 * drk_compact_segments: drop the empty segments of a segment index (ptr [num_segments+1], perm from drk_segment_index_build).
 *   rank [num_segments] int64 (nullable): compact id of every non-empty segment, -1 for empty ones (consecutive_cluster's inverse map
 *   is rank[cluster]); ptr_out [capacity+1]: compact offsets (entries beyond the count = total: empty segments); ids_out [capacity]
 *   (nullable): original id of every kept segment; last_out [capacity] int64 (nullable, needs perm): the LAST member of every kept
 *   segment = the largest node index of the cluster (what PyG's perm holds on CPU: last writer wins); count_out: device int32;
 *   workspace: drk_compact_segments_workspace_bytes (per-chunk counts of the two-launch scan).
 * drk_pool_edge_keys: key[e] = pair_ptr[g] + (inv[row_e] - cluster_ptr[g]) * C_g + (inv[col_e] - cluster_ptr[g]) with g =
 *   batch32[row_e], C_g = cluster_ptr[g+1] - cluster_ptr[g]: a dense id of the pooled pair, ascending in (row, col) order; pooled
 *   self loops get junk_key + (e mod DRK_POOL_JUNK_SEGMENTS) with junk_key = pair_ptr[num_graphs]: group the keys with
 *   drk_segment_index_build(key, E, junk_key + DRK_POOL_JUNK_SEGMENTS) and compact the first junk_key segments.  cluster_ptr / pair_ptr: int64 [num_graphs+1], pair_ptr[g+1] - pair_ptr[g] = C_g^2.
 * drk_pool_edge_decode: pooled edge_index [2, capacity] int64 from the dense ids kept by drk_compact_segments. */
#define DRK_POOL_JUNK_SEGMENTS 4096
DRK_API size_t drk_compact_segments_workspace_bytes(int32_t num_segments);
DRK_API int drk_compact_segments(const int32_t* ptr, int32_t num_segments, const int32_t* perm, int64_t* rank, int32_t* ptr_out, int32_t* ids_out,
                         int64_t* last_out, int32_t capacity, int32_t* count_out, int32_t* status, void* workspace, size_t workspace_bytes,
                         void* stream);
DRK_API int drk_pool_edge_keys(const int64_t* edge_index, int64_t num_edges, const int64_t* inv, int32_t num_nodes, const int32_t* batch32,
                       const int64_t* cluster_ptr, const int64_t* pair_ptr, int32_t num_graphs, int64_t junk_key, int64_t* key, int32_t* status,
                       void* stream);
DRK_API int drk_pool_edge_decode(const int32_t* ids, int32_t capacity, const int32_t* count, const int64_t* cluster_ptr, const int64_t* pair_ptr,
                         int32_t num_graphs, int64_t* edge_index_out, void* stream);
/* drk_pool_edge_blocked: the whole pool_edge of a COLLATED batch in one launch, one CTA per graph in shared memory (edge_ptr int32 [G+1]:
 *   the edges of a graph are one slice of edge_index; pooled_edge_ptr int32 [G+1]: where each graph's pooled edges go -- the collate
 *   counts them per graph).  Same outputs as the chain above: pooled_index int64 [2, num_pooled] sorted by (row, col), pooled_attr
 *   [num_pooled, Fe] = attributes of merged edges added in ascending edge id (NULL: no attributes).  A graph whose clustering yields
 *   another number of pooled edges than pooled_edge_ptr says raises DRK_STATUS_INDEX_RANGE, an edge between graphs DRK_STATUS_CROSS_GRAPH.
 *   drk_pool_edge_blocked_supported == 0 (more than 255 clusters per graph, or lists that do not fit a CTA) -> use the chain above. */
/* drk_consecutive_blocked: consecutive_cluster of a COLLATED batch in one launch, one CTA per graph (node_ptr int32 [G+1]: the nodes of a
 *   graph are contiguous; cluster ids globally unique after drk_cluster_offsets, so a graph's ids form a range of at most max_graph_ids
 *   values; cluster_ptr int64 [G+1]: first NEW id of every graph, counted by the collate).  Outputs as drk_segment_index_build +
 *   drk_compact_segments give them: inv [N] (new id of every node), last [C] (largest node index of every cluster), ptr_c [C+1] / perm
 *   [N] (nodes grouped by new id, ascending).  A graph with another number of distinct ids than cluster_ptr says, or ids spanning more
 *   than max_graph_ids, raises DRK_STATUS_INDEX_RANGE; capacity = the number of clusters last / ptr_c were allocated for (nothing is
 *   written beyond it). */
DRK_API int drk_consecutive_blocked_supported(int32_t max_graph_nodes, int32_t max_graph_ids);
DRK_API int drk_consecutive_blocked(const int64_t* cluster, int32_t num_nodes, const int32_t* node_ptr, const int64_t* cluster_ptr, int32_t num_graphs,
                            int32_t max_graph_nodes, int32_t max_graph_ids, int32_t capacity, int64_t* inv, int64_t* last, int32_t* ptr_c,
                            int32_t* perm, int32_t* status, void* stream);
DRK_API int drk_pool_edge_blocked_supported(int32_t max_graph_clusters, int32_t max_graph_edges);
DRK_API int drk_pool_edge_blocked(const int64_t* edge_index, int64_t num_edges, const int32_t* edge_ptr, const int64_t* inv, int32_t num_nodes,
                          const int64_t* cluster_ptr, const int32_t* pooled_edge_ptr, int32_t num_graphs, int32_t max_graph_clusters,
                          int32_t max_graph_edges, const float* edge_attr, int64_t ld_attr, int32_t num_edge_features, int64_t* pooled_index,
                          int64_t num_pooled, float* pooled_attr, int32_t* status, void* stream);

/* ------------------------------------------------------------------ GINet attention with a segment softmax per destination
 * The operator the reference's GINetConvLayer sets up (ginet.py:45-52: logit = leaky_relu(fc_attention([fc(x)[row], fc(x)[col],
 * fc_edge_attr(edge_attr)]))) normalised over the edges of each destination node -- BASELINE.json north_star, SURVEY 8f rank 4.
 * The reference itself normalises over a singleton axis (softmax(alpha, dim=1), ginet.py:54 => alpha == 1), which is what
 * drk_spmm / drk_ginet_step implement; this is the opt-in `attention="segment_softmax"` mode of the layer.
 *   p [n,width] = fc(x);  s [n,2] = (a_r.p[i], a_c.p[i]) with fc_attention.weight = [a_r | a_c | a_e];  u [fe] = We^T a_e;
 *   attr_csr [E,fe] = edge_attr in CSR-slot order (drk_gather_rows with perm, once per batch);  rowptr/colidx of drk_graph_index_build.
 *   z[i,:] = act( sum_{e in seg i} alpha_e p[col_e,:] ),  alpha = softmax_seg(leaky_relu(s_r[i] + s_c[col_e] + u.attr_e, slope))
 *   adq [E,2] (CSR-slot order): column 0 = alpha_e with the sign bit set where the logit is <= 0 (written by fwd), column 1 = the
 *   gradient of the logit (written by bwd_dst);  logit_scratch [E] is scratch between the two sweeps of the forward kernel.
 * One sub-warp per destination, edges in CSR order, no atomics.  width % 4 == 0, width <= 128, fe <= 32. */
DRK_API int drk_attn_supported(int32_t width, int32_t fe);
DRK_API int drk_attn_fwd(const int32_t* rowptr, const int32_t* colidx,
                 const float* p, int64_t ldp, const float* s, const float* attr_csr, int64_t ld_attr, int32_t fe,
                 const float* u, float slope, float* z, int64_t ldz, float* adq, float* logit_scratch,
                 int32_t n, int32_t width, int32_t act, void* stream);
/* Backward, destinations (CSR): dz = dy [* (y > 0) if act == RELU, written to `dz`]; adq[s,1] = dq_e = alpha_e lrelu'(q_e) (dz[i].p[col_e] - dz[i].y[i])
 * (the gradient of the logit before the leaky ReLU); ds[i,0] = sum_{e in seg i} dq_e. */
DRK_API int drk_attn_bwd_dst(const int32_t* rowptr, const int32_t* colidx,
                     const float* p, int64_t ldp, const float* dy, int64_t ld_dy, const float* y, int64_t ld_y,
                     float* adq, float slope, float* ds, float* dz, int64_t ld_dz,
                     int32_t n, int32_t width, int32_t act, void* stream);
/* CSC slot -> CSR slot of the same edge (perm/permT of drk_graph_index_build; inverse_scratch int32 [E]); once per batch. */
DRK_API int drk_attn_slot_map(const int32_t* perm, const int32_t* permT, int64_t num_edges, int32_t* inverse_scratch, int32_t* slot_map, void* stream);
/* Backward, sources (CSC): ds[j,1] = sum_{e: col_e = j} dq_e;  dp[j,:] = sum_{e: col_e = j} alpha_e dz[row_e,:] + ds[j,0] a_r + ds[j,1] a_c
 * (att = [a_r | a_c], 2*width floats).  The remaining gradients are dense contractions of these outputs:
 * d fc.weight = dp^T x (drk_weight_grad), d[a_r|a_c] = ds^T p, and g below -> d a_e = We g, d fc_edge_attr.weight = a_e (x) g. */
DRK_API int drk_attn_bwd_src(const int32_t* colptr, const int32_t* rowidx, const int32_t* slot_map,
                     const float* dz, int64_t ld_dz, const float* adq, float* ds, const float* att,
                     float* dp, int64_t ld_dp, int32_t n, int32_t width, void* stream);
/* g[k] = sum_s adq[s,1] attr_csr[s,k]  (two-stage fixed-order reduction). */
DRK_API size_t drk_attn_edge_grad_workspace_bytes(int32_t fe);
DRK_API int drk_attn_edge_grad(const float* adq, const float* attr_csr, int64_t ld_attr, int64_t num_edges, int32_t fe,
                       float* g, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRK_B200_H */
