/*
 * libmingraph_b200 — C ABI of the B200 (sm_100a) graph block of MinGraph-UNet.
 *
 * The reference (agent-charon/MinGraph-UNet) has no FFI: its boundary is the
 * Python nn.Module contract of the classes cited below (paths relative to the
 * reference root).  This header is what a binding of those classes needs: plain
 * device pointers, explicit sizes, a dtype enum and the caller's cudaStream_t.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - nothing here allocates, synchronises or touches the default stream:
 *     work is enqueued on `stream` (a cudaStream_t passed as void*), so calls are
 *     CUDA-graph capturable;
 *   - return value 0 = enqueued; <0 = error (MG_ERR_*), text via mg_last_error();
 *   - node features are row-major (N, D); graphs are CSR over int32:
 *       "in"  CSR: rowptr[j]..rowptr[j+1] lists the SOURCES of edges into j,
 *       "out" CSR: rowptr[i]..rowptr[i+1] lists the TARGETS of edges out of i,
 *     both in ascending COO edge id (the order torch's CPU scatter_add_ visits
 *     them), so segment sums round like the reference's;
 *   - batched calls are block-diagonal: image b owns nodes [b*N, (b+1)*N).
 */
#ifndef MINGRAPH_B200_H_
#define MINGRAPH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MG_API __attribute__((visibility("default")))
#else
#define MG_API
#endif

#define MG_VERSION 100            /* 0.1.0 */

#define MG_OK 0
#define MG_ERR_INVALID (-1)       /* bad argument / unsupported shape */
#define MG_ERR_CUDA (-2)          /* launch failed (sticky CUDA error text in mg_last_error) */
#define MG_ERR_UNSUPPORTED (-3)

#define MG_F32 0
#define MG_BF16 1
#define MG_I32 2                  /* label arrays only (mg_feature_consistency_loss) */
#define MG_I64 3

typedef void* mg_stream_t;        /* cudaStream_t */

MG_API int mg_version(void);
MG_API const char* mg_last_error(void);
/* Number of kernels this library has launched from the calling process (for bench.py's
 * gpu_launches claim). */
MG_API int64_t mg_launch_count(void);

/* ---- graph construction --------------------------------------------------------------
 * PatchGraphConstructor.construct_patch_graph — preprocessing/graph_construction/
 * patch_graph_construction.py:49-102.  4-connected patch grid, both directions, the
 * reference's emission order (per node, row-major: right pair then down pair). */
MG_API int64_t mg_grid_num_edges(int Hp, int Wp);      /* 2*(Hp*(Wp-1)+Wp*(Hp-1)) */
/* edge_index (2, B*E) int64: row 0 sources, row 1 targets; image b's ids are offset by
 * b*Hp*Wp when offset_nodes != 0 (B=1 reproduces the reference tensor bit for bit). */
MG_API int mg_grid_edge_index(int Hp, int Wp, int B, int offset_nodes, int64_t* edge_index, mg_stream_t stream);
/* Closed-form CSR of the same graph (block-diagonal over B images).  The grid is
 * symmetric, so the arrays serve as both the "in" and the "out" CSR; neighbour order is
 * up, left, right, down = ascending COO edge id in both views.  eid_in/eid_out (nullable)
 * receive the per-image COO edge id of each slot in the in/out view. */
MG_API int mg_grid_csr(int Hp, int Wp, int B, int32_t* rowptr, int32_t* col, int32_t* eid_in, int32_t* eid_out,
                mg_stream_t stream);
/* Complete digraph on K regions per image — scripts/train_end_to_end.py:376-380
 * (triu pairs (s,t) then the reversed pairs). */
MG_API int mg_complete_edge_index(int K, int B, int offset_nodes, int64_t* edge_index, mg_stream_t stream);
MG_API int mg_complete_csr(int K, int B, int32_t* rowptr, int32_t* col, mg_stream_t stream);
/* Arbitrary caller-supplied edge_index (2,E) int64 -> stable CSR by target (by_target=1)
 * or by source (0).  work: at least mg_csr_work_bytes(N,E) bytes.  status (1 int, nullable):
 * set non-zero if an index is outside [0,N) (the reference raises IndexError). */
MG_API int64_t mg_csr_work_bytes(int N, int64_t E);
MG_API int mg_csr_from_coo(const int64_t* edge_index, int64_t E, int N, int by_target, int32_t* rowptr, int32_t* col,
                    int32_t* eid, void* work, int32_t* status, mg_stream_t stream);

/* ---- kNN graph (north_star kernel (2); NOT in the reference, see csrc/knn.cu) -------------------------------
 * For every node j the k nearest OTHER nodes of its graph under squared Euclidean distance accumulated in fp32 in
 * feature order (rounded sub, mul, add; no FMA), ties to the lower id, sorted by (distance, id).  Outputs (each
 * nullable): edge_index (2, N*k) int64 (row 0 = neighbour = source, row 1 = j = target), in-CSR rowptr (N+1) /
 * col (N*k) int32, dist (N*k) f32.  nodes_per_graph > 0: block-diagonal batch of equal-size graphs.  k <= 32. */
MG_API int mg_knn_graph(const float* x, int N, int D, int k, int nodes_per_graph, int64_t* edge_index, int32_t* rowptr,
                 int32_t* col, float* dist, mg_stream_t stream);

/* ---- pooling ---------------------------------------------------------------------------
 * Patch mean pooling (B,C,Hf,Wf) -> (B, Hp*Wp, C): the documented intent of
 * PatchGraphConstructor.get_patch_features_from_unet_encoder (patch_graph_construction.py:
 * 104-136, raises NotImplementedError) expressed as image_to_patches(x).mean((2,3))
 * (:27-47; scripts/graph_refinement.py:78,98,103).  Right/bottom zero padding counts in the
 * mean: the divisor is always ph*pw. */
MG_API int mg_pool_patches(const void* x, int x_dtype, int B, int C, int Hf, int Wf, int ph, int pw, void* out,
                    int out_dtype, mg_stream_t stream);
/* Region mean pool — train_end_to_end.py:368-373: out[b,k,:] = mean(h[b, labels[b]==k, :]),
 * zero for empty regions.  h (B,N,D) f32, labels (B,N) int32, out (B,K,D) f32,
 * counts (B,K) int32 (nullable); work: mg_segment_work_bytes() bytes.  Two deterministic stages (per-chunk partials,
 * then a fixed-order sum), no atomics. */
MG_API int64_t mg_segment_work_bytes(int B, int N, int D, int K);
MG_API int mg_segment_mean(const float* h, const int32_t* labels, int B, int N, int D, int K, float* out,
                    int32_t* counts, void* work, mg_stream_t stream);

/* ---- graph attention ---------------------------------------------------------------------
 * MultiHeadGATLayer.forward (eval) — model/gat/graph_attention.py:40-118,150-160.
 *   x (N,in) f32|bf16; W (heads,F,in) f32 = heads.{h}.W.weight; a (heads,2F) f32 = heads.{h}.a.weight
 *   out (N, concat ? heads*F : F) f32|bf16
 *   nodes_per_graph > 0: the batch holds N/nodes_per_graph independent graphs and the softmax
 *   shift (torch.max(e), :86) is taken per graph; 0 = one graph.
 *   work: mg_gat_work_bytes() bytes of scratch (scores, per-graph max, spill of the
 *   aggregated features when the weights do not fit in shared memory).
 *   save_den (N,heads) f32 and save_z (N,heads,in) f32 are optional (nullable) outputs kept
 *   for mg_gat_backward.
 *   dropout_p > 0 (training): attention dropout (:97) with a counter-based mask keyed on (seed, in-CSR slot,
 *   head); statistically equivalent to torch's, not the same stream.  seed_dev (nullable DEVICE pointer): its value is
 *   added to `seed` when the kernel runs, so a captured CUDA graph draws a fresh mask per replay.
 * Empty graphs (E == 0) are rejected with MG_ERR_INVALID like the reference's RuntimeError. */
MG_API int64_t mg_gat_work_bytes(int N, int in_dim, int out_dim, int heads, int num_graphs);
/* 1 if mg_gat_forward runs the node transform of this shape on the tensor pipe (tcgen05 tf32 MMA, csrc/gat_tc.cu:
 * bf16 node features, inference, heads in {1,2,4}, in in {32,64,128,256} with heads*in <= 256, F % 16 == 0,
 * 2*heads*F <= 512, N >= 4096), else 0 (FP32-pipe kernels). */
MG_API int mg_gat_uses_tensor_pipe(int N, int in_dim, int out_dim, int heads, int concat, int x_dtype, int out_dtype);
MG_API int mg_gat_forward(const void* x, int x_dtype, const int32_t* rowptr, const int32_t* col, int N, int64_t E,
                   const float* W, const float* a, int in_dim, int out_dim, int heads, int concat, float slope,
                   int nodes_per_graph, float dropout_p, uint64_t seed, const uint64_t* seed_dev, void* out, int out_dtype,
                   void* work, float* save_den, float* save_z, mg_stream_t stream);

/* MultiHeadGATLayer backward (the reference relies on autograd: IndexBackward / ScatterAddBackward / MmBackward
 * over graph_attention.py:53-118).  Needs both CSR views and, per out-CSR slot, the in-CSR slot of the same edge
 * (mg_edge_slot_map from the eid arrays of mg_grid_csr / mg_csr_from_coo; work = E int32).
 *   den (N,heads), z (N,heads,in): saved by mg_gat_forward;  grad_out (N, out_w) f32
 *   -> grad_x (N,in) f32, grad_W (heads,F,in), grad_a (heads,2F).  dropout_p/seed must equal the forward's.
 * Deterministic (owner-computes passes by target and by source, fixed-order split reductions). */
MG_API int mg_edge_slot_map(const int32_t* eid_in, const int32_t* eid_out, int64_t E, int32_t* work, int32_t* slot_out2in,
                     mg_stream_t stream);
MG_API int64_t mg_gat_backward_work_bytes(int N, int64_t E, int in_dim, int out_dim, int heads, int num_graphs);
MG_API int mg_gat_backward(const void* x, int x_dtype, const int32_t* rowptr_in, const int32_t* col_in,
                    const int32_t* rowptr_out, const int32_t* col_out, const int32_t* slot_out2in, int N, int64_t E,
                    const float* W, const float* a, int in_dim, int out_dim, int heads, int concat, float slope,
                    int nodes_per_graph, float dropout_p, uint64_t seed, const uint64_t* seed_dev, const float* den,
                    const float* z, const float* grad_out, float* grad_x, float* grad_W, float* grad_a, void* work,
                    const void* fwd_work /* nullable: the work buffer mg_gat_forward used for the same call (its
                                            attention scalars are then reused instead of recomputed) */,
                    mg_stream_t stream);

/* Row softmax + argmax of the predictor logits — mincut_refinement.py:193 and
 * train_end_to_end.py:356.  logits (N,K) f32 -> S (N,K) f32, labels (N) int32 (first max). */
MG_API int mg_softmax_argmax(const float* logits, int N, int K, float* S, int32_t* labels, mg_stream_t stream);

/* ---- normalized cut ------------------------------------------------------------------------
 * MinCutRefinement.compute_edge_weights_for_ncut — mincut_refinement.py:30-52: w (E) f32 in
 * COO order for an int64 edge_index. */
MG_API int mg_ncut_edge_weights(const float* h, int N, int D, const int64_t* edge_index, int64_t E, float* w,
                         mg_stream_t stream);
/* MinCutRefinement.normalized_cut_loss — :55-160, per graph.  Uses the OUT CSR (degree is
 * summed by source, :96).  loss (G) f32; stats (G,2K) nullable: per-graph [assoc | cut] kept for the backward;
 * work: mg_ncut_work_bytes(). */
MG_API int64_t mg_ncut_work_bytes(int N, int K, int num_graphs);
MG_API int mg_ncut_loss(const float* h, const float* S, const int32_t* rowptr_out, const int32_t* col_out, int N, int D,
                 int K, int nodes_per_graph, float* loss, float* stats, void* work, mg_stream_t stream);
/* Backward of mg_ncut_loss w.r.t. h and S.  stats (G,2K) = [assoc | cut] as written by mg_ncut_loss (nullable
 * there); grad_loss (G); -> grad_h (N,D), grad_S (N,K).  Owner-computes over both CSR views, no atomics. */
MG_API int mg_ncut_backward(const float* h, const float* S, const int32_t* rowptr_out, const int32_t* col_out,
                     const int32_t* rowptr_in, const int32_t* col_in, int N, int D, int K, int nodes_per_graph,
                     const float* stats, const float* grad_loss, float* grad_h, float* grad_S, mg_stream_t stream);

/* ---- un-pool ---------------------------------------------------------------------------------
 * train_end_to_end.py:403-421: f_patch = table[labels]; (N,D)->(D,Hp,Wp); nearest up-sampling to
 * (D,H,W) with torch's index rule.  table (B,K,D) f32; labels (B,Hp*Wp) int32 or NULL (then
 * K == Hp*Wp and the table is the per-patch feature matrix).  out[b] starts at
 * out + b*out_batch_stride elements and holds (D,H,W) contiguous, so the result can be written
 * straight into a channel slice of a fusion buffer. */
MG_API int mg_unpool_nearest(const float* table, const int32_t* labels, int B, int K, int D, int Hp, int Wp, int H, int W,
                      void* out, int out_dtype, int64_t out_batch_stride, mg_stream_t stream);

/* ---- backward of the glue stages (training) -------------------------------------------------------
 * The reference gets these from autograd (UpsampleNearest2DBackward + IndexBackward for
 * train_end_to_end.py:403-421, IndexBackward/MeanBackward for :368-373, SoftmaxBackward for
 * mincut_refinement.py:193).  All deterministic, no atomics.
 * mg_unpool_nearest_backward: grad_out (B,D,H,W) planes f32|bf16 (image b at grad_out + b*grad_batch_stride
 * elements) -> grad_table (B,K,D) f32 = sum of the gradient over the pixels whose patch carries label k
 * (labels NULL: K == Hp*Wp, per-patch sums).  work: mg_unpool_backward_work_bytes() bytes. */
MG_API int64_t mg_unpool_backward_work_bytes(int B, int D, int Hp, int Wp, int K);
MG_API int mg_unpool_nearest_backward(const void* grad_out, int grad_dtype, int64_t grad_batch_stride, const int32_t* labels,
                               int B, int K, int D, int Hp, int Wp, int H, int W, void* work, float* grad_table,
                               mg_stream_t stream);
/* grad_h (B,N,D) (+)= grad_out[b, labels[b,n], :] / counts[b, labels[b,n]]  (counts from mg_segment_mean). */
MG_API int mg_segment_mean_backward(const float* grad_out, const int32_t* labels, const int32_t* counts, int B, int N, int D,
                             int K, int accumulate, float* grad_h, mg_stream_t stream);
/* grad_logits = S * (grad_S - rowsum(grad_S * S)). */
MG_API int mg_softmax_backward(const float* S, const float* grad_S, int N, int K, float* grad_logits, mg_stream_t stream);

/* ---- fused per-image block (forward) ------------------------------------------------------------
 * The per-image loop of scripts/train_end_to_end.py:329-389 for a batch of B images in ONE launch:
 * patch GAT -> predictor GAT -> softmax/argmax -> N-cut loss -> region mean-pool -> region GAT on the
 * 4-connected Hp x Wp patch grid (one thread-block cluster per image).  Weights are those of the
 * three 1-layer GATNetworks (heads.{h}.W.weight stacked: W1 (H1,D,in), W2 (H2,K,D), W3 (H3,D,D);
 * heads.{h}.a.weight stacked: a1 (H1,2D), a2 (H2,2K), a3 (H3,2D)); mg_block_prepare re-arranges them
 * into `prep` (mg_block_prep_floats() floats, 16-byte aligned) once per weight version.
 *   x (B,N,in) f32|bf16 -> h (B,N,D), S (B,N,K), labels (B,N) int32, loss (B), region_in (B,K,D, nullable),
 *   region_out (B,K,D); q_work: B*N*(2*H2 + H2*K) floats of scratch.
 * mg_block_supported() != 0 tells whether the shape fits the fused kernel (in <= 64, D in {32,64,128},
 * heads <= 4, K <= 8, shared-memory budget); otherwise compose the stand-alone entry points above. */
MG_API int64_t mg_block_prep_floats(int in_dim, int D, int H1, int H2, int H3, int K);
MG_API int mg_block_supported(int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3, int K);
MG_API int mg_block_prepare(const float* W1, const float* a1, const float* W2, const float* a2, const float* W3,
                            const float* a3, int in_dim, int D, int H1, int H2, int H3, int K, float* prep,
                            mg_stream_t stream);
MG_API int mg_block_forward(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3,
                            int K, float slope1, float slope2, float slope3, const float* prep, float* h, float* q_work,
                            float* S, int32_t* labels, float* loss, float* region_in, float* region_out,
                            mg_stream_t stream);

/* ---- losses on tensors the block already holds (scope row f4) -----------------------------------
 * FeatureConsistencyLoss.forward — model/unet/feature_loss.py:103-123, called per image at
 * scripts/train_end_to_end.py:344 with the patch-GAT output as f_graph:
 *   dist_sq = sum_d (f_unet - f_graph)^2 ; dist = sqrt(dist_sq + 1e-8)
 *   loss    = mean_b sum_n [ y*dist_sq + (1-y)*max(0, margin - dist)^2 ]
 * f_unet, f_graph (B,N,D) f32|bf16; y (B,N) of dtype MG_F32 | MG_BF16 | MG_I32 | MG_I64 (the reference
 * calls .float() on it).  work: mg_feature_loss_work_bytes(B,N) bytes.  per_image (B, nullable)
 * receives the per-image sums, loss (1) their mean.  Deterministic (fixed-order reductions). */
MG_API int64_t mg_feature_loss_work_bytes(int B, int N);
MG_API int mg_feature_consistency_loss(const void* f_unet, int fu_dtype, const void* f_graph, int fg_dtype, const void* y,
                                       int y_dtype, int B, int N, int D, float margin, void* work, float* per_image,
                                       float* loss, mg_stream_t stream);
/* grad_f_unet = grad_loss/B * (2y - (1-y)*2*hinge/dist) * (f_unet - f_graph); grad_f_graph = -grad_f_unet
 * (either may be NULL); grad_loss is a DEVICE scalar. */
MG_API int mg_feature_consistency_loss_backward(const void* f_unet, int fu_dtype, const void* f_graph, int fg_dtype,
                                                const void* y, int y_dtype, int B, int N, int D, float margin,
                                                const float* grad_loss, float* grad_f_unet, float* grad_f_graph,
                                                mg_stream_t stream);
/* TVLoss.forward — scripts/train_end_to_end.py:73-89:
 *   weight * ( sum (x[:,:,1:,:]-x[:,:,:-1,:])^2 / ((H-1)*W) + sum (x[:,:,:,1:]-x[:,:,:,:-1])^2 / (H*(W-1)) ) / B
 * x (B,C,H,W) f32|bf16, read once.  out3 (3 floats): {loss, h_tv, w_tv}.  H == 1 or W == 1 gives NaN like
 * the reference (0/0).  work: mg_tv_loss_work_bytes() bytes. */
MG_API int64_t mg_tv_loss_work_bytes(int dtype, int B, int C, int H, int W);
MG_API int mg_tv_loss(const void* x, int dtype, int B, int C, int H, int W, float weight, void* work, float* out3,
                      mg_stream_t stream);
MG_API int mg_tv_loss_backward(const void* x, int dtype, int B, int C, int H, int W, float weight, const float* grad_loss,
                               float* grad_x, mg_stream_t stream);

/* ---- fusion: per-region embeddings -> dense per-pixel map (next caller after the block, scope row f1) ----
 * FeatureFusion.forward, per-region branch — model/fusion_detection/feature_fusion.py:81-140:
 *   out[b, d, y, x] = table[map[b,y,x], d] if 0 <= map[b,y,x] < R, else 0   (valid_mask :119; zeros :85)
 * table (R, D) f32 = f_g; map (B,H,W) MG_I32 | MG_I64 = region_to_pixel_map (the reference calls .long() on it);
 * out: image b at out + b*out_batch_stride elements, (D,H,W) contiguous planes f32|bf16 — i.e. the channel slice
 * [C_u : C_u + D] of the fused (B, C_u + D, H, W) buffer, so Concat(F_u, F_g) (:143) needs no extra pass.
 * First run on a B200 in round 2: parity green, 0.80-0.87 of the measured HBM peak (profiles/r2_fusion_gather.md). */
MG_API int mg_region_map_gather(const float* table, int R, int D, const void* map, int map_dtype, int B, int H, int W,
                                void* out, int out_dtype, int64_t out_batch_stride, mg_stream_t stream);

/* ---- peer-memory exchange of the small per-image outputs (multi-GPU, scope row (e)) ----------------
 * The batch shards by image (scripts/train_end_to_end.py:300-425 builds one graph per image), so the only
 * per-step exchange is an all-gather of l_partition | region_features | hard_labels (74 KB per rank at cfg 2).  It is
 * FUSED into the block kernel: mg_block_forward_push is mg_block_forward plus, while it computes, stores of that
 * payload into slice `rank` of every rank's gathered buffer over NVLink peer mappings, and a sequence number published
 * in every rank's flag array by the last CTA (system-scope release).  No collective call, nothing waits for a peer.
 *
 * Buffers.  Every rank allocates one gathered buffer and one flag array with mg_peer_mem_alloc (cudaMalloc, zeroed,
 * plus a 64-byte CUDA IPC handle), ranks exchange the handles (any host channel) and map each other's allocations
 * with mg_peer_mem_open.  mg_peer_mem_* are host calls for set-up / tear-down: they allocate and synchronise, unlike
 * everything else in this header.  Gathered buffer layout (floats): [2 parity][slots][world][slice], one slice =
 * loss [B] | region_out [B][K][D] | labels [B][N] (int32 bits), the layout of mg_block_forward's small outputs.
 * Flags: [slots][world] uint32.
 *
 * mg_peer_out_t describes one (slot, rank): peer_bufs_dev / peer_flags_dev are DEVICE arrays of `world` pointers
 * (entry p = rank p's allocation as mapped on this GPU; entry `rank` = the local one); slice_offset = (slot * world +
 * rank) * slice floats; parity_stride = slots * world * slice; flag_index = slot * world + rank; seq, done = one
 * uint32 each on this device, zeroed once (seq counts the slot's completed steps and selects the parity half).
 * Step s of a slot lands in parity half s & 1 with flag value s + 1; see csrc/peer_exchange.cu for why that needs no
 * acknowledgement.  The launch must cover the slot's whole batch (B = images of the slice).
 *
 * mg_peer_wait: consumer side — the stream continues once my_flags[first_flag + r] >= *seq for every source rank r
 * (enqueue it after the slot's mg_block_forward_push on the same stream); bounded spin of a few seconds, on expiry
 * status[0] = 1 (nullable). */
typedef struct mg_peer_out {
  const void* const* peer_bufs_dev;
  const void* const* peer_flags_dev;
  int32_t world, rank;
  int64_t slice_offset, parity_stride, flag_index;
  uint32_t* seq;
  uint32_t* done;
} mg_peer_out_t;

MG_API int mg_block_forward_push(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2,
                                 int H3, int K, float slope1, float slope2, float slope3, const float* prep, float* h,
                                 float* q_work, float* S, int32_t* labels, float* loss, float* region_in,
                                 float* region_out, const mg_peer_out_t* peer, mg_stream_t stream);
/* mg_block_forward_ex: mg_block_forward_push plus FeatureConsistencyLoss folded into the kernel (scope row f4 — the loss
 * of model/unet/feature_loss.py:103-123 between the U-Net patch features and the patch-GAT output h, evaluated while
 * the rows of h are still in registers: no second pass over h).  floss (nullable): f_unet (B,N,D) f32, y (B,N) f32
 * patch labels, margin; loss_per_image (B) receives the per-image sums over the patches (the reference's value is
 * their batch mean, :124). */
typedef struct mg_block_feature_loss {
  const float* f_unet;
  const float* y;
  float margin;
  float* loss_per_image;
} mg_block_feature_loss_t;

MG_API int mg_block_forward_ex(const void* x, int x_dtype, int B, int Hp, int Wp, int in_dim, int D, int H1, int H2, int H3,
                               int K, float slope1, float slope2, float slope3, const float* prep, float* h,
                               float* q_work, float* S, int32_t* labels, float* loss, float* region_in,
                               float* region_out, const mg_peer_out_t* peer, const mg_block_feature_loss_t* floss,
                               mg_stream_t stream);
MG_API int mg_peer_wait(const uint32_t* my_flags, int64_t first_flag, int world, const uint32_t* seq, int32_t* status,
                        mg_stream_t stream);
MG_API int mg_peer_mem_alloc(int64_t nbytes, void** ptr_host, unsigned char* handle64_host);
MG_API int mg_peer_mem_open(const unsigned char* handle64_host, void** ptr_host);
MG_API int mg_peer_mem_close(void* ptr);
MG_API int mg_peer_mem_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* MINGRAPH_B200_H_ */
