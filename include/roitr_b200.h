/*
 * libroitr_b200 — C ABI of the B200-native RoITr forward hot path.
 *
 * Plain pointers and sizes only (no torch types). Every pointer is a DEVICE pointer unless it says "host".
 * Every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately:
 *   0            success (launch enqueued; asynchronous faults surface at the caller's next synchronisation)
 *   < 0          argument / unsupported-shape error, nothing was enqueued (ROITR_ERR_*)
 *   > 0          the cudaError_t of a failed launch or runtime call
 * roitr_last_error() returns a thread-local human-readable message for the last non-zero return.
 *
 * Layout conventions are the reference's: row-major contiguous f32 (n,3) coordinates, int32
 * `offset` arrays holding CUMULATIVE SEGMENT ENDS per batch element (a batch = several clouds concatenated),
 * int32 indices.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef ROITR_B200_H
#define ROITR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* roitr_last_error(void);
int roitr_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * Native point ops (replace the two live functions of the pybind module `pointops_cuda`,
 * cpp_wrappers/pointops/src/pointops_api.cpp:13-14).
 * ---------------------------------------------------------------------------------------------------------- */

/*
 * Drop-in for  knnquery_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)
 * (cpp_wrappers/pointops/src/knnquery/knnquery_cuda_kernel.h:13; kernel .cu:65-108), plus `b` and a stream.
 * Exact brute-force kNN of each query inside its own segment, ascending by squared distance computed as
 * fma(dz,dz,fma(dx,dx,dy*dy)) with d = query - ref (the reference's SASS). Unfilled slots (segment shorter than
 * nsample) keep (idx = segment start, dist2 = 1e10). Exact-distance ties are ordered by ascending index.
 * nsample <= 32.
 */
int roitr_knnquery(int b, int m, int nsample, const float* xyz, const float* new_xyz, const int* offset,
                   const int* new_offset, int* idx, float* dist2, void* stream);
/* Same, with n_total = rows of xyz (= offset[b-1]) supplied by the caller: no device->host read, no stream sync. */
int roitr_knnquery_n(int b, int m, int nsample, int n_total, const float* xyz, const float* new_xyz, const int* offset,
                     const int* new_offset, int* idx, float* dist2, void* stream);

/*
 * Fused kNN + point-pair features: the body of
 *   queryandgroup(nsample, p, n_p, ..., return_idx=True)  (pointops.py:79-92: kNN(nsample+1), drop column 0)
 *   -> p[group_idx], n[group_idx] -> calc_ppf_gpu          (lib/utils.py:358-389)
 * as used by TransitionDown / RIPointTransformerLayer (model/model.py:31-41,75-77), in ONE kernel.
 *   k_out   neighbours kept per query;  drop_first in {0,1}: leading (nearest) columns discarded (1 = reference)
 *   idx     (m,k_out) int32
 *   dist    (m,k_out) f32 or NULL: sqrt of the squared distance (what pointops.knnquery returns, pointops.py:43)
 *   ppf     (m,k_out,4) f32 or NULL: [ |d|, ang(n1,d)/pi, ang(n2,d)/pi, ang(n1,n2)/pi ], d = p_j - p_i;
 *           requires normals / new_normals.
 * k_out + drop_first <= 32.
 */
int roitr_knn_ppf(int b, int m, int k_out, int drop_first, const float* xyz, const float* normals,
                  const float* new_xyz, const float* new_normals, const int* offset, const int* new_offset, int* idx,
                  float* dist, float* ppf, void* stream);
int roitr_knn_ppf_n(int b, int m, int k_out, int drop_first, int n_total, const float* xyz, const float* normals,
                    const float* new_xyz, const float* new_normals, const int* offset, const int* new_offset, int* idx,
                    float* dist, float* ppf, void* stream);

/*
 * Grid-accelerated form of roitr_knn_ppf_n: identical results (same distance arithmetic on every candidate, same
 * (distance, index) order, same exact tie replay), but each query only visits the cells of a growing cube of a uniform
 * grid over its segment until the k-th distance is provably smaller than anything outside the cube.
 *   roitr_knn_grid_workspace_bytes(b, n)  bytes of device workspace for a reference set of b segments / n points
 *   roitr_knn_grid_build(...)             bounding boxes, cell size, counting sort into cell order (4 small kernels);
 *                                         the workspace can serve any number of queries against the same reference set
 *   roitr_knn_ppf_grid(...)               as roitr_knn_ppf_n, with the workspace (256-byte aligned)
 */
long long roitr_knn_grid_workspace_bytes(int b, int n);
/* roitr_knn_grid_build with an explicit average number of points per grid cell over the bounding box (default build: 0.5,
 * measured best for every query type of the forward on B200, scripts/tune_grid.py). */
int roitr_knn_grid_build_target(int b, int n, const float* xyz, const int* offset, float target_per_cell, void* workspace,
                                void* stream);
/* Surface normals of a (segmented) cloud: per point the `knn` (9, 17 or 33) nearest points of its own segment, itself
 * included, their covariance and the eigenvector of its smallest eigenvalue (fp64), oriented towards `view_point`
 * (3 HOST floats). Replaces Open3D estimate_normals(KDTreeSearchParamKNN(knn=33)) + normal_redirect
 * (dataset/tdmatch.py:120-127, dataset/common.py:312-320). `workspace` = roitr_knn_grid_build over the same cloud. */
int roitr_estimate_normals(int b, int n, int knn, const float* xyz, const int* offset, const void* workspace,
                           const float* view_point, float* normals, void* stream);
long long roitr_knn_grid_sorted_offset(int b);
int roitr_knn_grid_build(int b, int n, const float* xyz, const int* offset, void* workspace, void* stream);
int roitr_knn_ppf_grid(int b, int m, int k_out, int drop_first, int n_total, const float* xyz, const float* normals,
                       const float* new_xyz, const float* new_normals, const int* offset, const int* new_offset,
                       const void* workspace, int* idx, float* dist, float* ppf, void* stream);
/* Same, with the visiting order of the queries taken from the QUERY set's own grid (`query_workspace`, built by
 * roitr_knn_grid_build over new_xyz / new_offset; NULL = natural order, or the reference grid itself for self queries):
 * one thread per query, neighbouring threads in neighbouring cells. Results are identical to roitr_knn_ppf_grid. */
int roitr_knn_ppf_grid_q(int b, int m, int k_out, int drop_first, int n_total, const float* xyz, const float* normals,
                         const float* new_xyz, const float* new_normals, const int* offset, const int* new_offset,
                         const void* workspace, const void* query_workspace, int* idx, float* dist, float* ppf,
                         void* stream);

/*
 * Drop-in for  furthestsampling_cuda_launcher(b, n, xyz, offset, new_offset, tmp, idx)
 * (cpp_wrappers/pointops/src/sampling/sampling_cuda_kernel.h:13; kernel .cu:14-129).
 * Iterative furthest point sampling per segment; first sample = first point of the segment; identical distance
 * arithmetic; identical tie order (which depends on the reference's block size = f(n_max), src/cuda_utils.h:11-14):
 *   n_max > 0   one block size for the whole batch, derived from n_max exactly like the reference launcher;
 *   n_max == 0  per-segment block size derived from each segment's own length (equals b independent b=1 calls).
 * `tmp` (the reference's global running-distance scratch) is accepted for signature parity and ignored: running
 * distances live in registers of a thread-block cluster. new_xyz (sum m,3) optional (NULL): sampled coordinates.
 * cluster_hint: 0 = auto, else CTAs per cloud (1,2,4,8). Segments of up to 65536 points.
 */
int roitr_furthestsampling(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp,
                           int* idx, float* new_xyz, int cluster_hint, void* stream);

/* As above with the maximum segment length supplied by the caller (no device->host read). */
int roitr_furthestsampling_cfg(int b, int n_max, int n_seg_max, const float* xyz, const int* offset,
                               const int* new_offset, int* idx, float* new_xyz, int cluster_hint, void* stream);

/*
 * Inverse-distance interpolation after a k-NN: body of pointops.interpolation
 * (cpp_wrappers/pointops/functions/pointops.py:174-182) plus the skip add of TransitionUp (model/model.py:116).
 *   idx,dist (n,k) from roitr_knn_ppf(dist = sqrt distances); feat (m,c); base (n,c) or NULL; out (n,c). k <= 8.
 */
int roitr_interpolate(int n, int c, int k, const int* idx, const float* dist, const float* feat, const float* base,
                      float* out, void* stream);

/*
 * out[i,:] = src[index[i],:] for rows of c floats; index int32 or int64 (index_is_i64). Rows whose index equals
 * pad_row (>= 0) read as zeros: the zero row the reference appends before index_select (lib/utils.py:403-425,
 * model/RIGA_v2.py:86-89,138-142). pad_row = -1 disables.
 */
int roitr_gather_rows(long long rows, int c, const void* index, int index_is_i64, const float* src, float* out,
                      long long pad_row, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense layers and row epilogues (nn.Linear / nn.LayerNorm / F.relu / F.normalize chains of model/model.py:90-117,
 * 131-142, model/transformer/attention.py:166-170,317-319, geoattention.py:177-192, model/RIGA_v2.py:64-68).
 * ---------------------------------------------------------------------------------------------------------- */

/* C[M,N] = (A [+ a_add])[rows,:K] W[:N,:K]^T + bias (+ReLU). Row-major; lda/ldw/ldc leading dimensions so operands may be
 * column slices of wider buffers; a_index (int32, M entries) gathers rows of A / a_add; bias, a_add, a_index may be NULL.
 * fp32 FFMA arithmetic. */
int roitr_linear(int M, int N, int K, const float* A, const float* a_add, int lda, const int* a_index, const float* W,
                 int ldw, const float* bias, float* C, int ldc, int relu, void* stream);

/* The tensor-core dense layer: tcgen05.mma kind::tf32 with 3xTF32 split precision (x = hi + lo exactly; hi*hi + hi*lo + lo*hi
 * accumulated in fp32 TMEM; relative error ~2^-21, fp32-grade); persistent, warp-specialised (A loader warps, bulk-TMA weight producer, single-
 * thread tcgen05 MMA issuer, epilogue warps; two TMEM accumulators so the epilogue of a tile overlaps the next tile).
 * wpack = the weight pre-split into TF32 hi/lo and pre-swizzled by roitr_b200.engine.pack_linear_tc:
 * [ceil(N/bn)][ceil(K/32)][hi|lo][bn*32] floats, zero padded, bn in {64,128}. Same contract as roitr_linear otherwise.
 * Kernel selection (same arithmetic, same results): plain or gathered (a_index) 16-byte aligned rows -> the streaming kernel
 * (cp.async raw ring + split pass); N > 128 with >= 74 row tiles -> the row-group kernel (activations loaded and split once per
 * pair of 128-column weight tiles); a_add or unaligned inputs -> the coupled-ring kernel with register-staged loaders. */
int roitr_linear_tc_packed(int M, int N, int K, const float* A, const float* a_add, int lda, const int* a_index,
                           const float* wpack, int bn, const float* bias, float* C, int ldc, int relu, void* stream);

/* Launch configuration of the streaming dense-layer kernel for the calls issued from now on (a process-wide host-side mode;
 * baked into a CUDA graph at capture): 0 = deep rings, one CTA per SM (215 KB shared memory, 57 K registers); 3 = light
 * footprint (one operand stage, two raw stages, <= 64 registers: ~115 KB, 20 K registers) that shares an SM with the CTAs of
 * other streams' kernels - the engine sets it around the level-1 layers it issues while the FPS clusters are resident.
 * Results are identical in both. */
int roitr_set_linear_config(int config);
/* Dense layer with the row epilogue fused (one kernel instead of roitr_linear_tc_packed + roitr_row_epilogue):
 *   C = act( LayerNorm_N( A W^T + bias + res_pre[res_pre_index] ) * gamma + beta + res_post ),  act = ReLU if relu
 * (LocalRPEAttentionLayer output: attention.py:317-319; RIPointTransformerBlock: model/model.py:139-141; TransitionUp:
 * model/model.py:103-105). N must be a multiple of 32 and fit one weight tile (N <= bn) or, with bn = 128, two (N <= 256: both
 * halves of the row sit side by side in TMEM); res_pre / res_post rows have pitch ldr; res_pre_index (int32, optional) gathers
 * res_pre rows. eps = 1e-5. */
int roitr_linear_ln_tc_packed(int M, int N, int K, const float* A, int lda, const float* wpack, int bn, const float* bias,
                              const float* gamma, const float* beta, const float* res_pre, const int* res_pre_index,
                              const float* res_post, int ldr, int relu, float* C, int ldc, void* stream);

/* out = [L2norm] [ReLU] ( [LayerNorm_{gamma,beta,eps=1e-5}] (x + res_pre[res_pre_index]) + res_post ), one row of C<=1024
 * floats per warp. mode bits: 1 LayerNorm, 2 ReLU, 4 x / max(|x|_2, 1e-12). Any of the residuals may be NULL. */
int roitr_row_epilogue(int M, int C, const float* x, const float* res_pre, const int* res_pre_index, const float* gamma,
                       const float* beta, const float* res_post, float* out, int mode, void* stream);

/* TransitionUp head (model/model.py:101-112): per-segment column mean, and cat(x, repeat(g[segment])) -> (M, 2C). */
int roitr_segment_mean(int b, int C, const float* x, const int* offset, float* out, void* stream);
int roitr_concat_segment(int M, int C, int b, const float* x, const float* g, const int* offset, float* out,
                         void* stream);

/*
 * Fused local PPF attention: LocalRPEMultiHeadAttention.forward (model/transformer/attention.py:166-200) between the
 * q/k/v projections and the output linear, with the positional projections folded into (C,4)+(C) maps:
 *   Ap = W_p W_e, cp = W_p b_e + b_p, Avp = W_vp W_e, cvp = W_vp b_e + b_vp  (W_e,b_e = PPFStructualEmbedding.proj).
 * q/k/v: (n,C) row-major with leading dims ldq/ldk/ldv (may alias one (n,3C) buffer); node_idx (m) int32 or NULL;
 * group_idx (m,knb) int32; ppf (m,knb,4); out (m,C). heads = 4, C in {64,128,256,512}, knb in {8,16}.
 */
int roitr_local_attention(int m, int C, int heads, int knb, const float* q, int ldq, const float* k, int ldk,
                          const float* v, int ldv, const int* node_idx, const int* group_idx, const float* ppf,
                          const float* Ap, const float* cp, const float* Avp, const float* cvp, float* out, void* stream);
/* Same, visiting the queries in the order of `order_xyzi` ((m,4) floats: the cell-sorted (x, y, z, index) array of the
 * query set's grid, see roitr_knn_grid_sorted_offset; NULL = natural order). Results are identical; neighbouring
 * queries share neighbours, so the K/V gathers hit L1. */
int roitr_local_attention_ordered(int m, int C, int heads, int knb, const float* q, int ldq, const float* k, int ldk,
                                  const float* v, int ldv, const int* node_idx, const int* group_idx, const float* ppf,
                                  const float* Ap, const float* cp, const float* Avp, const float* cvp,
                                  const float* order_xyzi, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Global geometric transformer (model/transformer/positional_encoding.py:94-154, geoattention.py:43-136).
 * ---------------------------------------------------------------------------------------------------------- */

/* k nearest superpoints of every superpoint on dist = sqrt(clamp(x2 - 2xy + y2, 0)) (topk(k+1)[1:], :120-124). k = 3. */
int roitr_geo_knn(int N, int k, const float* pts, int* nn, void* stream);
/* `batch` clouds of N superpoints stacked along dim 0 (pts (batch*N,3), nn (batch*N,k) with cloud-local indices). */
int roitr_geo_knn_batched(int batch, int N, int k, const float* pts, int* nn, void* stream);

/* E (N,N,C) = proj_d(sinusoid(dist/sigma_d)) + max_{r<3} proj_a(sinusoid(angle_r * 180/(sigma_a*pi)))
 * (GeometricStructureEmbedding.forward, :139-154) as the GEMM the reference runs, on the tensor cores (tcgen05.mma
 * kind::tf32, 3xTF32 split precision, sinusoid A operand generated in-kernel, four accumulators per tile in TMEM). This is
 * the FALLBACK of roitr_geo_embedding_table: the engine takes it when the table's interpolation error bound cannot be met
 * for the loaded weights (roitr_b200.engine.build_geo_tables). C multiple of 128.
 * wpack: W_d and W_a pre-split into TF32 hi/lo and pre-swizzled into 32 KB SWIZZLE_128B blocks,
 * layout [mat(d,a)][C/128][C/32][hi|lo][4096] floats (roitr_b200.engine.pack_tf32_sw128); fetched by bulk TMA. */
int roitr_geo_embedding_tc(int N, int C, const float* pts, const int* nn3, const float* wpack, const float* bd,
                           const float* ba, const float* div_term, float sigma_d, float sigma_a, float* E, void* stream);
/* `batch` clouds per launch (blockIdx.z): pts (batch*N,3), nn3 (batch*N,3) cloud-local, E (batch,N,N,C). */
int roitr_geo_embedding_tc_batched(int batch, int N, int C, const float* pts, const int* nn3, const float* wpack,
                                   const float* bd, const float* ba, const float* div_term, float sigma_d, float sigma_a,
                                   float* E, void* stream);

/* Same E from weight-derived tables: F_d(t) = W_d s(t) + b_d and F_a(t) = W_a s(t) + b_a depend on the weights only, so
 * they are sampled once per weight load on a uniform grid of step h = 1/inv_h (a power of two) in fp64
 * (roitr_b200.engine.build_geo_tables) and evaluated per (n,m) scalar by 4-point Lagrange interpolation (error bound
 * (3/128) h^4 max|d4F/dt4|, chosen <= 2.5e-7 at pack time). tab_a [C/64][rows_a][64], tab_d [C/64][rows_d][64], row r
 * holds t = (r - 1) h. Scalars beyond the tables (and NaN / Inf) are evaluated directly from Wd / Wa. pts (batch*N,3), nn3 (batch*N,3) cloud-local, E (batch,N,N,C). C multiple of 64. */
long long roitr_geo_table_smem_rows(void);
int roitr_geo_embedding_table(int batch, int N, int C, const float* pts, const int* nn3, const float* tab_a, int rows_a,
                              const float* tab_d, int rows_d, float inv_h, const float* Wd, const float* bd,
                              const float* Wa, const float* ba, const float* div_term, float sigma_d, float sigma_a,
                              float* E, void* stream);

/* Batched dense contraction on the tensor cores (tcgen05 3xTF32, one CTA per 128 x {64,128} tile, 3-4 CTAs per SM, operands split on the fly) for
 * operands that are activations: for every (o, i) in batch_outer x batch_inner
 *     C_oi[M,N] = A_oi[M,K] W_oi[N,K]^T,   X_oi = X + o * sX_o + i * sX_i  (element strides)
 * w_transposed != 0: W_oi is given as (K, >= N) row-major with leading dimension ldw (W_oi[n][k] = Wt[k * ldw + n]).
 * Used for the global transformer's Q K^T (per cloud and head) and P V (V read transposed) - geoattention.py:50,62,107,128. */
int roitr_gemm_tc_batched(int batch_outer, int batch_inner, int M, int N, int K, const float* A, int lda, long long sA_o,
                          long long sA_i, const float* W, int ldw, long long sW_o, long long sW_i, int w_transposed,
                          float* C, int ldc, long long sC_o, long long sC_i, void* stream);

/* RPE self-attention between Q K^T and P V (geoattention.py:107-133), one streaming pass over E:
 *   S = (qk + gq_h . E[n,m] + q_h . b_p,h) / sqrt(c);  P (batch,H,N,N) = softmax_m(S);
 *   G (batch*N,H,C) = sum_m softmax_m(S without the diagonal)[m] E[n,m,:]   (the caller maps G through proj_vp per head).
 * qk (batch,H,N,N) raw q.k; q: (batch*N rows, C) view with leading dim ldq and cloud stride q_bs; gq (batch*N,H,C) folded
 * positional queries; bp = proj_p.bias. heads = 4, C in {256,512}. */
int roitr_geo_self_scores(int batch, int N, int C, int heads, const float* qk, const float* q, int ldq, long long q_bs,
                          const float* E, const float* gq, const float* bp, float* P, float* G, void* stream);
/* Same with an explicit row pitch for gq (gq may be a column slice of a wider buffer, e.g. of the folded [q|k|v|gq] layer). */
int roitr_geo_self_scores_ld(int batch, int N, int C, int heads, const float* qk, const float* q, int ldq, long long q_bs,
                             const float* E, const float* gq, int ldgq, const float* bp, float* P, float* G, void* stream);

/* out[row,:] = softmax(qk[row,:] / scale_div) over M entries per row (MultiHeadAttention, geoattention.py:50-60). */
int roitr_softmax_rows(long long rows, int M, const float* qk, float scale_div, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Matching head (lib/utils.py:428-471, model/modules.py:10-72,135-178,216-324, model/RIGA_v2.py:150-173).
 * ---------------------------------------------------------------------------------------------------------- */

/* Every entry point of the matching head takes a leading PAIR dimension B: `B` equally sized problems whose arrays lie
 * back to back (pair-major, contiguous), served by ONE launch per kernel (blockIdx.y / .z = pair). B = 1 is the single-pair
 * forward of lib/tester.py:53; the batched runner passes its 16 pairs at once. */

/* point_to_node_partition (lib/utils.py:448-463) for B clouds: pts (B,N,3), nodes (B,M,3) -> owner (B,N) int32 = nearest node
 * (matmul-form distance, clamp 1e-12, first minimum), dmin (B,N), count (B,M), and per node its `limit` nearest OWN points
 * sorted by (distance, index): knn_idx (B,M,limit) int32 (pad = N), knn_mask (B,M,limit), node_mask (B,M) bytes. A counting
 * sort by owner (buckets in `workspace`, roitr_point_to_node_workspace_bytes) lets a node touch only the points it owns. */
long long roitr_point_to_node_workspace_bytes(int B, int N, int M);
int roitr_point_to_node_batched(int B, int N, int M, int limit, const float* pts, const float* nodes, int* owner, float* dmin,
                                int* count, void* workspace, int* knn_idx, unsigned char* knn_mask, unsigned char* node_mask,
                                void* stream);

/* torch.nonzero replacement: ascending flat indices of the non-zero bytes, at most `capacity` written, the true total in
 * *count (device). chunk_scratch needs roitr_compact_scratch_ints(n) ints. The batched form compacts B segments of n flags
 * each: out (B, max(capacity,1)), count (B), chunk_scratch B * roitr_compact_scratch_ints(n) ints. */
int roitr_compact_flags(long long n, const unsigned char* flags, int* chunk_scratch, int* out, int capacity, int* count,
                        void* stream);
int roitr_compact_flags_batched(int B, long long n, const unsigned char* flags, int* chunk_scratch, int* out, int capacity,
                                int* count, void* stream);
long long roitr_compact_scratch_ints(long long n);

/* CoarseMatching.forward (model/modules.py:141-178) on (B,Mr,C) x (B,Ms,C) L2-normalised descriptors with validity masks
 * (B,Mr) / (B,Ms): exp(-sqdist), dual normalisation (fused into the selection kernel), flat top-k (sorted descending, ties
 * by ascending flat index). xy = ref @ src^T (B,Mr,Ms) is supplied by the caller. work: B * (Mr*Ms + 2*(Mr+Ms)) floats.
 * Outputs (B,k) padded; out_count (B) = min(k, #valid pairs). k <= 1024. */
int roitr_coarse_matching_batched(int B, int Mr, int Ms, int C, int k, int dual, const float* ref_feats, const float* src_feats,
                                  const unsigned char* ref_mask, const unsigned char* src_mask, const float* xy, float* work,
                                  int* out_ref, int* out_src, float* out_score, int* out_count, void* stream);

/* AdaptiveSuperPointMatching.forward (model/modules.py:81-123, 4DMatch head): sim = sqrt(clamp(2 - 2 a.b, 1e-12)) on valid
 * superpoints; all pairs with sim <= threshold in row-major (torch.nonzero) order, or the min_num smallest (ascending) when
 * fewer than min(min_num, #valid pairs) qualify; scores exp(-sim). Both candidate lists are built on the device and
 * selected from the device-side counts (no host sync). xy = a @ b^T (B,Ma,Mb). work: B * (2*Ma*Mb floats + Ma*Mb bytes
 * rounded up to floats); iwork: B * (cap + 3*min_num + 4 + roitr_compact_scratch_ints(Ma*Mb)) ints. Outputs (B,cap) padded
 * (cap <= Ma*Mb). min_num <= 1024. */
int roitr_coarse_matching_adaptive_batched(int B, int Ma, int Mb, int min_num, float threshold, const unsigned char* a_mask,
                                           const unsigned char* b_mask, const float* xy, float* work, int* iwork, int cap,
                                           int* out_a, int* out_b, float* out_score, int* out_count, void* stream);

/* One CTA per superpoint correspondence p < corr_count[pair]: gather the two 64-point patches' descriptors, scores =
 * Ft Fs^T / sqrt(C) (RIGA_v2.py:150-152; 32-term chunks summed into a compensated total), LearnableLogOptimalTransport
 * (modules.py:28-68, num_iter Sinkhorn iterations in registers / shared memory) -> scores (B,Pmax,65,65); then
 * FineMatching.compute_correspondence_matrix (modules.py:242-274) -> flags (B,Pmax,64,64) bytes. tgt_feat (B,Nt,C),
 * src_feat (B,Ns,C), tgt_knn / tgt_kmask (B,Mt,64), src_knn / src_kmask (B,Ms,64), corr_t / corr_s (B,Pmax), corr_count (B). */
int roitr_fine_matching_batched(int B, int Pmax, int Mt, int Ms, int Nt, int Ns, int C, const float* tgt_feat,
                                const float* src_feat, const int* tgt_knn, const int* src_knn, const unsigned char* tgt_kmask,
                                const unsigned char* src_kmask, const int* corr_t, const int* corr_s, const int* corr_count,
                                const float* alpha, int num_iter, int topk, int mutual, float threshold, float* scores,
                                unsigned char* flags, void* stream);

/* FineMatching.extract_correspondences (modules.py:276-283) from the compacted flat (p,row,col) indices: flat (B,capacity),
 * count (B), tgt_pts (B,Nt1,3) / src_pts (B,Ns1,3) -> out_t / out_s (B,capacity,3), out_score (B,capacity). */
int roitr_fine_gather_batched(int B, int capacity, int Pmax, int Mt, int Ms, int Nt1, int Ns1, const int* flat, const int* count,
                              const float* scores, const int* corr_t, const int* corr_s, const int* tgt_knn, const int* src_knn,
                              const float* tgt_pts, const float* src_pts, float* out_t, float* out_s, float* out_score,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Ground-truth bookkeeping that RIGA_v2.forward runs in test mode too (lib/utils.py:474-614).
 * ---------------------------------------------------------------------------------------------------------- */

/* out (N+1,3): [pts ; zero row] (RIGA_v2.py:86-87), optionally mapped through p R^T + t (lib/utils.py:505). */
int roitr_pad_transform(int N, const float* pts, const float* rot, const float* trans, float* out, void* stream);
/* B equally sized clouds per launch: pts (B*N,3), rot (B,3,3) or NULL, trans (B,3), out (B*(N+1),3). */
int roitr_pad_transform_batched(int B, int N, const float* pts, const float* rot, const float* trans, float* out,
                                void* stream);

/* get_node_occlusion_score tail (lib/utils.py:511-526) given the 1-NN distances of the padded clouds: knn / kmask (B,M,K),
 * nmask (B,M), nn_dist (B,nn_stride) -> occ (B,M). */
int roitr_node_occlusion_batched(int B, int M, int K, int nn_stride, const int* knn, const unsigned char* kmask,
                                 const unsigned char* nmask, const float* nn_dist, float thr, float* occ, void* stream);

/* get_node_correspondences (lib/utils.py:562-606): dense (B,Mr,Ms) overlap ratios and >0 flags (compact them with
 * roitr_compact_flags_batched, then roitr_corr_gather_batched). rot (B,3,3), trans (B,3). work: B * 4*(Mr+Ms) floats. K = 64. */
int roitr_node_overlaps_batched(int B, int Mr, int Ms, int K, int Nr, int Nsrc, const float* ref_nodes, const float* src_nodes,
                                const int* ref_knn, const int* src_knn, const unsigned char* ref_kmask,
                                const unsigned char* src_kmask, const unsigned char* ref_mask, const unsigned char* src_mask,
                                const float* ref_pts, const float* src_pts, const float* rot, const float* trans, float radius,
                                float* work, float* overlap, unsigned char* flag, void* stream);
int roitr_corr_gather_batched(int B, int capacity, int Mr, int Ms, const int* flat, const int* count, const float* overlap,
                              long long* out_idx, float* out_ov, void* stream);

/* Weighted Procrustes (lib/utils.py:159-218): per batch item the rigid transform (R (3,3) row-major, t (3)) that maps the
 * src points (batch, n, 3) onto the tgt points under the weights (batch, n) (NULL = all ones; weights below weight_thresh
 * count as zero; centroids use w / (sum w + eps)). R is always a proper rotation (the reference's sign det(V U^T) fix). */
int roitr_weighted_procrustes(int batch, int n, const float* src, const float* tgt, const float* weights,
                              float weight_thresh, float eps, float* R, float* t, void* stream);

/* Correspondence RANSAC: replaces ransac_pose_estimation_correspondences (registration/benchmark_utils.py:165-209, called
 * from registration/evaluate_registration_c2f.py:88) = Open3D registration_ransac_based_on_correspondence with
 * TransformationEstimationPointToPoint(False), ransac_n = 3, checkers EdgeLength(edge_similarity = 0.9) and
 * Distance(distance_threshold), RANSACConvergenceCriteria(iterations = 50000, confidence clamped to 1 => no early exit).
 * `pairs` problems in one launch: src / tgt (total, 3) f32 hold the matched points of every pair back to back (correspondence i
 * = row i of both), offset (pairs,) int32 cumulative ends, max_corr >= the largest per-pair count (<= 8192).
 * Hypothesis j of pair p samples rows hash(seed, p, 3 j + {0,1,2}) with replacement (deterministic; oracle/ransac_ref.py
 * shares the hash). Outputs per pair: transform (4,4) f64 row-major mapping src onto tgt (identity when no hypothesis
 * passes), fitness = inliers / n, rmse over the inliers, best_itr (-1 if none); ties go to the lower iteration.
 * workspace: roitr_ransac_workspace_bytes(pairs, iterations) bytes. All arithmetic is fp64, like Open3D. */
long long roitr_ransac_workspace_bytes(int pairs, int iterations);
int roitr_ransac_correspondences(int pairs, int iterations, const float* src, const float* tgt, const int* offset,
                                 int max_corr, double distance_threshold, double edge_similarity, unsigned seed,
                                 void* workspace, double* transform, double* fitness, double* rmse, int* best_itr,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROITR_B200_H */
