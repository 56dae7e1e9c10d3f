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

#ifdef __cplusplus
}
#endif
#endif /* ROITR_B200_H */
