// Ground-truth bookkeeping that RIGA_v2.forward also runs in test mode (sm_100a): superpoint occlusion scores and
// ground-truth superpoint correspondences from (rot, trans). Observable outputs of the forward
// (gt_node_corr_indices, gt_node_corr_overlaps, gt_{tgt,src}_node_occ; lib/tester.py:64-65 saves the latter).
//
// Replaces get_node_occlusion_score (lib/utils.py:474-527) and get_node_correspondences (lib/utils.py:530-614).
#include <math_constants.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

struct Rt { float r[9]; float t[3]; };

// torch.matmul(p, rot.T) + trans.T for one point: fma chain over k like the matmul, then the add
__device__ __forceinline__ void apply_rt(const float* R, const float* t, float x, float y, float z, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = __fadd_rn(fmaf(z, R[3 * i + 2], fmaf(y, R[3 * i + 1], __fmul_rn(x, R[3 * i]))), t[i]);
}

// out (N+1,3): rows < N = (transformed) points, row N = the zero pad row of RIGA_v2.py:86-87 (transformed too, since the
// reference transforms the padded array, lib/utils.py:505).
__global__ void pad_transform_kernel(int N, const float* __restrict__ pts, const float* __restrict__ rot,
                                     const float* __restrict__ trans, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    {   // blockIdx.y = pair of a batch: pts (B*N,3), rot (B,3,3), trans (B,3), out (B*(N+1),3)
        const size_t b = blockIdx.y;
        pts += b * (size_t)N * 3; out += b * (size_t)(N + 1) * 3;
        if (rot) { rot += b * 9; trans += b * 3; }
    }
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < N) { x = __ldg(pts + 3 * i); y = __ldg(pts + 3 * i + 1); z = __ldg(pts + 3 * i + 2); }
    float o[3] = {x, y, z};
    if (rot) {
        float R[9], t[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = __ldg(trans + k);
        apply_rt(R, t, x, y, z, o);
    }
    out[3 * (size_t)i] = o[0]; out[3 * (size_t)i + 1] = o[1]; out[3 * (size_t)i + 2] = o[2];
}

// occ[node] = node_mask * sum_j (nn_dist[knn[node,j]] < thr) * kmask / (sum_j kmask + 1e-10)     (lib/utils.py:511-526)
__global__ void node_occ_kernel(int M, int K, const int* __restrict__ knn, const unsigned char* __restrict__ kmask,
                                const unsigned char* __restrict__ nmask, const float* __restrict__ nn_dist, int nn_stride,
                                float thr, float* __restrict__ occ) {
    const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (node >= M) return;
    {   // blockIdx.y = pair: knn / kmask (B,M,K), nmask / occ (B,M), nn_dist (B, nn_stride)
        const size_t b = blockIdx.y;
        knn += b * M * K; kmask += b * M * K; nmask += b * M; occ += b * M; nn_dist += b * nn_stride;
    }
    float hit = 0.f, cnt = 0.f;
    for (int j = lane; j < K; j += 32) {
        const float mk = kmask[(size_t)node * K + j] ? 1.f : 0.f;
        const float ov = (__ldg(nn_dist + __ldg(knn + (size_t)node * K + j)) < thr) ? 1.f : 0.f;
        hit += ov * mk;
        cnt += mk;
    }
    hit = warp_sum(hit); cnt = warp_sum(cnt);
    if (lane == 0) occ[node] = (hit / (cnt + 1e-10f)) * (nmask[node] ? 1.f : 0.f);
}

// per node: (optionally transformed) node centre and the radius of its patch: max_j kmask * |p_j - node|  (:573-578)
__global__ void node_radius_kernel(int M, int K, int N, const float* __restrict__ nodes, const int* __restrict__ knn,
                                   const unsigned char* __restrict__ kmask, const float* __restrict__ pts,
                                   const float* __restrict__ rot, const float* __restrict__ trans,
                                   float* __restrict__ nodes_out, float* __restrict__ radius) {
    const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (node >= M) return;
    {   // blockIdx.y = pair
        const size_t b = blockIdx.y;
        nodes += b * M * 3; knn += b * M * K; kmask += b * M * K; pts += b * N * 3; nodes_out += b * M * 3; radius += b * M;
        if (rot) { rot += b * 9; trans += b * 3; }
    }
    float R[9], t[3];
    if (rot) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = __ldg(trans + k);
    }
    float c[3] = {__ldg(nodes + 3 * node), __ldg(nodes + 3 * node + 1), __ldg(nodes + 3 * node + 2)};
    if (rot) { float o[3]; apply_rt(R, t, c[0], c[1], c[2], o); c[0] = o[0]; c[1] = o[1]; c[2] = o[2]; }
    float mx = 0.f;
    for (int j = lane; j < K; j += 32) {
        if (!kmask[(size_t)node * K + j]) continue;
        const int pi = __ldg(knn + (size_t)node * K + j);
        float p[3] = {0.f, 0.f, 0.f};
        if (pi < N) { p[0] = __ldg(pts + 3 * (size_t)pi); p[1] = __ldg(pts + 3 * (size_t)pi + 1); p[2] = __ldg(pts + 3 * (size_t)pi + 2); }
        if (rot) { float o[3]; apply_rt(R, t, p[0], p[1], p[2], o); p[0] = o[0]; p[1] = o[1]; p[2] = o[2]; }
        const float dx = __fsub_rn(p[0], c[0]), dy = __fsub_rn(p[1], c[1]), dz = __fsub_rn(p[2], c[2]);
        mx = fmaxf(mx, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
    }
    mx = warp_max(mx);
    if (lane == 0) {
        radius[node] = mx;
        nodes_out[3 * node] = c[0]; nodes_out[3 * node + 1] = c[1]; nodes_out[3 * node + 2] = c[2];
    }
}

__device__ __forceinline__ float sqd_mm(const float* a, const float* b) {  // square_distance, lib/utils.py:139-156
    const float xy = fmaf(a[2], b[2], fmaf(a[1], b[1], __fmul_rn(a[0], b[0])));
    const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(a[0], a[0]), __fmul_rn(a[1], a[1])), __fmul_rn(a[2], a[2]));
    const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(b[0], b[0]), __fmul_rn(b[1], b[1])), __fmul_rn(b[2], b[2]));
    return fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, xy), a2), b2), 1e-12f);
}

// One CTA (4 warps) per ref node i (blockIdx.y = pair): (1) all threads run the enclosing-sphere prefilter (:580-590) over
// the src nodes and write the zero cells of row i directly (coalesced); the few pairs that pass go to a shared-memory list;
// (2) one WARP per listed pair runs the 64x64 point test (:592-606): a lane holds two (transformed) src patch points, the
// ref patch sits in shared memory. The first version launched one 64-thread CTA per (i, j) - 1.5 M CTAs per step of which
// ~90 % returned after the prefilter (0.83 ms per step, profiles/r02 kernel shares); counts are integers, so the result does
// not depend on the order of evaluation.
constexpr int NO_THREADS = 128;
__global__ void __launch_bounds__(NO_THREADS) node_overlap_kernel(int Mr, int Ms, int K, int Nr, int Nsrc,
                                                                  const float* __restrict__ rnodes, const float* __restrict__ snodes_t,
                                                                  const float* __restrict__ rrad, const float* __restrict__ srad,
                                                                  const unsigned char* __restrict__ rmask, const unsigned char* __restrict__ smask,
                                                                  const int* __restrict__ rknn, const int* __restrict__ sknn,
                                                                  const unsigned char* __restrict__ rkmask, const unsigned char* __restrict__ skmask,
                                                                  const float* __restrict__ rpts, const float* __restrict__ spts,
                                                                  const float* __restrict__ rot, const float* __restrict__ trans,
                                                                  float radius, float radius2, float* __restrict__ overlap,
                                                                  unsigned char* __restrict__ flag) {
    extern __shared__ int s_list[];                       // Ms candidate src nodes
    __shared__ float ra[64][3];
    __shared__ unsigned char rok[64];
    __shared__ int s_n, s_rn;
    const int i = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // blockIdx.y = pair
        const size_t b = blockIdx.y;
        rnodes += b * Mr * 3; snodes_t += b * Ms * 3; rrad += b * Mr; srad += b * Ms; rmask += b * Mr; smask += b * Ms;
        rknn += b * Mr * K; sknn += b * Ms * K; rkmask += b * Mr * K; skmask += b * Ms * K; rpts += b * Nr * 3; spts += b * Nsrc * 3;
        rot += b * 9; trans += b * 3; overlap += b * Mr * Ms; flag += b * Mr * Ms;
    }
    if (tid == 0) s_n = 0;
    if (tid < 64) {                                       // the ref patch (not transformed)
        const int ri = __ldg(rknn + (size_t)i * K + tid);
        float a[3] = {0.f, 0.f, 0.f};
        if (ri < Nr) { a[0] = __ldg(rpts + 3 * (size_t)ri); a[1] = __ldg(rpts + 3 * (size_t)ri + 1); a[2] = __ldg(rpts + 3 * (size_t)ri + 2); }
        ra[tid][0] = a[0]; ra[tid][1] = a[1]; ra[tid][2] = a[2];
        rok[tid] = rkmask[(size_t)i * K + tid];
    }
    __syncthreads();
    if (tid == 0) { int rn = 0; for (int c = 0; c < 64; ++c) rn += rok[c]; s_rn = rn; }
    // ---- (1) prefilter ----
    const bool iv = rmask[i];
    const float a0[3] = {__ldg(rnodes + 3 * i), __ldg(rnodes + 3 * i + 1), __ldg(rnodes + 3 * i + 2)};
    const float ri_rad = __ldg(rrad + i);
    for (int j = tid; j < Ms; j += NO_THREADS) {
        bool go = iv && smask[j];
        if (go) {
            float b3[3] = {__ldg(snodes_t + 3 * j), __ldg(snodes_t + 3 * j + 1), __ldg(snodes_t + 3 * j + 2)};
            const float d = __fsqrt_rn(sqd_mm(a0, b3));
            go = __fsub_rn(__fadd_rn(__fadd_rn(ri_rad, __ldg(srad + j)), radius), d) > 0.f;
        }
        if (go) s_list[atomicAdd(&s_n, 1)] = j;
        else { overlap[(size_t)i * Ms + j] = 0.f; flag[(size_t)i * Ms + j] = 0; }
    }
    __syncthreads();
    const int ncand = s_n, rn = s_rn;
    if (ncand == 0) return;
    float R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = __ldg(trans + k);
    // ---- (2) one warp per candidate pair ----
    for (int q = warp; q < ncand; q += NO_THREADS / 32) {
        const int j = s_list[q];
        float sp[2][3];
        bool sok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c = lane + 32 * u;
            const int pi = __ldg(sknn + (size_t)j * K + c);
            float p[3] = {0.f, 0.f, 0.f};
            if (pi < Nsrc) { p[0] = __ldg(spts + 3 * (size_t)pi); p[1] = __ldg(spts + 3 * (size_t)pi + 1); p[2] = __ldg(spts + 3 * (size_t)pi + 2); }
            apply_rt(R, t, p[0], p[1], p[2], sp[u]);
            sok[u] = skmask[(size_t)j * K + c];
        }
        bool colhit[2] = {false, false};
        int rc = 0;
        for (int r = 0; r < 64; ++r) {
            if (!rok[r]) continue;                         // warp-uniform
            const float a[3] = {ra[r][0], ra[r][1], ra[r][2]};
            const bool h0 = sok[0] && sqd_mm(a, sp[0]) < radius2;
            const bool h1 = sok[1] && sqd_mm(a, sp[1]) < radius2;
            colhit[0] |= h0; colhit[1] |= h1;
            rc += __any_sync(FULL_MASK, h0 || h1) ? 1 : 0;
        }
        const int sc = __popc(__ballot_sync(FULL_MASK, colhit[0])) + __popc(__ballot_sync(FULL_MASK, colhit[1]));
        const int sn = __popc(__ballot_sync(FULL_MASK, sok[0])) + __popc(__ballot_sync(FULL_MASK, sok[1]));
        if (lane == 0) {
            const float ov = __fdiv_rn(__fadd_rn(__fdiv_rn((float)rc, (float)rn), __fdiv_rn((float)sc, (float)sn)), 2.0f);
            overlap[(size_t)i * Ms + j] = ov;
            flag[(size_t)i * Ms + j] = ov > 0.f;
        }
    }
}

__global__ void corr_gather_kernel(const int* __restrict__ flat, const int* __restrict__ count, int capacity, int Ms, long long stride,
                                   const float* __restrict__ overlap, long long* __restrict__ out_idx,
                                   float* __restrict__ out_ov) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    {   // blockIdx.y = pair: flat / out (B, capacity), overlap (B, Mr*Ms = stride)
        const size_t b = blockIdx.y;
        flat += b * capacity; count += b; overlap += b * stride; out_idx += b * capacity * 2; out_ov += b * capacity;
    }
    if (i >= min(__ldg(count), capacity)) return;
    const int f = __ldg(flat + i);
    out_idx[2 * (size_t)i] = f / Ms;
    out_idx[2 * (size_t)i + 1] = f % Ms;
    out_ov[i] = __ldg(overlap + f);
}

}  // namespace

extern "C" int roitr_pad_transform_batched(int B, int N, const float* pts, const float* rot, const float* trans,
                                           float* out, void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && N >= 0 && pts && out && (!rot || trans), "pad_transform: bad arguments");
    pad_transform_kernel<<<dim3(ceil_div(N + 1, 256), B), 256, 0, (cudaStream_t)stream>>>(N, pts, rot, trans, out);
    ROITR_CHECK_LAUNCH("pad_transform_kernel");
    return ROITR_OK;
}

extern "C" int roitr_pad_transform(int N, const float* pts, const float* rot, const float* trans, float* out,
                                   void* stream) {
    return roitr_pad_transform_batched(1, N, pts, rot, trans, out, stream);
}

extern "C" int roitr_node_occlusion_batched(int B, int M, int K, int nn_stride, const int* knn, const unsigned char* kmask,
                                            const unsigned char* nmask, const float* nn_dist, float thr, float* occ,
                                            void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && M >= 1 && K >= 1 && knn && kmask && nmask && nn_dist && occ, "node_occlusion: bad arguments");
    node_occ_kernel<<<dim3(ceil_div(M * 32, 256), B), 256, 0, (cudaStream_t)stream>>>(M, K, knn, kmask, nmask, nn_dist, nn_stride, thr, occ);
    ROITR_CHECK_LAUNCH("node_occ_kernel");
    return ROITR_OK;
}

extern "C" int roitr_node_overlaps_batched(int B, int Mr, int Ms, int K, int Nr, int Nsrc, const float* ref_nodes,
                                           const float* src_nodes, const int* ref_knn, const int* src_knn,
                                           const unsigned char* ref_kmask, const unsigned char* src_kmask,
                                           const unsigned char* ref_mask, const unsigned char* src_mask, const float* ref_pts,
                                           const float* src_pts, const float* rot, const float* trans, float radius,
                                           float* work, float* overlap, unsigned char* flag, void* stream) {
    // work: B * (4*Ms + 4*Mr) floats
    ROITR_CHECK_ARG(K == 64, "node_overlaps: point_per_patch must be 64, got %d", K);
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && Mr >= 1 && Ms >= 1, "node_overlaps: bad sizes");
    ROITR_CHECK_ARG(ref_nodes && src_nodes && ref_knn && src_knn && rot && trans && work && overlap && flag, "node_overlaps: null");
    cudaStream_t st = (cudaStream_t)stream;
    float* snodes_t = work;
    float* srad = snodes_t + 3 * (size_t)B * Ms;
    float* rnodes_c = srad + (size_t)B * Ms;
    float* rrad = rnodes_c + 3 * (size_t)B * Mr;
    node_radius_kernel<<<dim3(ceil_div(Mr * 32, 256), B), 256, 0, st>>>(Mr, K, Nr, ref_nodes, ref_knn, ref_kmask, ref_pts, nullptr,
                                                                        nullptr, rnodes_c, rrad);
    node_radius_kernel<<<dim3(ceil_div(Ms * 32, 256), B), 256, 0, st>>>(Ms, K, Nsrc, src_nodes, src_knn, src_kmask, src_pts, rot,
                                                                        trans, snodes_t, srad);
    const double r2 = (double)radius * (double)radius;  // pos_radius ** 2 in double, then cast (lib/utils.py:597)
    ROITR_CHECK_ARG((size_t)Ms * 4 <= 40 * 1024, "node_overlaps: too many src nodes (%d)", Ms);
    dim3 grid(Mr, B);
    node_overlap_kernel<<<grid, NO_THREADS, (size_t)Ms * sizeof(int), st>>>(Mr, Ms, K, Nr, Nsrc, rnodes_c, snodes_t, rrad, srad, ref_mask, src_mask,
                                             ref_knn, src_knn, ref_kmask, src_kmask, ref_pts, src_pts, rot, trans,
                                             radius, (float)r2, overlap, flag);
    ROITR_CHECK_LAUNCH("node_overlap_kernel");
    return ROITR_OK;
}

extern "C" int roitr_corr_gather_batched(int B, int capacity, int Mr, int Ms, const int* flat, const int* count,
                                         const float* overlap, long long* out_idx, float* out_ov, void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && capacity >= 0, "corr_gather: bad sizes");
    if (capacity == 0) return ROITR_OK;
    corr_gather_kernel<<<dim3(ceil_div(capacity, 256), B), 256, 0, (cudaStream_t)stream>>>(flat, count, capacity, Ms, (long long)Mr * Ms,
                                                                                           overlap, out_idx, out_ov);
    ROITR_CHECK_LAUNCH("corr_gather_kernel");
    return ROITR_OK;
}
