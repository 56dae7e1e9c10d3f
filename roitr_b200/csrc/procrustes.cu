// Weighted Procrustes (rigid transform from weighted correspondences), the step after the hot path.
//
// Replaces weighted_procrustes (lib/utils.py:159-218): thresholded weights, weighted centroids with
// w / (sum w + eps), H = sum_i w_i (p_i - cp)(q_i - cq)^T, R = V diag(1, 1, sign det(V U^T)) U^T from H = U S V^T,
// t = cq - R cp. The reference calls torch.svd on a 3x3 matrix per batch item and ~15 elementwise kernels.
//
// One CTA per batch item: two block reductions (moments, then the 3x3 cross-covariance, both accumulated in fp64), then
// thread 0 finds the optimal proper rotation as the dominant eigenvector of Horn's symmetric 4x4 matrix built from H
// (cyclic Jacobi in fp64). For every H with a unique optimum this IS the reference's result: V diag(1,1,sign det) U^T is
// the proper rotation maximising trace(R H), which is what the quaternion form maximises; it never needs the
// sign-ambiguous singular vectors themselves.
#include "../../include/roitr_b200.h"
#include "common.cuh"
#include "horn.cuh"

namespace {

constexpr int PR_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < PR_THREADS / 32; ++i) t += sh[i];
    return t;
}

__global__ void __launch_bounds__(PR_THREADS) procrustes_kernel(int n, const float* __restrict__ src, const float* __restrict__ tgt,
                                                               const float* __restrict__ weights, float thresh, float eps,
                                                               float* __restrict__ R_out, float* __restrict__ t_out) {
    __shared__ double sh[PR_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* P = src + (size_t)b * n * 3;
    const float* Q = tgt + (size_t)b * n * 3;
    const float* Wt = weights ? weights + (size_t)b * n : nullptr;
    auto wgt = [&](int i) {
        const float w = Wt ? __ldg(Wt + i) : 1.0f;
        return (double)(w < thresh ? 0.0f : w);
    };
    double m[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += PR_THREADS) {
        const double w = wgt(i);
        m[0] += w;
        for (int d = 0; d < 3; ++d) { m[1 + d] += w * (double)__ldg(P + 3 * i + d); m[4 + d] += w * (double)__ldg(Q + 3 * i + d); }
    }
    for (int k = 0; k < 7; ++k) m[k] = block_sum(m[k], sh);
    const double inv = 1.0 / (m[0] + (double)eps);           // weights / (sum + eps): lib/utils.py:189
    const double cp[3] = {m[1] * inv, m[2] * inv, m[3] * inv}, cq[3] = {m[4] * inv, m[5] * inv, m[6] * inv};
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += PR_THREADS) {
        const double w = wgt(i);
        double p[3], q[3];
        for (int d = 0; d < 3; ++d) { p[d] = (double)__ldg(P + 3 * i + d) - cp[d]; q[d] = (double)__ldg(Q + 3 * i + d) - cq[d]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] += w * p[r] * q[c];
    }
    for (int k = 0; k < 9; ++k) H[k] = block_sum(H[k], sh);
    if (tid != 0) return;
    double R[9];
    horn_rotation(H, R);
    for (int k = 0; k < 9; ++k) R_out[9 * (size_t)b + k] = (float)R[k];
    for (int r = 0; r < 3; ++r)
        t_out[3 * (size_t)b + r] = (float)(cq[r] - (R[3 * r] * cp[0] + R[3 * r + 1] * cp[1] + R[3 * r + 2] * cp[2]));
}

}  // namespace

extern "C" int roitr_weighted_procrustes(int batch, int n, const float* src, const float* tgt, const float* weights,
                                         float weight_thresh, float eps, float* R, float* t, void* stream) {
    ROITR_CHECK_ARG(batch >= 1 && n >= 1 && src && tgt && R && t, "weighted_procrustes: bad arguments");
    procrustes_kernel<<<batch, PR_THREADS, 0, (cudaStream_t)stream>>>(n, src, tgt, weights, weight_thresh, eps, R, t);
    ROITR_CHECK_LAUNCH("procrustes_kernel");
    return ROITR_OK;
}
