// Superpoint 3-NN for the geometric structure embedding (sm_100a).
//
// Replaces the k-NN part of GeometricStructureEmbedding.get_embedding_indices (model/transformer/positional_encoding.py:
// 111-137, with pairwise_distance :9-34). The embedding itself is csrc/geo_table.cu (tabulated) with csrc/geo_tc.cu (tcgen05
// GEMM) as its fallback; the attention is csrc/geo_attn2.cu + csrc/gemm_tc.cu. The first-generation fp32 FFMA embedding and
// attention kernels that lived here were test comparators only and are gone (tests compare against float64 PyTorch).
#include <math_constants.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ 3-NN of the nodes
// dist = sqrt(clamp(x2 - 2 xy + y2, 0)) exactly as pairwise_distance (positional_encoding.py:24-33): xy is the matmul
// (fma chain over the 3 coordinates), x2/y2 are separately rounded sums of squares.
__device__ __forceinline__ float sq3(const float* p) {
    return __fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2]));
}
__device__ __forceinline__ float pair_dist(const float* a, float a2, const float* b, float b2) {
    const float xy = fmaf(a[2], b[2], fmaf(a[1], b[1], __fmul_rn(a[0], b[0])));
    const float d2 = __fadd_rn(__fsub_rn(a2, __fmul_rn(2.0f, xy)), b2);
    return __fsqrt_rn(fmaxf(d2, 0.0f));
}

// one warp per node: the k+1 smallest distances of its row (topk(k+1, largest=False), :124), first one dropped.
// Ties: lower index first.
template <int KP1>
__global__ void geo_knn_kernel(int N, const float* __restrict__ pts, int* __restrict__ nn) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    pts += (size_t)blockIdx.y * N * 3;               // blockIdx.y = cloud of a batch (indices stay cloud-local)
    nn += (size_t)blockIdx.y * N * (KP1 - 1);
    float a[3] = {__ldg(pts + 3 * n), __ldg(pts + 3 * n + 1), __ldg(pts + 3 * n + 2)};
    const float a2 = sq3(a);
    float bd[KP1];
    int bi[KP1];
#pragma unroll
    for (int i = 0; i < KP1; ++i) { bd[i] = CUDART_INF_F; bi[i] = 0x7fffffff; }
    for (int m = lane; m < N; m += 32) {
        float b[3] = {__ldg(pts + 3 * m), __ldg(pts + 3 * m + 1), __ldg(pts + 3 * m + 2)};
        float d = pair_dist(a, a2, b, sq3(b));
        int id = m;
#pragma unroll
        for (int i = 0; i < KP1; ++i) {  // sorted insert (ascending by (d, idx))
            const bool lt = (d < bd[i]) || (d == bd[i] && id < bi[i]);
            const float td = lt ? bd[i] : d; const int ti = lt ? bi[i] : id;
            bd[i] = lt ? d : bd[i]; bi[i] = lt ? id : bi[i];
            d = td; id = ti;
        }
    }
    // merge the 32 sorted lists: KP1 rounds of warp-wide arg-min over the list heads
#pragma unroll
    for (int r = 0; r < KP1; ++r) {
        float hd = bd[0]; int hi = bi[0];
        float md = hd; int mi = hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(FULL_MASK, md, o);
            const int oi = __shfl_xor_sync(FULL_MASK, mi, o);
            if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
        }
        if (hd == md && hi == mi) {  // pop my head
#pragma unroll
            for (int i = 0; i + 1 < KP1; ++i) { bd[i] = bd[i + 1]; bi[i] = bi[i + 1]; }
            bd[KP1 - 1] = CUDART_INF_F; bi[KP1 - 1] = 0x7fffffff;
        }
        if (r >= 1 && lane == 0) nn[(size_t)n * (KP1 - 1) + (r - 1)] = mi;
    }
}

// ------------------------------------------------------------------------------------- geometric structure embedding
// E[n,m,:] = proj_d(sinusoid(D_nm / sigma_d)) + max_r proj_a(sinusoid(angle_nmr * factor_a))     (:139-154)
// A GEMM whose A operand (4 sinusoid rows of width C per (n,m) pair) is generated on the fly, K-slice by K-slice, and
// never touches memory; the (N,N,3,C) intermediate of the reference (299 MB at N=312) does not exist.
}  // namespace

extern "C" int roitr_geo_knn_batched(int batch, int N, int k, const float* pts, int* nn, void* stream) {
    ROITR_CHECK_ARG(batch >= 1 && batch <= 65535 && N >= 1 && pts && nn, "geo_knn: bad arguments");
    ROITR_CHECK_ARG(k == 3, "geo_knn: angle_k = 3 only (model/model.py:165), got %d", k);
    geo_knn_kernel<4><<<dim3(ceil_div(N * 32, 256), batch), 256, 0, (cudaStream_t)stream>>>(N, pts, nn);
    ROITR_CHECK_LAUNCH("geo_knn_kernel");
    return ROITR_OK;
}

extern "C" int roitr_geo_knn(int N, int k, const float* pts, int* nn, void* stream) {
    return roitr_geo_knn_batched(1, N, k, pts, nn, stream);
}
