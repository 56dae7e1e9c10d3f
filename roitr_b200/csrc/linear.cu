// Dense layers of the forward path (sm_100a): y = x W^T + b with fused epilogues, and the row-wise epilogue
// (residual add, LayerNorm, ReLU, L2-normalise) that follows them.
//
// Replaces nn.Linear / nn.LayerNorm / F.relu / F.normalize call chains of model/model.py:90-117,131-142,
// model/transformer/attention.py:166-170,317-319, ppftransformer.py:244-251, geoattention.py:177-192,
// model/RIGA_v2.py:64-68. All arithmetic is fp32 (reference tolerance 1e-4 rules out TF32).
#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256;

// C[M,N] = (A [+ A2])[M,K] * W[N,K]^T (+ bias) (+ ReLU). A rows optionally gathered through a_index (int32).
// 64x64 CTA tile, 16-wide K slices staged in shared memory (k-major, so the inner product reads two float4 per k),
// 4x4 register micro-tile per thread.
__global__ void __launch_bounds__(GEMM_THREADS)
linear_kernel(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ A2, int lda,
              const int* __restrict__ a_index, const float* __restrict__ W, int ldw, const float* __restrict__ bias, float* __restrict__ C, int ldc,
              int relu) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 (m) x 4 (n)
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // loader mapping: 64 rows x 16 k = 1024 elements per operand, 4 per thread: row = tid/4, k = (tid%4)*4..+3
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int am = m0 + lrow;
    long long arow = -1;
    if (am < M) arow = a_index ? (long long)__ldg(a_index + am) : (long long)am;
    const int wn = n0 + lrow;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto load_tile = [&](int kt, float (&ra)[4], float (&rw)[4]) {
        const int k = kt * BK + lk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ra[i] = (arow >= 0 && k + i < K) ? __ldg(A + arow * lda + k + i) : 0.f;
            if (A2 && arow >= 0 && k + i < K) ra[i] += __ldg(A2 + arow * lda + k + i);  // x + pos (geoattention.py:43-44)
            rw[i] = (wn < N && k + i < K) ? __ldg(W + (long long)wn * ldw + k + i) : 0.f;
        }
    };
    auto store_tile = [&](int buf, const float (&ra)[4], const float (&rw)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[buf][lk + i][lrow] = ra[i];
            Ws[buf][lk + i][lrow] = rw[i];
        }
    };

    const int nkt = (K + BK - 1) / BK;
    float ra[4], rw[4];
    load_tile(0, ra, rw);
    store_tile(0, ra, rw);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_tile(kt + 1, ra, rw);  // global loads in flight during the math
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nkt) store_tile(buf ^ 1, ra, rw);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[(long long)m * ldc + n] = v;
        }
    }
}

// One warp per row of width C <= 1024:
//   v = x[row] (+ res_pre[idx?row])            residual before the norm   (attention.py:319: norm(h + x[node_idx]))
//   v = LayerNorm(v) * gamma + beta  (eps 1e-5)                            optional
//   v = v + res_post[row]                      residual after the norm    (model/model.py:138-139: bn2(x) + identity)
//   v = relu(v)                                                            optional
//   v = v / max(|v|_2, 1e-12)                  F.normalize (RIGA_v2.py:64) optional
// mode bits: 1 = LN, 2 = ReLU, 4 = L2-normalise.
template <int VPL>  // values per lane: C <= 32*VPL
__global__ void row_epilogue_kernel(int M, int C, const float* __restrict__ x, const float* __restrict__ res_pre,
                                    const int* __restrict__ res_pre_index, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ res_post,
                                    float* __restrict__ out, int mode) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float v[VPL];
    const long long pre_row = res_pre ? (res_pre_index ? (long long)__ldg(res_pre_index + row) : (long long)row) : 0;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + 32 * i;
        v[i] = 0.f;
        if (c < C) {
            v[i] = __ldg(x + (long long)row * C + c);
            if (res_pre) v[i] += __ldg(res_pre + pre_row * C + c);
        }
    }
    if (mode & 1) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s += (lane + 32 * i < C) ? v[i] : 0.f;
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float d = (lane + 32 * i < C) ? v[i] - mean : 0.f;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) v[i] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
    }
    if (res_post) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) v[i] += __ldg(res_post + (long long)row * C + c);
        }
    }
    if (mode & 2) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (mode & 4) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s += (lane + 32 * i < C) ? v[i] * v[i] : 0.f;
        const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = v[i] / nrm;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + 32 * i;
        if (c < C) out[(long long)row * C + c] = v[i];
    }
}

// column mean over a segment of rows (TransitionUp head: x.sum(0)/cnt, model/model.py:108) -> out[seg, c]
__global__ void segment_mean_kernel(int b, int C, const float* __restrict__ x, const int* __restrict__ offset,
                                    float* __restrict__ out) {
    const int seg = blockIdx.x;
    const int s = seg == 0 ? 0 : __ldg(offset + seg - 1), e = __ldg(offset + seg);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int r = s; r < e; ++r) acc += __ldg(x + (long long)r * C + c);
        out[(long long)seg * C + c] = acc / (float)(e - s);
    }
}

// out[row, 0:C] = x[row, :], out[row, C:2C] = g[segment(row), :]   (torch.cat((x_b, g.repeat(cnt,1)), 1), model.py:108)
__global__ void concat_segment_kernel(int M, int C, int b, const float* __restrict__ x, const float* __restrict__ g,
                                      const int* __restrict__ offset, float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)M * 2 * C) return;
    const int row = (int)(e / (2 * C)), c = (int)(e % (2 * C));
    if (c < C) { out[e] = __ldg(x + (long long)row * C + c); return; }
    int seg = 0;
    while (seg < b - 1 && row >= __ldg(offset + seg)) ++seg;
    out[e] = __ldg(g + (long long)seg * C + (c - C));
}

}  // namespace

extern "C" int roitr_linear(int M, int N, int K, const float* A, const float* a_add, int lda, const int* a_index,
                            const float* W, int ldw, const float* bias, float* C, int ldc, int relu, void* stream) {
    ROITR_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && A && W && C, "linear: bad arguments M=%d N=%d K=%d", M, N, K);
    ROITR_CHECK_ARG(lda >= K && ldc >= N && ldw >= K, "linear: bad leading dimensions");
    if (M == 0) return ROITR_OK;
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    linear_kernel<<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(M, N, K, A, a_add, lda, a_index, W, ldw, bias, C, ldc,
                                                                   relu);
    ROITR_CHECK_LAUNCH("linear_kernel");
    return ROITR_OK;
}

extern "C" int roitr_row_epilogue(int M, int C, const float* x, const float* res_pre, const int* res_pre_index,
                                  const float* gamma, const float* beta, const float* res_post, float* out, int mode,
                                  void* stream) {
    ROITR_CHECK_ARG(M >= 0 && C >= 1 && C <= 1024 && x && out, "row_epilogue: bad arguments M=%d C=%d", M, C);
    ROITR_CHECK_ARG(!(mode & 1) || (gamma && beta), "row_epilogue: LayerNorm needs gamma/beta");
    if (M == 0) return ROITR_OK;
    const int warps = 8;
    dim3 grid(ceil_div(M, warps));
    cudaStream_t st = (cudaStream_t)stream;
#define RE(V) row_epilogue_kernel<V><<<grid, warps * 32, 0, st>>>(M, C, x, res_pre, res_pre_index, gamma, beta, res_post, out, mode)
    if (C <= 64) RE(2);
    else if (C <= 128) RE(4);
    else if (C <= 256) RE(8);
    else if (C <= 512) RE(16);
    else RE(32);
#undef RE
    ROITR_CHECK_LAUNCH("row_epilogue_kernel");
    return ROITR_OK;
}

extern "C" int roitr_segment_mean(int b, int C, const float* x, const int* offset, float* out, void* stream) {
    ROITR_CHECK_ARG(b >= 1 && C >= 1 && x && offset && out, "segment_mean: bad arguments");
    segment_mean_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(b, C, x, offset, out);
    ROITR_CHECK_LAUNCH("segment_mean_kernel");
    return ROITR_OK;
}

extern "C" int roitr_concat_segment(int M, int C, int b, const float* x, const float* g, const int* offset, float* out,
                                    void* stream) {
    ROITR_CHECK_ARG(M >= 0 && C >= 1 && b >= 1 && x && g && offset && out, "concat_segment: bad arguments");
    if (M == 0) return ROITR_OK;
    const long long total = (long long)M * 2 * C;
    concat_segment_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(M, C, b, x, g, offset, out);
    ROITR_CHECK_LAUNCH("concat_segment_kernel");
    return ROITR_OK;
}
