// Dense layers of the forward path (sm_100a): y = x W^T + b with fused epilogues, and the row-wise epilogue
// (residual add, LayerNorm, ReLU, L2-normalise) that follows them.
//
// Replaces nn.Linear / nn.LayerNorm / F.relu / F.normalize call chains of model/model.py:90-117,131-142,
// model/transformer/attention.py:166-170,317-319, ppftransformer.py:244-251, geoattention.py:177-192,
// model/RIGA_v2.py:64-68. All arithmetic is fp32 (reference tolerance 1e-4 rules out TF32).
#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

constexpr int GEMM_THREADS = 256, BK = 8;

// C[M,N] = (A [+ A2])[M,K] * W[N,K]^T (+ bias) (+ ReLU). A rows optionally gathered through a_index (int32).
// 128 x BN CTA tile (BN = 128 or 64), 8-wide K slices double-buffered in shared memory (k-major), 8 x TN register
// micro-tile per thread (TN = BN/16): 4 (BN=128) or 3 (BN=64) LDS.128 per 64 / 32 FFMA, so the FMA pipe, not shared
// memory, is the limit. The micro-tile is split in 4-wide halves (rows ty*4.. and 64+ty*4.., cols tx*4.. and BN/2+tx*4..)
// so that the float4 shared-memory reads of a quarter-warp hit distinct banks.
template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
linear_kernel(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ A2, int lda,
              const int* __restrict__ a_index, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
              float* __restrict__ C, int ldc, int relu) {
    constexpr int BM = 128, TN = BN / 16, NH = TN / 4;   // NH column halves of 4
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Ws[2][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // loaders: A tile 128 rows x 8 k = 256 float4 -> one per thread (row = tid/2, k half = tid%2);
    //          W tile BN rows x 8 k -> BN*2 float4: threads < 2*BN
    const int lrow = tid >> 1, lk = (tid & 1) * 4;
    const int am = m0 + lrow;
    long long arow = -1;
    if (am < M) arow = a_index ? (long long)__ldg(a_index + am) : (long long)am;
    const int wn = n0 + lrow;
    const bool w_thread = lrow < BN;
    const bool vecA = (lda % 4 == 0) && ((uintptr_t)A % 16 == 0) && (!A2 || (uintptr_t)A2 % 16 == 0);
    const bool vecW = (ldw % 4 == 0) && ((uintptr_t)W % 16 == 0);

    auto ld4 = [&](const float* base, const float* base2, long long row, int ld, int k, bool vec) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < 0) return x;
        const float* p = base + row * ld + k;
        if (vec && k + 3 < K) {
            x = __ldg(reinterpret_cast<const float4*>(p));
            if (base2) { const float4 y = __ldg(reinterpret_cast<const float4*>(base2 + row * ld + k)); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k + e < K) t[e] = __ldg(p + e) + (base2 ? __ldg(base2 + row * ld + k + e) : 0.f);   // x + pos (geoattention.py:43-44)
            x = make_float4(t[0], t[1], t[2], t[3]);
        }
        return x;
    };
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nkt = (K + BK - 1) / BK;
    float4 ra = ld4(A, A2, arow, lda, lk, vecA);
    float4 rw = w_thread ? ld4(W, nullptr, wn < N ? (long long)wn : -1, ldw, lk, vecW) : make_float4(0.f, 0.f, 0.f, 0.f);
    As[0][lk][lrow] = ra.x; As[0][lk + 1][lrow] = ra.y; As[0][lk + 2][lrow] = ra.z; As[0][lk + 3][lrow] = ra.w;
    if (w_thread) { Ws[0][lk][lrow] = rw.x; Ws[0][lk + 1][lrow] = rw.y; Ws[0][lk + 2][lrow] = rw.z; Ws[0][lk + 3][lrow] = rw.w; }
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) {   // global loads in flight during the math
            ra = ld4(A, A2, arow, lda, (kt + 1) * BK + lk, vecA);
            if (w_thread) rw = ld4(W, nullptr, wn < N ? (long long)wn : -1, ldw, (kt + 1) * BK + lk, vecW);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][h * (BN / 2) + tx * 4]);
                bv[4 * h] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nkt) {
            const int nb = buf ^ 1;
            As[nb][lk][lrow] = ra.x; As[nb][lk + 1][lrow] = ra.y; As[nb][lk + 2][lrow] = ra.z; As[nb][lk + 3][lrow] = ra.w;
            if (w_thread) { Ws[nb][lk][lrow] = rw.x; Ws[nb][lk + 1][lrow] = rw.y; Ws[nb][lk + 2][lrow] = rw.z; Ws[nb][lk + 3][lrow] = rw.w; }
        }
        __syncthreads();
    }
    const bool vecC = (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int n = n0 + h * (BN / 2) + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][4 * h + j] + ((bias && n + j < N) ? __ldg(bias + n + j) : 0.f);
                if (relu) v[j] = fmaxf(v[j], 0.f);
            }
            float* dst = C + (long long)m * ldc + n;
            if (vecC && n + 3 < N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) dst[j] = v[j];
            }
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
    if (N <= 64) {
        dim3 grid(ceil_div(M, 128), 1);
        linear_kernel<64><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(M, N, K, A, a_add, lda, a_index, W, ldw, bias, C,
                                                                          ldc, relu);
    } else {
        dim3 grid(ceil_div(M, 128), ceil_div(N, 128));
        linear_kernel<128><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(M, N, K, A, a_add, lda, a_index, W, ldw, bias, C,
                                                                           ldc, relu);
    }
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
