// Small gather-style kernels of the forward path (sm_100a): inverse-distance interpolation, row gathers.
#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

// pointops.interpolation (cpp_wrappers/pointops/functions/pointops.py:168-182), after its kNN:
//   w_i = (1/(d_i + 1e-8)) / sum_i(1/(d_i + 1e-8));  out = sum_i feat[idx_i] * w_i   (+ base, model/model.py:116)
// The sum runs i = 0..k-1 with separately rounded multiply and add, like the eager loop it replaces.
__global__ void interpolate_kernel(int n, int c, int k, const int* __restrict__ idx, const float* __restrict__ dist,
                                   const float* __restrict__ feat, const float* __restrict__ base,
                                   float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c) return;
    const int row = (int)(e / c), ch = (int)(e % c);
    float r[8], s = 0.f;
    for (int i = 0; i < k; ++i) {
        r[i] = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + (size_t)row * k + i), 1e-8f));
        s = __fadd_rn(s, r[i]);
    }
    float acc = 0.f;
    for (int i = 0; i < k; ++i) {
        const float w = __fdiv_rn(r[i], s);
        const int j = __ldg(idx + (size_t)row * k + i);
        acc = __fadd_rn(acc, __fmul_rn(__ldg(feat + (size_t)j * c + ch), w));
    }
    if (base) acc = __fadd_rn(__ldg(base + e), acc);
    out[e] = acc;
}

// Same arithmetic, k = 3 and c a multiple of 4: a thread owns 4 consecutive channels of a row (float4 gathers / stores), so
// the six IEEE divisions of the weights are evaluated once per 4 outputs instead of once per output.
__global__ void interpolate3_vec4_kernel(int n, int c4, const int* __restrict__ idx, const float* __restrict__ dist,
                                         const float4* __restrict__ feat, const float4* __restrict__ base,
                                         float4* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c4) return;
    const int row = (int)(e / c4), q = (int)(e % c4);
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + (size_t)row * 3), 1e-8f));
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + (size_t)row * 3 + 1), 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__ldg(dist + (size_t)row * 3 + 2), 1e-8f));
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(0.f, r0), r1), r2);
    const float w0 = __fdiv_rn(r0, s), w1 = __fdiv_rn(r1, s), w2 = __fdiv_rn(r2, s);
    const float4 f0 = __ldg(feat + (size_t)__ldg(idx + (size_t)row * 3) * c4 + q);
    const float4 f1 = __ldg(feat + (size_t)__ldg(idx + (size_t)row * 3 + 1) * c4 + q);
    const float4 f2 = __ldg(feat + (size_t)__ldg(idx + (size_t)row * 3 + 2) * c4 + q);
    float4 a;
    a.x = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(f0.x, w0)), __fmul_rn(f1.x, w1)), __fmul_rn(f2.x, w2));
    a.y = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(f0.y, w0)), __fmul_rn(f1.y, w1)), __fmul_rn(f2.y, w2));
    a.z = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(f0.z, w0)), __fmul_rn(f1.z, w1)), __fmul_rn(f2.z, w2));
    a.w = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(f0.w, w0)), __fmul_rn(f1.w, w1)), __fmul_rn(f2.w, w2));
    if (base) {
        const float4 b = __ldg(base + e);
        a.x = __fadd_rn(b.x, a.x); a.y = __fadd_rn(b.y, a.y); a.z = __fadd_rn(b.z, a.z); a.w = __fadd_rn(b.w, a.w);
    }
    out[e] = a;
}

// out[i, :] = src[index[i], :]   (advanced indexing p[idx], model/model.py:67-68 and the index_select helper,
// lib/utils.py:403-425). index may be int32 or int64; rows of `c` floats. pad_row >= 0: index == pad_row -> zeros
// (the appended zero row of RIGA_v2.py:86-87,138-139 without materialising the padded copy).
template <typename IndexT>
__global__ void gather_rows_kernel(long long rows, int c, const IndexT* __restrict__ index,
                                   const float* __restrict__ src, float* __restrict__ out, long long pad_row) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * c) return;
    const long long r = e / c;
    const int ch = (int)(e % c);
    const long long j = (long long)index[r];
    out[e] = (j == pad_row) ? 0.f : __ldg(src + j * c + ch);
}

}  // namespace

extern "C" int roitr_interpolate(int n, int c, int k, const int* idx, const float* dist, const float* feat,
                                 const float* base, float* out, void* stream) {
    ROITR_CHECK_ARG(n >= 0 && c >= 1 && k >= 1 && k <= 8, "interpolate: bad n=%d c=%d k=%d", n, c, k);
    ROITR_CHECK_ARG(idx && dist && feat && out, "interpolate: null pointer");
    if (n == 0) return ROITR_OK;
    if (k == 3 && c % 4 == 0 && ((uintptr_t)feat | (uintptr_t)base | (uintptr_t)out) % 16 == 0) {
        const long long total4 = (long long)n * (c / 4);
        interpolate3_vec4_kernel<<<(unsigned)ceil_div_ll(total4, 256), 256, 0, (cudaStream_t)stream>>>(
            n, c / 4, idx, dist, (const float4*)feat, (const float4*)base, (float4*)out);
        ROITR_CHECK_LAUNCH("interpolate3_vec4_kernel");
        return ROITR_OK;
    }
    const long long total = (long long)n * c;
    interpolate_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(n, c, k, idx, dist, feat,
                                                                                            base, out);
    ROITR_CHECK_LAUNCH("interpolate_kernel");
    return ROITR_OK;
}

extern "C" int roitr_gather_rows(long long rows, int c, const void* index, int index_is_i64, const float* src,
                                 float* out, long long pad_row, void* stream) {
    ROITR_CHECK_ARG(rows >= 0 && c >= 1 && index && src && out, "gather_rows: bad arguments");
    if (rows == 0) return ROITR_OK;
    const unsigned grid = (unsigned)ceil_div_ll(rows * c, 256);
    if (index_is_i64)
        gather_rows_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, c, (const long long*)index, src, out,
                                                                             pad_row);
    else
        gather_rows_kernel<int><<<grid, 256, 0, (cudaStream_t)stream>>>(rows, c, (const int*)index, src, out, pad_row);
    ROITR_CHECK_LAUNCH("gather_rows_kernel");
    return ROITR_OK;
}
