// Global geometric transformer kernels (sm_100a): superpoint 3-NN, geometric structure embedding with the sinusoids
// generated in-kernel, and the RPE self- / plain cross-attention core.
//
// Replaces GeometricStructureEmbedding (model/transformer/positional_encoding.py:94-154, with pairwise_distance :9-34
// and SinusoidalPositionalEmbedding :38-62), RPEMultiHeadAttention.forward (geoattention.py:101-136) and
// MultiHeadAttention.forward (geoattention.py:43-66).
//
// Exact folds used by the attention kernel (DESIGN.md "global fold"): with p_nm = W_p E_nm + b_p and
// vp_nm = W_vp E_nm + b_vp,
//     q_n,h . p_nm,h       = (W_p,h^T q_n,h) . E_nm + q_n,h . b_p,h     -> gq (N, H, C) from one small GEMM per head
//     sum_m A-_nm vp_nm,h  = W_vp,h (sum_m A-_nm E_nm) + b_vp,h          (softmax rows sum to 1)
// so the two N^2 x C x C GEMMs per self layer (25.5 GFLOP per layer per cloud at N=312) disappear and E is read twice.
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
constexpr int GE_PAIRS = 32, GE_ROWS = GE_PAIRS * 4, GE_BN = 128, GE_BK = 16, GE_THREADS = 256;

struct GeoEmbParams {
    const float* pts; const int* nn3;
    const float* Wd; const float* bd; const float* Wa; const float* ba; const float* div_term;
    float* E;
    int N, C;
    float sigma_d, factor_a;
};

__global__ void __launch_bounds__(GE_THREADS) geo_embedding_kernel(const GeoEmbParams P) {
    __shared__ __align__(16) float As[2][GE_BK][GE_ROWS];
    __shared__ __align__(16) float Wds[2][GE_BK][GE_BN];
    __shared__ __align__(16) float Was[2][GE_BK][GE_BN];
    const int tid = threadIdx.x;
    const int N = P.N, C = P.C;
    const long long p0 = (long long)blockIdx.x * GE_PAIRS;
    const long long npairs = (long long)N * N;
    const int n0 = blockIdx.y * GE_BN;

    // embedding index t of the row this thread generates (row r = tid % 128: pair r/4, kind r%4)
    const int gr = tid & (GE_ROWS - 1);
    float t_row = 0.f;
    {
        const long long pidx = p0 + (gr >> 2);
        const int kind = gr & 3;
        if (pidx < npairs) {
            const int n = (int)(pidx / N), m = (int)(pidx % N);
            float a[3] = {__ldg(P.pts + 3 * n), __ldg(P.pts + 3 * n + 1), __ldg(P.pts + 3 * n + 2)};
            float b[3] = {__ldg(P.pts + 3 * m), __ldg(P.pts + 3 * m + 1), __ldg(P.pts + 3 * m + 2)};
            if (kind == 0) {
                t_row = __fdiv_rn(pair_dist(a, sq3(a), b, sq3(b)), P.sigma_d);
            } else {
                const int r = __ldg(P.nn3 + 3 * n + (kind - 1));
                const float rx = __fsub_rn(__ldg(P.pts + 3 * r), a[0]), ry = __fsub_rn(__ldg(P.pts + 3 * r + 1), a[1]),
                            rz = __fsub_rn(__ldg(P.pts + 3 * r + 2), a[2]);
                const float ax = __fsub_rn(b[0], a[0]), ay = __fsub_rn(b[1], a[1]), az = __fsub_rn(b[2], a[2]);
                const float cx = __fsub_rn(__fmul_rn(ry, az), __fmul_rn(rz, ay));
                const float cy = __fsub_rn(__fmul_rn(rz, ax), __fmul_rn(rx, az));
                const float cz = __fsub_rn(__fmul_rn(rx, ay), __fmul_rn(ry, ax));
                const float sinv = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
                // torch.sum starts from +0: (+0) + (-0) = +0, so a zero dot product is +0 and atan2(0, +0) = 0 (not pi)
                const float cosv = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(rx, ax)), __fmul_rn(ry, ay)), __fmul_rn(rz, az));
                t_row = __fmul_rn(atan2f(sinv, cosv), P.factor_a);
            }
        }
    }
    const int gj = tid >> 7;  // 0/1: which of the (8 frequencies per slice) this thread starts at

    auto gen_a = [&](int kt, int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int jj = gj + 2 * i;                      // frequency within the slice (0..7)
            const int j = kt * (GE_BK / 2) + jj;            // global frequency index
            float s, c;
            sincosf(__fmul_rn(t_row, __ldg(P.div_term + j)), &s, &c);   // omegas = idx * div_term (:57)
            As[buf][2 * jj][gr] = s;                        // interleaved [sin, cos] (:60-61)
            As[buf][2 * jj + 1][gr] = c;
        }
    };
    // weight loader: 128 output columns x 16 k for each of Wd, Wa: 2048 floats each, 8 per thread per matrix
    const int wrow = tid >> 1, wk = (tid & 1) * 8;
    auto load_w = [&](int kt, float (&rd)[8], float (&ra)[8]) {
        const float* pd = P.Wd + (size_t)(n0 + wrow) * C + kt * GE_BK + wk;
        const float* pa = P.Wa + (size_t)(n0 + wrow) * C + kt * GE_BK + wk;
        const float4 d0 = __ldg(reinterpret_cast<const float4*>(pd)), d1 = __ldg(reinterpret_cast<const float4*>(pd) + 1);
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(pa)), a1 = __ldg(reinterpret_cast<const float4*>(pa) + 1);
        rd[0] = d0.x; rd[1] = d0.y; rd[2] = d0.z; rd[3] = d0.w; rd[4] = d1.x; rd[5] = d1.y; rd[6] = d1.z; rd[7] = d1.w;
        ra[0] = a0.x; ra[1] = a0.y; ra[2] = a0.z; ra[3] = a0.w; ra[4] = a1.x; ra[5] = a1.y; ra[6] = a1.z; ra[7] = a1.w;
    };
    auto store_w = [&](int buf, const float (&rd)[8], const float (&ra)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { Wds[buf][wk + i][wrow] = rd[i]; Was[buf][wk + i][wrow] = ra[i]; }
    };

    // 8 rows (= 2 pairs x 4 kinds) x 8 cols per thread; the cols are tx*4..+3 and 64+tx*4..+3 so that the float4
    // shared-memory reads of a quarter-warp hit distinct banks
    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nkt = C / GE_BK;
    float rd[8], ra[8];
    load_w(0, rd, ra);
    gen_a(0, 0);
    store_w(0, rd, ra);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_w(kt + 1, rd, ra);
#pragma unroll
        for (int k = 0; k < GE_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 d0 = *reinterpret_cast<const float4*>(&Wds[buf][k][tx * 4]);
            const float4 d1 = *reinterpret_cast<const float4*>(&Wds[buf][k][64 + tx * 4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Was[buf][k][tx * 4]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Was[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], ((i & 3) == 0) ? dv[j] : wv[j], acc[i][j]);
        }
        if (kt + 1 < nkt) { gen_a(kt + 1, buf ^ 1); store_w(buf ^ 1, rd, ra); }
        __syncthreads();
    }
    // epilogue: rows [0..3] = pair 2*ty, rows [4..7] = pair 2*ty+1; kind 0 = distance, 1..3 = angles
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long pidx = p0 + ty * 2 + h;
        if (pidx >= npairs) continue;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            const float b_a = __ldg(P.ba + col);
            const float d = acc[4 * h][j] + __ldg(P.bd + col);
            const float a = fmaxf(fmaxf(acc[4 * h + 1][j] + b_a, acc[4 * h + 2][j] + b_a), acc[4 * h + 3][j] + b_a);
            o[j] = d + a;
        }
        float* dst = P.E + (size_t)pidx * C + n0 + tx * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 64) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// ------------------------------------------------------------------------------------------------ attention core
// One CTA per query row n. RPE = self layer with embedding E (N,N,C): also emits G[n,h,:] = sum_m A-_nm E_nm for the
// position branch, A- = softmax of the scores with the diagonal removed (geoattention.py:117-123,132-133).
struct GeoAttnParams {
    const float* q; int ldq;     // (N, C)
    const float* k; int ldk;     // (M, C)
    const float* v; int ldv;     // (M, C)
    const float* E;              // (N, M=N, C) or NULL
    const float* gq;             // (N, 4, C) folded positional queries or NULL
    const float* bp;             // (C) proj_p bias or NULL
    float* hidden;               // (N, C)
    float* G;                    // (N, 4, C) or NULL
    int N, M;
    float sqrt_c;
    long long q_bs, k_bs, v_bs;  // element strides between the clouds of a batch (blockIdx.y) for q / k / v
};

template <int C, bool RPE>
__global__ void __launch_bounds__(256) geo_attention_kernel(const GeoAttnParams P_) {
    constexpr int CPL = C / 32, H = 4;
    extern __shared__ float sm[];
    GeoAttnParams P = P_;
    {   // batch element: every operand advances by one cloud
        const long long b = blockIdx.y;
        P.q += b * P.q_bs; P.k += b * P.k_bs; P.v += b * P.v_bs;
        P.hidden += b * (long long)P.N * C;
        if (RPE) { P.E += b * (long long)P.N * P.M * C; P.gq += b * (long long)P.N * H * C; P.G += b * (long long)P.N * H * C; }
    }
    const int M = P.M;
    float* S = sm;               // [H][M] scores -> attention
    float* Sm = sm + H * M;      // [H][M] attention without self (RPE only)
    const int n = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = lane * CPL;
    const int head = lane >> 3;

    float q[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) q[i] = __ldg(P.q + (size_t)n * P.ldq + c0 + i);
    float gq[RPE ? H : 1][CPL];
    float qb = 0.f;
    if (RPE) {
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int i = 0; i < CPL; ++i) gq[h][i] = __ldg(P.gq + ((size_t)n * H + h) * C + c0 + i);
#pragma unroll
        for (int i = 0; i < CPL; ++i) qb = fmaf(q[i], __ldg(P.bp + c0 + i), qb);
        qb += __shfl_xor_sync(FULL_MASK, qb, 1); qb += __shfl_xor_sync(FULL_MASK, qb, 2); qb += __shfl_xor_sync(FULL_MASK, qb, 4);
    }
    // ---- phase 1: scores ----
    for (int m = warp; m < M; m += 8) {
        float kr[CPL];
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(P.k + (size_t)m * P.ldk + c0) + i);
            kr[4 * i] = t.x; kr[4 * i + 1] = t.y; kr[4 * i + 2] = t.z; kr[4 * i + 3] = t.w;
        }
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) se = fmaf(q[i], kr[i], se);
        se += __shfl_xor_sync(FULL_MASK, se, 1); se += __shfl_xor_sync(FULL_MASK, se, 2); se += __shfl_xor_sync(FULL_MASK, se, 4);
        float sp = 0.f;
        if (RPE) {
            float er[CPL];
            const float* ep = P.E + ((size_t)n * M + m) * C + c0;
#pragma unroll
            for (int i = 0; i < CPL / 4; ++i) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(ep) + i);
                er[4 * i] = t.x; er[4 * i + 1] = t.y; er[4 * i + 2] = t.z; er[4 * i + 3] = t.w;
            }
            float ph[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < CPL; ++i) a = fmaf(gq[h][i], er[i], a);
                ph[h] = warp_sum(a);
            }
            sp = (head == 0 ? ph[0] : head == 1 ? ph[1] : head == 2 ? ph[2] : ph[3]) + qb;
        }
        if ((lane & 7) == 0) S[head * M + m] = __fdiv_rn(se + sp, P.sqrt_c);
    }
    __syncthreads();
    // ---- phase 2: softmax per head (warps 0-3) and self-excluded softmax (warps 4-7, RPE only) ----
    {
        const int h = warp & 3;
        const bool noself = warp >= 4;
        const bool active = !noself || RPE;
        float mx = -CUDART_INF_F, den = 0.f;
        if (active) {
            for (int m = lane; m < M; m += 32)
                if (!(noself && m == n)) mx = fmaxf(mx, S[h * M + m]);
            mx = warp_max(mx);
            for (int m = lane; m < M; m += 32)
                if (!(noself && m == n)) den += expf(S[h * M + m] - mx);
            den = warp_sum(den);
            if (noself)
                for (int m = lane; m < M; m += 32) Sm[h * M + m] = (m == n) ? 0.f : expf(S[h * M + m] - mx) / den;
        }
        __syncthreads();  // warps 4-7 have read S before warps 0-3 overwrite it in place
        if (active && !noself)
            for (int m = lane; m < M; m += 32) S[h * M + m] = expf(S[h * M + m] - mx) / den;
    }
    __syncthreads();
    // ---- phase 3: value aggregate (+ positional aggregate) ----
    constexpr int CPT = C / 256;
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
        const int ch = tid + 256 * u;
        const int hh = ch / (C / H);
        float acc = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        const float* vp = P.v + ch;
        const float* ep = RPE ? P.E + (size_t)n * M * C + ch : nullptr;
#pragma unroll 4
        for (int m = 0; m < M; ++m) {
            acc = fmaf(S[hh * M + m], __ldg(vp + (size_t)m * P.ldv), acc);
            if (RPE) {
                const float e = __ldg(ep + (size_t)m * C);
                g0 = fmaf(Sm[m], e, g0); g1 = fmaf(Sm[M + m], e, g1);
                g2 = fmaf(Sm[2 * M + m], e, g2); g3 = fmaf(Sm[3 * M + m], e, g3);
            }
        }
        P.hidden[(size_t)n * C + ch] = acc;
        if (RPE) {
            float* g = P.G + (size_t)n * H * C + ch;
            g[0] = g0; g[C] = g1; g[2 * C] = g2; g[3 * C] = g3;
        }
    }
}

template <int C, bool RPE>
int launch_attn(const GeoAttnParams& P, int batch, cudaStream_t st) {
    const size_t smem = (size_t)(RPE ? 2 : 1) * 4 * P.M * sizeof(float);
    if (smem > 48 * 1024) ROITR_CUDA(cudaFuncSetAttribute(geo_attention_kernel<C, RPE>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    geo_attention_kernel<C, RPE><<<dim3(P.N, batch), 256, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("geo_attention_kernel");
    return ROITR_OK;
}

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

extern "C" int roitr_geo_embedding(int N, int C, const float* pts, const int* nn3, const float* Wd, const float* bd,
                                   const float* Wa, const float* ba, const float* div_term, float sigma_d,
                                   float sigma_a, float* E, void* stream) {
    ROITR_CHECK_ARG(N >= 1 && pts && nn3 && Wd && bd && Wa && ba && div_term && E, "geo_embedding: bad arguments");
    ROITR_CHECK_ARG(C % GE_BN == 0, "geo_embedding: C must be a multiple of %d, got %d", GE_BN, C);
    ROITR_CHECK_ARG(((uintptr_t)Wd | (uintptr_t)Wa | (uintptr_t)E) % 16 == 0, "geo_embedding: 16-byte alignment");
    GeoEmbParams P;
    P.pts = pts; P.nn3 = nn3; P.Wd = Wd; P.bd = bd; P.Wa = Wa; P.ba = ba; P.div_term = div_term; P.E = E; P.N = N;
    P.C = C; P.sigma_d = sigma_d;
    P.factor_a = (float)(180.0 / ((double)sigma_a * 3.14159265358979323846));  // positional_encoding.py:99
    const long long npairs = (long long)N * N;
    dim3 grid((unsigned)ceil_div_ll(npairs, GE_PAIRS), C / GE_BN);
    geo_embedding_kernel<<<grid, GE_THREADS, 0, (cudaStream_t)stream>>>(P);
    ROITR_CHECK_LAUNCH("geo_embedding_kernel");
    return ROITR_OK;
}

extern "C" int roitr_geo_attention_batched(int batch, int N, int M, int C, int heads, const float* q, int ldq,
                                           long long q_bs, const float* k, int ldk, long long k_bs, const float* v,
                                           int ldv, long long v_bs, const float* E, const float* gq, const float* bp,
                                           float* hidden, float* G, void* stream) {
    ROITR_CHECK_ARG(heads == 4 && (C == 256 || C == 512), "geo_attention: heads=4, C in {256,512} only");
    ROITR_CHECK_ARG(batch >= 1 && N >= 1 && M >= 1 && q && k && v && hidden, "geo_attention: bad arguments");
    ROITR_CHECK_ARG(!E || (gq && bp && G && N == M), "geo_attention: RPE needs gq, bp, G and N == M");
    ROITR_CHECK_ARG(ldk % 4 == 0 && k_bs % 4 == 0 && ((uintptr_t)k % 16 == 0) && (!E || (uintptr_t)E % 16 == 0), "geo_attention: alignment");
    GeoAttnParams P;
    P.q = q; P.ldq = ldq; P.k = k; P.ldk = ldk; P.v = v; P.ldv = ldv; P.E = E; P.gq = gq; P.bp = bp; P.hidden = hidden;
    P.G = G; P.N = N; P.M = M; P.sqrt_c = sqrtf((float)(C / heads)); P.q_bs = q_bs; P.k_bs = k_bs; P.v_bs = v_bs;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 256) return E ? launch_attn<256, true>(P, batch, st) : launch_attn<256, false>(P, batch, st);
    return E ? launch_attn<512, true>(P, batch, st) : launch_attn<512, false>(P, batch, st);
}

extern "C" int roitr_geo_attention(int N, int M, int C, int heads, const float* q, int ldq, const float* k, int ldk,
                                   const float* v, int ldv, const float* E, const float* gq, const float* bp,
                                   float* hidden, float* G, void* stream) {
    return roitr_geo_attention_batched(1, N, M, C, heads, q, ldq, 0, k, ldk, 0, v, ldv, 0, E, gq, bp, hidden, G, stream);
}
