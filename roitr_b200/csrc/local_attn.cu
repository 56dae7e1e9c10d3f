// Fused local PPF attention over the k-neighbourhood (sm_100a).
//
// Replaces the ~20 eager ops of LocalRPEMultiHeadAttention.forward (model/transformer/attention.py:166-200) between
// the q/k/v projections and the output `linear`:
//     S_hj = ( q_h . k_{j,h} + q_h . p_{ij,h} ) / sqrt(c) ;  A = softmax_j S ;  out_h = sum_j A_hj (v_{j,h} + vp_{ij,h})
// with p_ij = proj_p(embedding.proj(ppf_ij)), vp_ij = proj_vp(embedding.proj(ppf_ij)).
//
// Exact algebraic fold (DESIGN.md "local fold"): PPFStructualEmbedding('local') is a bare Linear(4->C)
// (positional_encoding.py:68-70,78-79) and proj_p / proj_vp are Linear(C->C) with nothing in between, so
//     p_ij = Ap ppf_ij + cp,  vp_ij = Avp ppf_ij + cvp,   Ap = W_p W_e (C x 4), cp = W_p b_e + b_p   (same for vp)
//     q_h . p_ij,h = (Ap_h^T q_h) . ppf_ij + q_h . cp_h          -> 5 numbers per (query, head)
//     sum_j A_hj vp_ij,h = Avp_h (sum_j A_hj ppf_ij) + cvp_h      (softmax rows sum to 1)
// which removes the two (m k) x C x C GEMMs and the four (m,k,C) tensors the reference materialises; the kernel is a
// pure gather of K and V rows.
//
// One WARP per query. Lane l owns channels [l*CPL, (l+1)*CPL), CPL = C/32, so a head (c = C/4 channels) is 8 lanes and
// the per-head dot products are 3-step shuffle reductions. K/V rows are contiguous C*4 bytes: every gather is a fully
// coalesced 256 B..2 KB warp read, served from L2 (K,V of a level fit in the 126 MB L2).
#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

template <int CPL>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&r)[CPL]) {
    if constexpr (CPL == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
    }
}

__device__ __forceinline__ float head_sum(float v) {  // reduce over the 8 lanes of a head
    v += __shfl_xor_sync(FULL_MASK, v, 1);
    v += __shfl_xor_sync(FULL_MASK, v, 2);
    v += __shfl_xor_sync(FULL_MASK, v, 4);
    return v;
}

struct LocalAttnParams {
    const float* q; int ldq;       // (n, C) rows addressed through node_idx
    const float* k; int ldk;       // (n, C)
    const float* v; int ldv;       // (n, C)
    const int* node_idx;           // (m,) int32 or NULL (identity)
    const int* group_idx;          // (m, KNB) int32
    const float* ppf;              // (m, KNB, 4)
    const float* Ap; const float* cp;     // (C,4), (C)
    const float* Avp; const float* cvp;   // (C,4), (C)
    float* out;                    // (m, C)
    const float4* order;           // (m,) cell-sorted (x, y, z, query index) of the query set's grid, or NULL (natural order)
    int m;
    float sqrt_c;
};

template <int C, int KNB>
__global__ void __launch_bounds__(256) local_attn_kernel(const LocalAttnParams P) {
    constexpr int CPL = C / 32;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= P.m) return;
    // visiting order: with the cell-sorted order of the query set's grid the queries resident on an SM are spatial
    // neighbours, their k-neighbourhoods overlap and most K/V row gathers hit L1 instead of L2
    const int qi = P.order ? __float_as_int(__ldg(P.order + warp).w) : warp;
    const int node = P.node_idx ? __ldg(P.node_idx + qi) : qi;
    const int c0 = lane * CPL;

    float q[CPL];
    load_row<CPL>(P.q + (size_t)node * P.ldq + c0, q);

    // folded positional query terms (per head): qa[0..3] = Ap_h^T q_h, qb = q_h . cp_h
    float qa0 = 0.f, qa1 = 0.f, qa2 = 0.f, qa3 = 0.f, qb = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.Ap) + c0 + i);
        qa0 = fmaf(q[i], a.x, qa0); qa1 = fmaf(q[i], a.y, qa1); qa2 = fmaf(q[i], a.z, qa2); qa3 = fmaf(q[i], a.w, qa3);
        qb = fmaf(q[i], __ldg(P.cp + c0 + i), qb);
    }
    qa0 = head_sum(qa0); qa1 = head_sum(qa1); qa2 = head_sum(qa2); qa3 = head_sum(qa3); qb = head_sum(qb);

    const int* gi = P.group_idx + (size_t)qi * KNB;
    const float4* pf = reinterpret_cast<const float4*>(P.ppf) + (size_t)qi * KNB;
    const int my_nb = (lane < KNB) ? __ldg(gi + lane) : 0;

    // ---- scores ----
    // Lane l of a head computes the partial dot products of its CPL channels for all KNB neighbours; a halving butterfly
    // over the 8 lanes of the head (exchange half of the values with lane^4, then lane^2, lane^1) leaves lane li with the
    // COMPLETE dot product of neighbours li (and li + 8): 7 shuffles per 8 neighbours instead of 24, and the per-neighbour
    // scalar work (positional term, division, exp) is done once per neighbour instead of once per lane.
    constexpr int R = KNB / 8;
    const int li = lane & 7;
    float d[KNB];
#pragma unroll
    for (int j0 = 0; j0 < KNB; j0 += 4) {
        float kr[4][CPL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int nb = __shfl_sync(FULL_MASK, my_nb, j0 + u);
            load_row<CPL>(P.k + (size_t)nb * P.ldk + c0, kr[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) t = fmaf(q[i], kr[u][i], t);
            d[j0 + u] = t;
        }
    }
    {
        const bool b2 = li & 4, b1 = li & 2, b0 = li & 1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float* e = d + 8 * r;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float send = b2 ? e[t] : e[t + 4];
                const float keep = b2 ? e[t + 4] : e[t];
                e[t] = keep + __shfl_xor_sync(FULL_MASK, send, 4);
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float send = b1 ? e[t] : e[t + 2];
                const float keep = b1 ? e[t + 2] : e[t];
                e[t] = keep + __shfl_xor_sync(FULL_MASK, send, 2);
            }
            {
                const float send = b0 ? e[0] : e[1];
                const float keep = b0 ? e[1] : e[0];
                e[0] = keep + __shfl_xor_sync(FULL_MASK, send, 1);
            }
        }
    }
    // lane li now owns neighbours li + 8 r: score, softmax over the head's KNB values, positional weights
    float sc[R];
    float4 fr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fr[r] = __ldg(pf + li + 8 * r);
        const float sp = fmaf(qa3, fr[r].w, fmaf(qa2, fr[r].z, fmaf(qa1, fr[r].y, fmaf(qa0, fr[r].x, qb))));
        sc[r] = __fdiv_rn(d[8 * r] + sp, P.sqrt_c);  // attention.py:187: (e + p) / c ** 0.5
    }
    float mx = sc[0];
#pragma unroll
    for (int r = 1; r < R; ++r) mx = fmaxf(mx, sc[r]);
    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, 4));
    float den = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) { sc[r] = expf(sc[r] - mx); den += sc[r]; }
    den = head_sum(den);
    const float inv = 1.0f / den;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;  // sum_j A_j ppf_j (per head)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sc[r] *= inv;
        w0 = fmaf(sc[r], fr[r].x, w0); w1 = fmaf(sc[r], fr[r].y, w1); w2 = fmaf(sc[r], fr[r].z, w2); w3 = fmaf(sc[r], fr[r].w, w3);
    }
    w0 = head_sum(w0); w1 = head_sum(w1); w2 = head_sum(w2); w3 = head_sum(w3);

    // ---- value aggregate ----
    float acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
    const int head_base = lane & 24;
#pragma unroll
    for (int j0 = 0; j0 < KNB; j0 += 4) {
        float vr[4][CPL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int nb = __shfl_sync(FULL_MASK, my_nb, j0 + u);
            load_row<CPL>(P.v + (size_t)nb * P.ldv + c0, vr[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            const float a = __shfl_sync(FULL_MASK, sc[j >> 3], head_base | (j & 7));   // the weight lives in lane (head, j % 8)
#pragma unroll
            for (int i = 0; i < CPL; ++i) acc[i] = fmaf(a, vr[u][i], acc[i]);
        }
    }
    float* o = P.out + (size_t)qi * C + c0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.Avp) + c0 + i);
        o[i] = acc[i] + fmaf(a.w, w3, fmaf(a.z, w2, fmaf(a.y, w1, fmaf(a.x, w0, __ldg(P.cvp + c0 + i)))));
    }
}

template <int C, int KNB>
int launch(const LocalAttnParams& P, cudaStream_t st) {
    const int warps_per_cta = 8;
    local_attn_kernel<C, KNB><<<ceil_div(P.m, warps_per_cta), warps_per_cta * 32, 0, st>>>(P);
    ROITR_CHECK_LAUNCH("local_attn_kernel");
    return ROITR_OK;
}

}  // namespace

extern "C" int roitr_local_attention_ordered(int m, int C, int heads, int knb, const float* q, int ldq, const float* k,
                                             int ldk, const float* v, int ldv, const int* node_idx, const int* group_idx,
                                             const float* ppf, const float* Ap, const float* cp, const float* Avp,
                                             const float* cvp, const float* order_xyzi, float* out, void* stream) {
    ROITR_CHECK_ARG(heads == 4, "local_attention: 4 heads only (model/model.py:149), got %d", heads);
    ROITR_CHECK_ARG(q && k && v && group_idx && ppf && Ap && cp && Avp && cvp && out, "local_attention: null pointer");
    ROITR_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "local_attention: leading dims must be multiples of 4");
    ROITR_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)ppf | (uintptr_t)Ap | (uintptr_t)Avp) % 16 == 0,
                    "local_attention: pointers must be 16-byte aligned");
    ROITR_CHECK_ARG((uintptr_t)order_xyzi % 16 == 0, "local_attention: order must be 16-byte aligned");
    if (m == 0) return ROITR_OK;
    LocalAttnParams P;
    P.q = q; P.ldq = ldq; P.k = k; P.ldk = ldk; P.v = v; P.ldv = ldv; P.node_idx = node_idx; P.group_idx = group_idx;
    P.ppf = ppf; P.Ap = Ap; P.cp = cp; P.Avp = Avp; P.cvp = cvp; P.out = out; P.m = m;
    P.order = reinterpret_cast<const float4*>(order_xyzi);
    P.sqrt_c = sqrtf((float)(C / heads));
    cudaStream_t st = (cudaStream_t)stream;
#define LA(CV, KV) if (C == CV && knb == KV) return launch<CV, KV>(P, st)
    LA(64, 8); LA(64, 16); LA(128, 8); LA(128, 16); LA(256, 8); LA(256, 16); LA(512, 8); LA(512, 16);
#undef LA
    roitr_set_error("local_attention: unsupported C=%d k=%d (C in {64,128,256,512}, k in {8,16})", C, knb);
    return ROITR_ERR_UNSUPPORTED;
}

extern "C" int roitr_local_attention(int m, int C, int heads, int knb, const float* q, int ldq, const float* k, int ldk,
                                     const float* v, int ldv, const int* node_idx, const int* group_idx,
                                     const float* ppf, const float* Ap, const float* cp, const float* Avp,
                                     const float* cvp, float* out, void* stream) {
    return roitr_local_attention_ordered(m, C, heads, knb, q, ldq, k, ldk, v, ldv, node_idx, group_idx, ppf, Ap, cp, Avp, cvp,
                                         nullptr, out, stream);
}
