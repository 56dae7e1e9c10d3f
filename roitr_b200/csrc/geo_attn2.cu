// Global transformer attention, second generation (sm_100a): the dense contractions (Q K^T and P V) run on tcgen05
// (roitr_gemm_tc_batched, csrc/gemm_tc.cu); what is left is a streaming pass over the geometric embedding E.
//
// Replaces RPEMultiHeadAttention.forward (model/transformer/geoattention.py:101-136) and MultiHeadAttention.forward
// (geoattention.py:43-66) between the q/k/v projections and the output linear:
//
//   QK[b,h,n,m] = q_h[n] . k_h[m]                                   tcgen05 batched GEMM (per cloud and head)
//   self:  S = (QK + gq_h[n] . E[n,m] + q_h[n] . b_p,h) / sqrt(c)   this file, geo_self_scores_kernel
//          P  = softmax_m(S)                 -> global, for P V
//          G[n,h,:] = sum_m softmax_m(S with the diagonal removed)[m] E[n,m,:]      (position branch, :117-133)
//   cross: P = softmax_m(QK / sqrt(c))                              this file, softmax_rows_kernel
//   hidden[b,n,h*c:(h+1)*c] = P[b,h,n,:] V_h                        tcgen05 batched GEMM (V read transposed)
//
// geo_self_scores_kernel reads E exactly ONCE from HBM (the first generation, csrc/geo.cu, read it twice: 3.17 GB per
// 16-cloud launch at 42 % of HBM peak, profiles/r01e_geo_attention_kernel_raw.txt): one CTA per query row n streams
// the row's E[n,:,:] (M x C floats = 320 KB at M=312, C=256) through a two-stage shared-memory ring with bulk TMA copies
// (cp.async.bulk + mbarrier) in chunks of ~48 KB. Per chunk: (A) scores of the chunk's keys from shared memory,
// (B) chunk-level online softmax of the diagonal-free variant (running max / denominator per head, accumulators
// rescaled once per chunk, i.e. ~7 times per row), (C) G accumulation from the same shared-memory chunk. The full softmax
// that feeds P V is computed exactly (two-pass) at the end from the row's raw scores, which stay in shared memory.
#include <math_constants.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

constexpr int GA_THREADS = 256;
constexpr int GA_H = 4;
constexpr int GA_CHUNK_BYTES = 32 * 1024;    // x 2 stages + scores: 3 CTAs (24 warps) per SM at C = 256

struct SelfParams {
    const float* qk;       // (batch, H, N, M) raw q.k
    const float* q; int ldq; long long q_bs;    // (batch*N rows, C) view for q . b_p
    const float* E;        // (batch, N, M, C)
    const float* gq;       // (batch*N, H, C), row pitch ldgq floats
    int ldgq;
    const float* bp;       // (C)
    float* P;              // (batch, H, N, M) softmax with self
    float* G;              // (batch*N, H, C)
    int N, M;
    float sqrt_c;
};

// sum of 4 values over the warp with 6 shuffles instead of 20: after the two folding rounds each lane holds the partial
// sum of ONE head (head = bit4*2 + bit3 of the lane), then a 3-step butterfly inside the 8-lane group finishes it.
// Result: every lane of group g = lane>>3 ... holds the total of head ((lane>>4)&1)*2 + ((lane>>3)&1).
__device__ __forceinline__ float reduce4_to_head(float a0, float a1, float a2, float a3, int lane) {
    const bool hi16 = lane & 16;
    // round 1: lanes with bit4 = 0 keep heads {0,1}, bit4 = 1 keep heads {2,3}
    float s0 = hi16 ? a0 : a2, s1 = hi16 ? a1 : a3;        // what I send away
    float k0 = hi16 ? a2 : a0, k1 = hi16 ? a3 : a1;        // what I keep
    k0 += __shfl_xor_sync(FULL_MASK, s0, 16);
    k1 += __shfl_xor_sync(FULL_MASK, s1, 16);
    // round 2: bit3 = 0 keeps the first of the pair, bit3 = 1 the second
    const bool hi8 = lane & 8;
    float send = hi8 ? k0 : k1, keep = hi8 ? k1 : k0;
    keep += __shfl_xor_sync(FULL_MASK, send, 8);
    keep += __shfl_xor_sync(FULL_MASK, keep, 4);
    keep += __shfl_xor_sync(FULL_MASK, keep, 2);
    keep += __shfl_xor_sync(FULL_MASK, keep, 1);
    return keep;
}

template <int C>
__global__ void __launch_bounds__(GA_THREADS, (C <= 256 ? 3 : 2)) geo_self_scores_kernel(const SelfParams P) {
    constexpr int CPL = C / 32;            // channels per lane in phase A
    constexpr int CPT = C / GA_THREADS;    // channels per thread in phase C
    constexpr int H = GA_H;
    constexpr int CHK = GA_CHUNK_BYTES / (C * 4);   // keys per chunk
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);              // 2 x [CHK][C]
    float* S = ring + 2 * CHK * C;                                 // [H][M] raw scores of the row
    float* pn = S + H * P.M;                                       // [CHK][H] chunk weights of the diagonal-free softmax
    __shared__ __align__(8) uint64_t full[2];
    __shared__ float s_alpha[H], s_mx[H], s_den[H];

    const int n = blockIdx.x, b = blockIdx.y;
    const int N = P.N, M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* Erow = P.E + ((size_t)b * N + n) * (size_t)M * C;
    const size_t rowid = (size_t)b * N + n;
    const int nch = (M + CHK - 1) / CHK;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    if (tid < H) { s_mx[tid] = -CUDART_INF_F; s_den[tid] = 0.f; }
    __syncthreads();
    auto issue = [&](int c) {   // thread 0: chunk c -> stage c & 1
        const int rows = min(CHK, M - c * CHK);
        const uint32_t bytes = (uint32_t)rows * C * 4;
        mbar_expect_tx(&full[c & 1], bytes);
        tma_load_1d(ring + (size_t)(c & 1) * CHK * C, Erow + (size_t)c * CHK * C, bytes, &full[c & 1]);
    };
    if (tid == 0) { issue(0); if (nch > 1) issue(1); }

    // folded positional queries of this row: lane holds CPL consecutive channels of each head's gq
    const int c0 = lane * CPL;
    float gq[H][CPL];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int i = 0; i < CPL; ++i) gq[h][i] = __ldg(P.gq + rowid * (size_t)P.ldgq + h * C + c0 + i);
    // q_h . b_p,h per head (lanes of a head are the 8-lane groups: head = c0 / (C/H) = lane >> 3)
    float qb = 0.f;
    {
        const float* q = P.q + (size_t)b * P.q_bs + (size_t)n * P.ldq;
#pragma unroll
        for (int i = 0; i < CPL; ++i) qb = fmaf(__ldg(q + c0 + i), __ldg(P.bp + c0 + i), qb);
        qb += __shfl_xor_sync(FULL_MASK, qb, 1); qb += __shfl_xor_sync(FULL_MASK, qb, 2); qb += __shfl_xor_sync(FULL_MASK, qb, 4);
    }
    // the head whose total reduce4_to_head leaves in this lane, and that head's q.b_p
    const int myh = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    const float qb_h = __shfl_sync(FULL_MASK, qb, myh * 8);
    const float* qk_row = P.qk + (((size_t)b * H) * N + n) * M;     // + h * N * M

    float G[CPT][H];
#pragma unroll
    for (int u = 0; u < CPT; ++u)
#pragma unroll
        for (int h = 0; h < H; ++h) G[u][h] = 0.f;

    for (int c = 0; c < nch; ++c) {
        const int m0 = c * CHK, rows = min(CHK, M - m0);
        const float* Ec = ring + (size_t)(c & 1) * CHK * C;
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        // ---- A: scores of the chunk's keys ----
        for (int r = warp; r < rows; r += GA_THREADS / 32) {
            const float* e = Ec + (size_t)r * C + c0;
            float a[H] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < CPL / 4; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(e + 4 * i);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    a[h] = fmaf(gq[h][4 * i], t.x, a[h]); a[h] = fmaf(gq[h][4 * i + 1], t.y, a[h]);
                    a[h] = fmaf(gq[h][4 * i + 2], t.z, a[h]); a[h] = fmaf(gq[h][4 * i + 3], t.w, a[h]);
                }
            }
            const float sp = reduce4_to_head(a[0], a[1], a[2], a[3], lane);
            if ((lane & 7) == 0) {
                const int m = m0 + r;
                const float se = __ldg(qk_row + (size_t)myh * N * M + m);
                S[myh * M + m] = __fdiv_rn(se + (sp + qb_h), P.sqrt_c);
            }
        }
        __syncthreads();
        // ---- B: chunk-level online softmax of the diagonal-free variant (warp h < 4 owns head h) ----
        if (warp < H) {
            const int h = warp;
            float cm = -CUDART_INF_F;
            for (int r = lane; r < rows; r += 32)
                if (m0 + r != n) cm = fmaxf(cm, S[h * M + m0 + r]);
            cm = warp_max(cm);
            const float old = s_mx[h];
            const float nw = fmaxf(old, cm);
            float sum = 0.f;
            for (int r = lane; r < rows; r += 32) {
                const float w = (m0 + r == n || nw == -CUDART_INF_F) ? 0.f : expf(S[h * M + m0 + r] - nw);
                pn[r * H + h] = w;
                sum += w;
            }
            sum = warp_sum(sum);
            if (lane == 0) {
                const float alpha = (old == -CUDART_INF_F) ? 0.f : expf(old - nw);
                s_alpha[h] = alpha;
                s_mx[h] = nw;
                s_den[h] = s_den[h] * alpha + sum;
            }
        }
        __syncthreads();
        // ---- C: G[h][ch] = G * alpha + sum_r pn[r][h] E[r][ch] ----
        {
            const float4 al = make_float4(s_alpha[0], s_alpha[1], s_alpha[2], s_alpha[3]);
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                G[u][0] *= al.x; G[u][1] *= al.y; G[u][2] *= al.z; G[u][3] *= al.w;
            }
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                const float4 w = *reinterpret_cast<const float4*>(pn + r * H);
#pragma unroll
                for (int u = 0; u < CPT; ++u) {
                    const float e = Ec[(size_t)r * C + tid + u * GA_THREADS];
                    G[u][0] = fmaf(w.x, e, G[u][0]); G[u][1] = fmaf(w.y, e, G[u][1]);
                    G[u][2] = fmaf(w.z, e, G[u][2]); G[u][3] = fmaf(w.w, e, G[u][3]);
                }
            }
        }
        __syncthreads();                       // every thread is done with this stage (and with pn / s_alpha)
        if (tid == 0 && c + 2 < nch) issue(c + 2);
    }
    // ---- G / denominator ----
#pragma unroll
    for (int u = 0; u < CPT; ++u)
#pragma unroll
        for (int h = 0; h < H; ++h)
            P.G[(rowid * H + h) * C + tid + u * GA_THREADS] = G[u][h] / s_den[h];
    // ---- full softmax (with the diagonal) of the row's scores -> P (warp h < 4 owns head h) ----
    if (warp < H) {
        const int h = warp;
        float mx = -CUDART_INF_F, den = 0.f;
        for (int m = lane; m < M; m += 32) mx = fmaxf(mx, S[h * M + m]);
        mx = warp_max(mx);
        for (int m = lane; m < M; m += 32) den += expf(S[h * M + m] - mx);
        den = warp_sum(den);
        float* out = P.P + (((size_t)b * H + h) * N + n) * M;
        for (int m = lane; m < M; m += 32) out[m] = expf(S[h * M + m] - mx) / den;
    }
}

// ---- C = 256: barrier-free variant ------------------------------------------------------------------------------------
// ncu on the kernel above (profiles/r01j_geo_self_scores_kernel_*): 52 % issue slots, 36 % of the warp slots active and 3.1 TB/s
// of E - the three phases per chunk are separated by CTA barriers, phase B runs on 4 of the 8 warps, and E is read from
// shared memory twice (scores, then accumulation). Here every WARP owns a subset of the keys and carries its own online
// softmax: for a key it loads the row once (8 channels per lane), forms the four per-head dot products, folds them so that
// lane group g (8 lanes) holds head g's score, updates that head's running (max, denominator), broadcasts the four weights
// and accumulates G[4][8] from the registers that still hold the row. No barrier, no second pass over the chunk; a stage of
// the three-stage TMA ring is released by one mbarrier arrival per warp. The eight partial (max, denominator, G) states
// are merged once per row through shared memory.
constexpr int GV_STAGES = 3;
constexpr int GV_CHK = 32;      // keys per stage (32 KB at C = 256)

__global__ void __launch_bounds__(GA_THREADS, 2) geo_self_scores_v3_kernel(const SelfParams P) {
    constexpr int C = 256, H = GA_H, CPL = 8, WARPS = GA_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);              // GV_STAGES x [GV_CHK][C]; reused as the merge buffer
    float* S = ring + GV_STAGES * GV_CHK * C;                      // [H][M] raw q.k on entry, final scores on exit
    __shared__ __align__(8) uint64_t full[GV_STAGES], empty[GV_STAGES];
    __shared__ float s_mx[WARPS][H], s_den[WARPS][H];

    const int n = blockIdx.x, b = blockIdx.y;
    const int N = P.N, M = P.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* Erow = P.E + ((size_t)b * N + n) * (size_t)M * C;
    const size_t rowid = (size_t)b * N + n;
    const int nch = (M + GV_CHK - 1) / GV_CHK;

    if (tid == 0) {
        for (int i = 0; i < GV_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {   // thread 0: chunk c -> stage c % GV_STAGES
        const int rows = min(GV_CHK, M - c * GV_CHK);
        const uint32_t bytes = (uint32_t)rows * C * 4;
        const int st = c % GV_STAGES;
        mbar_expect_tx(&full[st], bytes);
        tma_load_1d(ring + (size_t)st * GV_CHK * C, Erow + (size_t)c * GV_CHK * C, bytes, &full[st]);
    };
    if (tid == 0)
        for (int c = 0; c < GV_STAGES && c < nch; ++c) issue(c);

    // raw q.k of the row (tcgen05 GEMM output) -> S, coalesced
    const float* qk_row = P.qk + (((size_t)b * H) * N + n) * M;
    for (int i = tid; i < H * M; i += GA_THREADS) S[i] = __ldg(qk_row + (size_t)(i / M) * N * M + (i % M));

    // Packed fp32 arithmetic (fma.rn.f32x2: two IEEE FMAs per instruction): channel pairs (2 j, 2 j + 1) of the lane's eight
    // channels share a register pair, which halves the instruction count of the two 32-FMA blocks of the key loop.
    const int c0 = lane * CPL;
    float2 gq[H][CPL / 2];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.gq + rowid * (size_t)P.ldgq + h * C + c0));
        const float4 d = __ldg(reinterpret_cast<const float4*>(P.gq + rowid * (size_t)P.ldgq + h * C + c0) + 1);
        gq[h][0] = make_float2(a.x, a.y); gq[h][1] = make_float2(a.z, a.w); gq[h][2] = make_float2(d.x, d.y); gq[h][3] = make_float2(d.z, d.w);
    }
    float qb = 0.f;
    {
        const float* q = P.q + (size_t)b * P.q_bs + (size_t)n * P.ldq;
#pragma unroll
        for (int i = 0; i < CPL; ++i) qb = fmaf(__ldg(q + c0 + i), __ldg(P.bp + c0 + i), qb);
        qb += __shfl_xor_sync(FULL_MASK, qb, 1); qb += __shfl_xor_sync(FULL_MASK, qb, 2); qb += __shfl_xor_sync(FULL_MASK, qb, 4);
    }
    const int myh = lane >> 3;                                      // the head whose total reduce4_to_head leaves in this lane
    const float qb_h = qb;                                          // lanes 8h .. 8h+7 hold channels of head h: qb IS q_h . b_p,h
    float2 G[H][CPL / 2];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int i = 0; i < CPL / 2; ++i) G[h][i] = make_float2(0.f, 0.f);
    const float inv_sqrt_c = 1.0f / P.sqrt_c;                       // sqrt(256 / 4) = 8: the product equals the reference's division bit for bit
    float mx = -CUDART_INF_F, den = 0.f;                            // running state of head myh (diagonal-free softmax)
    __syncthreads();                                                // S is in place

    for (int c = 0; c < nch; ++c) {
        const int st = c % GV_STAGES;
        const int m0 = c * GV_CHK, rows = min(GV_CHK, M - m0);
        const float* Ec = ring + (size_t)st * GV_CHK * C;
        mbar_wait(&full[st], (c / GV_STAGES) & 1);
        // (a two-phase variant - the four scores of the chunk first, one running-maximum update, then the four accumulations
        // with the rows re-read from shared memory - was measured: 0.96 ms per launch against 0.87 ms for this loop)
        for (int r = warp; r < rows; r += WARPS) {
            const int m = m0 + r;
            const float4 ea = *reinterpret_cast<const float4*>(Ec + (size_t)r * C + c0);
            const float4 eb = *reinterpret_cast<const float4*>(Ec + (size_t)r * C + c0 + 4);
            const float2 e[CPL / 2] = {make_float2(ea.x, ea.y), make_float2(ea.z, ea.w), make_float2(eb.x, eb.y), make_float2(eb.z, eb.w)};
            float a[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float2 t = __fmul2_rn(gq[h][0], e[0]);
#pragma unroll
                for (int i = 1; i < CPL / 2; ++i) t = __ffma2_rn(gq[h][i], e[i], t);
                a[h] = t.x + t.y;
            }
            const float sp = reduce4_to_head(a[0], a[1], a[2], a[3], lane);
            const float sc = (S[myh * M + m] + (sp + qb_h)) * inv_sqrt_c;
            if ((lane & 7) == 0) S[myh * M + m] = sc;
            if (m == n) continue;                                   // the position branch excludes the diagonal (warp-uniform)
            const float nm = fmaxf(mx, sc);
            const bool grew = nm > mx;
            // hardware ex2 on (sc - nm) log2 e: relative error 2^-22 + 6e-8 |sc - nm| log2 e (a few 1e-6 at the far tail, whose
            // weights are < 1e-20): two instructions instead of expf's eight, once per key and lane
            float w;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"((sc - nm) * 1.4426950408889634f));
            if (__any_sync(FULL_MASK, grew)) {                      // some head's running maximum moved: rescale (rare after the first keys)
                const float f = grew ? (mx == -CUDART_INF_F ? 0.f : expf(mx - nm)) : 1.f;
                den *= f;
                const float f0 = __shfl_sync(FULL_MASK, f, 0), f1 = __shfl_sync(FULL_MASK, f, 8);
                const float f2 = __shfl_sync(FULL_MASK, f, 16), f3 = __shfl_sync(FULL_MASK, f, 24);
#pragma unroll
                for (int i = 0; i < CPL / 2; ++i) {
                    G[0][i] = __fmul2_rn(G[0][i], make_float2(f0, f0)); G[1][i] = __fmul2_rn(G[1][i], make_float2(f1, f1));
                    G[2][i] = __fmul2_rn(G[2][i], make_float2(f2, f2)); G[3][i] = __fmul2_rn(G[3][i], make_float2(f3, f3));
                }
                mx = nm;
            }
            den += w;
            const float w0 = __shfl_sync(FULL_MASK, w, 0), w1 = __shfl_sync(FULL_MASK, w, 8);
            const float w2 = __shfl_sync(FULL_MASK, w, 16), w3 = __shfl_sync(FULL_MASK, w, 24);
            const float2 v0 = make_float2(w0, w0), v1 = make_float2(w1, w1), v2 = make_float2(w2, w2), v3 = make_float2(w3, w3);
#pragma unroll
            for (int i = 0; i < CPL / 2; ++i) {
                G[0][i] = __ffma2_rn(v0, e[i], G[0][i]); G[1][i] = __ffma2_rn(v1, e[i], G[1][i]);
                G[2][i] = __ffma2_rn(v2, e[i], G[2][i]); G[3][i] = __ffma2_rn(v3, e[i], G[3][i]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);                     // this warp is done with the stage
        if (tid == 0 && c + GV_STAGES < nch) {                      // refill it once every warp is
            mbar_wait(&empty[st], (c / GV_STAGES) & 1);
            issue(c + GV_STAGES);
        }
    }
    // ---- merge the eight warps' states: G = sum_w G_w exp(mx_w - mx*) / sum_w den_w exp(mx_w - mx*) ----
    if ((lane & 7) == 0) { s_mx[warp][myh] = mx; s_den[warp][myh] = den; }
    __syncthreads();                                                // also: every warp is past its last read of the ring
    float gmx = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) gmx = fmaxf(gmx, s_mx[w][myh]);
    float gden = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) gden += s_mx[w][myh] == -CUDART_INF_F ? 0.f : s_den[w][myh] * expf(s_mx[w][myh] - gmx);
    const float fm = mx == -CUDART_INF_F ? 0.f : expf(mx - gmx);    // this warp's factor for head myh
    const float f0 = __shfl_sync(FULL_MASK, fm, 0), f1 = __shfl_sync(FULL_MASK, fm, 8);
    const float f2 = __shfl_sync(FULL_MASK, fm, 16), f3 = __shfl_sync(FULL_MASK, fm, 24);
    float* mg = ring + (size_t)warp * H * C;                        // [warp][H][C]
#pragma unroll
    for (int i = 0; i < CPL / 2; i += 2) {
        *reinterpret_cast<float4*>(mg + 0 * C + c0 + 2 * i) = make_float4(G[0][i].x * f0, G[0][i].y * f0, G[0][i + 1].x * f0, G[0][i + 1].y * f0);
        *reinterpret_cast<float4*>(mg + 1 * C + c0 + 2 * i) = make_float4(G[1][i].x * f1, G[1][i].y * f1, G[1][i + 1].x * f1, G[1][i + 1].y * f1);
        *reinterpret_cast<float4*>(mg + 2 * C + c0 + 2 * i) = make_float4(G[2][i].x * f2, G[2][i].y * f2, G[2][i + 1].x * f2, G[2][i + 1].y * f2);
        *reinterpret_cast<float4*>(mg + 3 * C + c0 + 2 * i) = make_float4(G[3][i].x * f3, G[3][i].y * f3, G[3][i + 1].x * f3, G[3][i + 1].y * f3);
    }
    if ((lane & 7) == 0 && warp == 0) s_den[0][myh] = gden;          // every warp computed the same gden; publish one copy
    __syncthreads();
    for (int o = tid; o < H * C; o += GA_THREADS) {                  // o = h * C + channel
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) acc += ring[(size_t)w * H * C + o];
        P.G[rowid * H * C + o] = acc / s_den[0][o / C];
    }
    // ---- full softmax (with the diagonal) of the row's scores -> P (warp h < 4 owns head h) ----
    if (warp < H) {
        const int h = warp;
        float m2 = -CUDART_INF_F, d2 = 0.f;
        for (int m = lane; m < M; m += 32) m2 = fmaxf(m2, S[h * M + m]);
        m2 = warp_max(m2);
        for (int m = lane; m < M; m += 32) d2 += expf(S[h * M + m] - m2);
        d2 = warp_sum(d2);
        float* out = P.P + (((size_t)b * H + h) * N + n) * M;
        for (int m = lane; m < M; m += 32) out[m] = expf(S[h * M + m] - m2) / d2;
    }
}

// cross attention: P[row,:] = softmax(QK[row,:] / sqrt(c)); one warp per row of M scores, in place allowed. The scaled row
// (IEEE division, like the reference's `/ sqrt(c)`) is read ONCE into registers for rows of up to 512 scores.
__global__ void softmax_rows_kernel(long long rows, int M, const float* __restrict__ qk, float sqrt_c, float* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = qk + row * M;
    float* o = out + row * M;
    if (M <= 512) {
        float v[16];
        float mx = -CUDART_INF_F, den = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int m = lane + 32 * i;
            v[i] = m < M ? __fdiv_rn(__ldg(s + m), sqrt_c) : -CUDART_INF_F;
            mx = fmaxf(mx, v[i]);
        }
        mx = warp_max(mx);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = (lane + 32 * i < M) ? expf(v[i] - mx) : 0.f;
            den += v[i];
        }
        den = warp_sum(den);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (lane + 32 * i < M) o[lane + 32 * i] = v[i] / den;
        return;
    }
    float mx = -CUDART_INF_F, den = 0.f;
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, __fdiv_rn(s[m], sqrt_c));
    mx = warp_max(mx);
    for (int m = lane; m < M; m += 32) den += expf(__fdiv_rn(s[m], sqrt_c) - mx);
    den = warp_sum(den);
    for (int m = lane; m < M; m += 32) o[m] = expf(__fdiv_rn(s[m], sqrt_c) - mx) / den;
}

template <int C>
int launch_self(const SelfParams& P, int batch, cudaStream_t st) {
    constexpr int CHK = GA_CHUNK_BYTES / (C * 4);
    const size_t smem = (size_t)2 * CHK * C * 4 + (size_t)GA_H * P.M * 4 + (size_t)CHK * GA_H * 4;
    ROITR_CHECK_ARG(smem <= 226 * 1024, "geo_self_attention: %d keys do not fit the shared-memory score buffer", P.M);
    static size_t configured_dev[ROITR_MAX_DEVICES] = {};
    size_t& configured = configured_dev[roitr_cur_device()];
    if (smem > configured) {
        ROITR_CUDA(cudaFuncSetAttribute(geo_self_scores_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    geo_self_scores_kernel<C><<<dim3(P.N, batch), GA_THREADS, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("geo_self_scores_kernel");
    return ROITR_OK;
}

}  // namespace

extern "C" int roitr_geo_self_scores_ld(int batch, int N, int C, int heads, const float* qk, const float* q, int ldq,
                                     long long q_bs, const float* E, const float* gq, int ldgq, const float* bp, float* P, float* G,
                                     void* stream) {
    ROITR_CHECK_ARG(heads == GA_H && (C == 256 || C == 512), "geo_self_scores: heads=4, C in {256,512} only");
    ROITR_CHECK_ARG(batch >= 1 && batch <= 65535 && N >= 1 && qk && q && E && gq && bp && P && G, "geo_self_scores: bad arguments");
    ROITR_CHECK_ARG((uintptr_t)E % 16 == 0, "geo_self_scores: E must be 16-byte aligned");
    SelfParams S;
    S.qk = qk; S.q = q; S.ldq = ldq; S.q_bs = q_bs; S.E = E; S.gq = gq; S.ldgq = ldgq; S.bp = bp; S.P = P; S.G = G; S.N = N; S.M = N;
    S.sqrt_c = sqrtf((float)(C / heads));
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 256) {      // barrier-free kernel; C = 512 (factor-2 backbone) needs twice the registers per lane and keeps the chunk-synchronous one
        const size_t smem = (size_t)GV_STAGES * GV_CHK * 256 * 4 + (size_t)GA_H * S.M * 4;
        ROITR_CHECK_ARG(smem <= 226 * 1024 && (ldgq % 4) == 0 && (uintptr_t)gq % 16 == 0, "geo_self_scores: %d keys do not fit / gq alignment", S.M);
        static size_t configured_dev[ROITR_MAX_DEVICES] = {};
        size_t& configured = configured_dev[roitr_cur_device()];
        if (smem > configured) {
            ROITR_CUDA(cudaFuncSetAttribute(geo_self_scores_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        geo_self_scores_v3_kernel<<<dim3(S.N, batch), GA_THREADS, smem, st>>>(S);
        ROITR_CHECK_LAUNCH("geo_self_scores_v3_kernel");
        return ROITR_OK;
    }
    return launch_self<512>(S, batch, st);
}

extern "C" int roitr_geo_self_scores(int batch, int N, int C, int heads, const float* qk, const float* q, int ldq,
                                     long long q_bs, const float* E, const float* gq, const float* bp, float* P, float* G,
                                     void* stream) {
    return roitr_geo_self_scores_ld(batch, N, C, heads, qk, q, ldq, q_bs, E, gq, heads * C, bp, P, G, stream);
}

extern "C" int roitr_softmax_rows(long long rows, int M, const float* qk, float scale_div, float* out, void* stream) {
    ROITR_CHECK_ARG(rows >= 0 && M >= 1 && qk && out, "softmax_rows: bad arguments");
    if (rows == 0) return ROITR_OK;
    softmax_rows_kernel<<<(unsigned)ceil_div_ll(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(rows, M, qk, scale_div, out);
    ROITR_CHECK_LAUNCH("softmax_rows_kernel");
    return ROITR_OK;
}
