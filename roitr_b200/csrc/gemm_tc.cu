// Tensor-core dense layer for sm_100a: C[M,N] = (A [+ A2])[M,K] W[N,K]^T + bias (+ReLU), fp32 in / fp32 out,
// computed on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) with 3xTF32 split precision.
//
// Why split precision: the parity bar is fp32 results within 1e-4 of the reference. A single TF32 product keeps 11
// mantissa bits (~5e-4 relative on K=256 contractions: too coarse). Each fp32 operand is split exactly into
//     x = hi + lo,   hi = x with the low 13 mantissa bits cleared (a valid TF32 number),   lo = x - hi  (exact in fp32)
// and the product is accumulated in fp32 (TMEM) as  hi*hi + hi*lo + lo*hi  (lo is truncated to TF32 by the MMA, the
// lo*lo term is dropped): relative error ~2^-21, i.e. fp32-grade, at 3 MMAs per K-step (measured against an fp64
// reference in tests/test_gemm_gpu.py).
//
// Structure (one CTA = 128 threads = one 128 x BN output tile, BN in {64,128}):
//   loader (all threads)   global fp32 -> registers -> split -> 128B-swizzled K-major smem tiles A_hi/A_lo/B_hi/B_lo
//                          (K chunks of 32 floats = one 128-byte swizzle row), 2 stages
//   MMA (thread 0)         per chunk: 4 K-steps x 3 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8), then
//                          tcgen05.commit -> mbarrier; the next chunk's loads overlap the asynchronous MMAs
//   epilogue (4 warps)     tcgen05.ld 32x32b.x32 (warp w owns TMEM lanes 32w..32w+31 = rows), bias / ReLU, 128-byte row
//                          segments to global
// Operand rows may be gathered (a_index) or be the sum of two matrices (a_add), as in roitr_linear.
#include "../../include/roitr_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_THREADS = 128;
constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // floats per chunk = 128 bytes = one swizzle row

struct TcParams {
    int M, N, K;
    const float* A; const float* A2; int lda; const int* a_index;
    const float* W; int ldw; const float* bias;
    float* C; int ldc; int relu;
    // batched form (blockIdx.z = outer * inner + inner index): element strides of A / W / C per outer and inner index
    int inner;
    long long sA_o, sA_i, sW_o, sW_i, sC_o, sC_i;
    int w_transposed;       // W is given as Wt (K, >=N) row-major with leading dimension ldw: W[n][k] = Wt[k * ldw + n]
};

template <int BN, int TC_STAGES>
__global__ void __launch_bounds__(TC_THREADS) linear_tc_kernel(const TcParams P_) {
    TcParams P = P_;
    {
        const long long bo = blockIdx.z / P.inner, bi = blockIdx.z % P.inner;
        P.A += bo * P.sA_o + bi * P.sA_i;
        P.W += bo * P.sW_o + bi * P.sW_i;
        P.C += bo * P.sC_o + bi * P.sC_i;
    }
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte alignment: align manually (the launch reserves 1 KB of slack)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // per stage: A_hi (16 KB) | A_lo (16 KB) | B_hi (BN*128 B) | B_lo (BN*128 B)
    constexpr int A_BYTES = TC_BM * 128, B_BYTES = BN * 128, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    __shared__ __align__(8) uint64_t mma_bar[TC_STAGES];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;

    if (tid == 0) {
        for (int i = 0; i < TC_STAGES; ++i) mbar_init(&mma_bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = s_tmem;

    // loader row assignment: thread t owns A row t and B rows t, t+128 (BN=256) / t (BN<=128, t < BN)
    const int am = m0 + tid;
    long long arow = -1;
    if (am < P.M) arow = P.a_index ? (long long)__ldg(P.a_index + am) : (long long)am;
    const bool vecA = (P.lda % 4 == 0) && ((uintptr_t)P.A % 16 == 0) && (!P.A2 || (uintptr_t)P.A2 % 16 == 0);
    const bool vecW = (P.ldw % 4 == 0) && ((uintptr_t)P.W % 16 == 0);

    auto load_row = [&](const float* base, const float* base2, long long row, int ld, int k0, bool vec, float4 (&v)[8]) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int k = k0 + 4 * c;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) {
                const float* p = base + row * ld + k;
                if (vec && k + 3 < P.K) {
                    x = __ldg(reinterpret_cast<const float4*>(p));
                    if (base2) { const float4 y = __ldg(reinterpret_cast<const float4*>(base2 + row * ld + k)); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (k + e < P.K) t[e] = __ldg(p + e) + (base2 ? __ldg(base2 + row * ld + k + e) : 0.f);
                    x = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            v[c] = x;
        }
    };

    // transposed weight operand: thread t owns output column n0+t and reads Wt[k][n0+t] (coalesced across threads)
    auto load_row_t = [&](long long row, int k0, float4 (&v)[8]) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + 4 * c + e;
                if (row >= 0 && k < P.K) t[e] = __ldg(P.W + (long long)k * P.ldw + row);
            }
            v[c] = make_float4(t[0], t[1], t[2], t[3]);
        }
    };

    const int nk = (P.K + TC_BK - 1) / TC_BK;
    constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
    constexpr int BROWS = (BN + TC_THREADS - 1) / TC_THREADS;  // B rows per thread

    for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc % TC_STAGES;
        unsigned char* st = smem + buf * STAGE_BYTES;
        // global loads of this chunk are issued before waiting for the stage to drain
        float4 va[8], vb[BROWS][8];
        load_row(P.A, P.A2, arow, P.lda, kc * TC_BK, vecA, va);
#pragma unroll
        for (int j = 0; j < BROWS; ++j) {
            const int br = tid + j * TC_THREADS;
            const long long wrow = (br < BN && n0 + br < P.N) ? (long long)(n0 + br) : -1;
            if (P.w_transposed) load_row_t(wrow, kc * TC_BK, vb[j]);
            else load_row(P.W, nullptr, wrow, P.ldw, kc * TC_BK, vecW, vb[j]);
        }
        if (kc >= TC_STAGES) {  // the MMAs that read this stage (chunk kc - TC_STAGES) must have completed
            mbar_wait(&mma_bar[buf], ((kc / TC_STAGES) - 1) & 1);
            tc_fence_after();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) split_store(st, st + A_BYTES, tid, c, va[c]);
#pragma unroll
        for (int j = 0; j < BROWS; ++j) {
            const int br = tid + j * TC_THREADS;
            if (br < BN) {
#pragma unroll
                for (int c = 0; c < 8; ++c) split_store(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, br, c, vb[j][c]);
            }
        }
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {   // K = 8 tf32 = 32 bytes per MMA
                const uint64_t dah = make_desc_sw128(a_hi + ks * 32), dal = make_desc_sw128(a_lo + ks * 32);
                const uint64_t dbh = make_desc_sw128(b_hi + ks * 32), dbl = make_desc_sw128(b_lo + ks * 32);
                umma_tf32(tmem_d, dal, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);   // small terms first
                umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                umma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
            umma_commit(&mma_bar[buf]);
        }
    }
    // drain: the last commit covers every earlier MMA (commits complete in order)
    {
        const int last = nk - 1;
        mbar_wait(&mma_bar[last % TC_STAGES], (last / TC_STAGES) & 1);
        tc_fence_after();
    }
    // ---- epilogue: warp w reads TMEM lanes 32w..32w+31 (= output rows), 32 columns at a time ----
    const int row = m0 + warp * 32 + lane;
    const bool vecC = (P.ldc % 4 == 0) && ((uintptr_t)P.C % 16 == 0);
    // fragment layout (tc_common.cuh: tmem_ld_16x256b_x4): a quad of lanes holds 32 contiguous bytes of one row, so float2
    // stores fill whole sectors; the row layout below writes 16 bytes per lane into 32 different rows per instruction
    const bool frag = (P.ldc % 2 == 0) && (P.N % 2 == 0) && ((uintptr_t)P.C % 8 == 0) && ((uintptr_t)P.bias % 8 == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN && frag; c0 += 32) {
        if (n0 + c0 >= P.N) break;   // warp-uniform
        float v[2][16];
        const uint32_t tq = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        tmem_ld_16x256b_x4(tq, v[0]);
        tmem_ld_16x256b_x4(tq + (16u << 16), v[1]);
        const int cq = 2 * (lane & 3), rq = lane >> 2;
        float2 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + c0 + 8 * j + cq;
            b[j] = (P.bias && n < P.N) ? __ldg(reinterpret_cast<const float2*>(P.bias + n)) : make_float2(0.f, 0.f);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {                       // rows 8 i + rq of this warp's 32
            const int r = m0 + warp * 32 + 8 * i + rq;
            if (r < P.M) {
                float* dst = P.C + (long long)r * P.ldc + n0 + c0 + cq;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n0 + c0 + 8 * j + cq < P.N) {
                        float2 x = make_float2(v[i >> 1][4 * j + 2 * (i & 1)] + b[j].x, v[i >> 1][4 * j + 2 * (i & 1) + 1] + b[j].y);
                        if (P.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); }
                        *reinterpret_cast<float2*>(dst + 8 * j) = x;
                    }
                }
            }
        }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN && !frag; c0 += 32) {
        if (n0 + c0 >= P.N) break;   // warp-uniform
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (row < P.M) {
            float* dst = P.C + (long long)row * P.ldc + n0 + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = n0 + c0 + i;
                if (n < P.N) {
                    float x = v[i] + (P.bias ? __ldg(P.bias + n) : 0.f);
                    v[i] = P.relu ? fmaxf(x, 0.f) : x;
                }
            }
            if (vecC && n0 + c0 + 32 <= P.N) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + c0 + i < P.N) dst[i] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, BN);
}

template <int BN, int TC_STAGES>
int launch_tc(const TcParams& P, cudaStream_t st, int batch = 1) {
    constexpr int smem = TC_STAGES * (2 * TC_BM * 128 + 2 * BN * 128) + 1024;
    static bool attr_dev[ROITR_MAX_DEVICES] = {};
    bool& attr = attr_dev[roitr_cur_device()];
    if (!attr) {
        ROITR_CUDA(cudaFuncSetAttribute(linear_tc_kernel<BN, TC_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    dim3 grid(ceil_div(P.M, TC_BM), ceil_div(P.N, BN), batch);
    linear_tc_kernel<BN, TC_STAGES><<<grid, TC_THREADS, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("linear_tc_kernel");
    return ROITR_OK;
}

}  // namespace

extern "C" int roitr_gemm_tc_batched(int batch_outer, int batch_inner, int M, int N, int K, const float* A, int lda,
                                     long long sA_o, long long sA_i, const float* W, int ldw, long long sW_o, long long sW_i,
                                     int w_transposed, float* C, int ldc, long long sC_o, long long sC_i, void* stream) {
    ROITR_CHECK_ARG(batch_outer >= 1 && batch_inner >= 1 && (long long)batch_outer * batch_inner <= 65535,
                    "gemm_tc_batched: bad batch %d x %d", batch_outer, batch_inner);
    ROITR_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && A && W && C, "gemm_tc_batched: bad arguments M=%d N=%d K=%d", M, N, K);
    ROITR_CHECK_ARG(lda >= K && ldc >= N && ldw >= (w_transposed ? N : K), "gemm_tc_batched: bad leading dimensions");
    if (M == 0) return ROITR_OK;
    TcParams P;
    P.M = M; P.N = N; P.K = K; P.A = A; P.A2 = nullptr; P.lda = lda; P.a_index = nullptr; P.W = W; P.ldw = ldw; P.bias = nullptr;
    P.C = C; P.ldc = ldc; P.relu = 0;
    P.inner = batch_inner; P.sA_o = sA_o; P.sA_i = sA_i; P.sW_o = sW_o; P.sW_i = sW_i; P.sC_o = sC_o; P.sC_i = sC_i;
    P.w_transposed = w_transposed;
    cudaStream_t st = (cudaStream_t)stream;
    const int batch = batch_outer * batch_inner;
    // One operand stage and tiles of at most 128 columns: 49 / 65 KB of shared memory per CTA, so 4 / 3 CTAs share an SM and hide
    // each other's load -> split -> MMA -> epilogue chain (the products here are short: K = 64 or 312). Two stages and
    // 256-column tiles (one CTA per SM) measured 1.42 ms per step against 1.10 ms (profiles/r02f_gemm_batched_ab.txt).
    if (N <= 64) return launch_tc<64, 1>(P, st, batch);
    return launch_tc<128, 1>(P, st, batch);
}
