// Persistent warp-specialised tensor-core dense layer (sm_100a):  C[M,N] = (A [+ A2])[M,K] W[N,K]^T + bias (+ReLU)
// fp32 in / fp32 out, 3xTF32 split precision on tcgen05 (see gemm_tc.cu for the numerics).
//
// The dense layers of the forward are skinny (K = 64..512, N = 64..768) over very tall activations (up to 640 000 rows per
// step): they are memory-streaming problems, so the kernel is organised to keep HBM/L2 requests in flight at all times:
//
//   warps 0-3   EPILOGUE   tcgen05.ld the finished accumulator (warp w = TMEM lanes 32w..), bias / ReLU, 128-byte row
//                          segments to global; releases the accumulator (tmem_empty)
//   warps 4-7   A LOADERS  one activation row per thread: 8 independent 16-byte loads per K chunk (optional row gather /
//                          second addend), TF32 hi/lo split, 128B-swizzled shared-memory stores; run up to STAGES chunks
//                          ahead of the tensor core
//   warp 8      W PRODUCER the weights are static: pre-split and pre-swizzled at load time into per-(n-tile, K-chunk)
//                          blocks, fetched with ONE bulk TMA copy per chunk (cp.async.bulk + mbarrier complete_tx)
//   warp 9      MMA        one thread issues 4 K-steps x 3 tcgen05.mma.kind::tf32 per chunk and commits to the stage's
//                          "empty" barrier; a second commit per tile hands the accumulator to the epilogue
//
// CTAs are persistent (grid = #SMs): each walks tiles (row tile, n tile) with a static stride. Two TMEM accumulators
// alternate, so the epilogue of tile t overlaps the loads and MMAs of tile t+1.
#include "../../include/roitr_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int T2_THREADS = 320;
constexpr int T2_BM = 128, T2_BK = 32;
constexpr int A_HALF = T2_BM * 128;     // 16 KB: hi or lo of a 128-row x 32-float chunk

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


struct Tc2Params {
    int M, N, K;
    const float* A; const float* A2; int lda; const int* a_index;
    const float* wpack;     // [ceil(N/BN)][ceil(K/32)][hi|lo][BN rows x 128 B, SWIZZLE_128B], zero padded
    const float* bias;
    float* C; int ldc; int relu;
    int tiles_m, tiles_n, nkc;
    // fused row epilogue (streaming kernel, N <= one tile): out = act(LN(acc + bias + res_pre[idx]) * gamma + beta + res_post)
    const float* ln_gamma; const float* ln_beta; const float* res_pre; const int* res_pre_index; const float* res_post;
    int ln, ldr;            // ln != 0: LayerNorm over the N columns (eps 1e-5); ldr = row pitch of res_pre / res_post
};

// Epilogue of one 32x32 accumulator block held as "thread = row, v[j] = column j" (the tcgen05.ld 32x32b layout): bias,
// ReLU and the global store. The block is transposed through a per-warp shared-memory pad (row stride 36 floats: the
// 16-byte writes of 8 consecutive rows and the 16-byte reads of one row both touch all 32 banks once) so that 8 lanes
// write one 128-byte row segment with float4 stores: 8 store instructions per block, each covering 4 full lines.
constexpr int PAD_STRIDE = 36;
__device__ __forceinline__ void store_block32(const Tc2Params& P, float* pad, const float (&v)[32], int lane, int row0, int ncol0,
                                              const float* bias, const float* post) {
    const bool vec = (P.ldc % 4 == 0) && (P.N % 4 == 0) && ((uintptr_t)P.C % 16 == 0) && (!post || P.ldr % 4 == 0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(pad + lane * PAD_STRIDE + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    if (vec) {
        const int sub = lane >> 3, n = ncol0 + 4 * (lane & 7);
        if (n < P.N) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + n));
            float* dst = P.C + (long long)(row0 + sub) * P.ldc + n;
            const float* pp = post ? post + (long long)(row0 + sub) * P.ldr + n : nullptr;
            const long long step = 4ll * P.ldc, pstep = 4ll * P.ldr;
            float4 r[8];                                     // all eight residual segments in flight before the first use
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pp && row0 + 4 * i + sub < P.M) r[i] = __ldg(reinterpret_cast<const float4*>(pp + i * pstep));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 x = *reinterpret_cast<const float4*>(pad + (4 * i + sub) * PAD_STRIDE + 4 * (lane & 7));
                x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                x.x += r[i].x; x.y += r[i].y; x.z += r[i].z; x.w += r[i].w;
                if (P.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                if (row0 + 4 * i + sub < P.M) *reinterpret_cast<float4*>(dst + i * step) = x;
            }
        }
    } else {
        const int n = ncol0 + lane;
        if (n < P.N) {
            const float bv = bias ? __ldg(bias + n) : 0.f;
            for (int i = 0; i < 32; ++i) {
                if (row0 + i >= P.M) break;
                float x = pad[i * PAD_STRIDE + lane] + bv;
                if (post) x += __ldg(post + (long long)(row0 + i) * P.ldr + n);
                if (P.relu) x = fmaxf(x, 0.f);
                P.C[(long long)(row0 + i) * P.ldc + n] = x;
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void store_block32(const Tc2Params& P, float* pad, const float (&v)[32], int lane, int row0, int ncol0) {
    store_block32(P, pad, v, lane, row0, ncol0, P.bias, nullptr);
}

#ifndef EPI_FRAG
#define EPI_FRAG 1
#endif
// The same 32x32 block straight from TMEM in the fragment layout (tc_common.cuh: tmem_ld_16x256b_x4): bias / ReLU and float2
// stores, 8 rows x 32 bytes per instruction, no shared-memory transposition. `tq` = TMEM address of the quadrant's first lane
// at the block's first column. Needs even N / ldc and 8-byte aligned C / bias (checked per launch by the caller).
__device__ __forceinline__ bool frag_store_ok(const Tc2Params& P) {
    return EPI_FRAG && (P.ldc % 2 == 0) && (P.N % 2 == 0) && ((uintptr_t)P.C % 8 == 0) && ((uintptr_t)P.bias % 8 == 0);
}
__device__ __forceinline__ void store_block32_frag(const Tc2Params& P, uint32_t tq, int lane, int row0, int ncol0) {
    float v[2][16];
    tmem_ld_16x256b_x4(tq, v[0]);
    tmem_ld_16x256b_x4(tq + (16u << 16), v[1]);
    const int cq = 2 * (lane & 3), rq = lane >> 2;
    float2 b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = ncol0 + 8 * j + cq;
        b[j] = (P.bias && n < P.N) ? __ldg(reinterpret_cast<const float2*>(P.bias + n)) : make_float2(0.f, 0.f);
    }
    tmem_ld_wait();
#pragma unroll
    for (int hl = 0; hl < 2; ++hl)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = row0 + 16 * hl + 8 * h + rq;
            if (row < P.M) {
                float* dst = P.C + (long long)row * P.ldc + ncol0 + cq;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ncol0 + 8 * j + cq < P.N) {
                        float2 x = make_float2(v[hl][4 * j + 2 * h] + b[j].x, v[hl][4 * j + 2 * h + 1] + b[j].y);
                        if (P.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); }
                        *reinterpret_cast<float2*>(dst + 8 * j) = x;
                    }
                }
            }
        }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(T2_THREADS, 1) linear_tc2_kernel(const Tc2Params P) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int B_HALF = BN * 128;
    constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_HALF;
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_pad[4][32 * PAD_STRIDE];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 128 + 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int total_tiles = P.tiles_m * P.tiles_n;
    const int nkc = P.nkc;

    if (warp >= 4 && warp < 8) {
        // ================================================= A loaders =================================================
        const int r = tid - 128;                                   // row of the tile owned by this thread
        const bool vec = (P.lda % 4 == 0) && ((uintptr_t)P.A % 16 == 0) && (P.K % 4 == 0);
        if (vec && !P.A2 && !P.a_index) {
            // Streaming path: cp.async (LDGSTS) 16-byte copies straight into the swizzled operand tile, LOOK chunks ahead
            // of the split pass, so STAGES-1 x 16 KB of loads are in flight per SM without holding registers. Each thread
            // copies and later splits ONLY its own row, so per-thread cp.async groups are all the synchronisation needed.
            constexpr int LOOK = STAGES - 1;
            // this thread's chunk sequence: (tile, kc) pairs in order; `issue` runs LOOK steps ahead of `it`
            int i_tile = blockIdx.x, i_kc = 0;
            uint32_t i_it = 0;
            auto issue_one = [&]() {
                if (i_tile < total_tiles) {
                    const int st = i_it % STAGES;
                    mbar_wait(&empty_bar[st], ((i_it / STAGES) & 1) ^ 1);
                    const int am = (i_tile / P.tiles_n) * T2_BM + r;
                    const uint32_t dst = smem_u32(smem + st * STAGE_BYTES);
                    const float* src = P.A + (long long)(am < P.M ? am : 0) * P.lda + i_kc * T2_BK;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int k = i_kc * T2_BK + 4 * c;
                        cp_async16(dst + sw128_offset(r, c), src + 4 * c, (am < P.M && k < P.K) ? 16u : 0u);
                    }
                    ++i_it;
                    if (++i_kc == nkc) { i_kc = 0; i_tile += gridDim.x; }
                }
                cp_async_commit();                                 // always commit: keeps the group count uniform
            };
#pragma unroll
            for (int j = 0; j < LOOK; ++j) issue_one();
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    issue_one();
                    cp_async_wait<LOOK>();                         // this thread's copies of chunk `it` have landed
                    unsigned char* a_hi = smem + (it % STAGES) * STAGE_BYTES;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {                  // in-place split: raw -> TF32 hi (rounded), lo = raw - hi
                        const uint32_t o = sw128_offset(r, c);
                        split_store(a_hi, a_hi + A_HALF, r, c, *reinterpret_cast<const float4*>(a_hi + o));
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&full_bar[it % STAGES]);
                }
            }
            cp_async_wait<0>();
        } else {
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int tm = tile / P.tiles_n;
            const int am = tm * T2_BM + r;
            long long arow = -1;
            if (am < P.M) arow = P.a_index ? (long long)__ldg(P.a_index + am) : (long long)am;
            const bool vec2 = vec && (!P.A2 || (uintptr_t)P.A2 % 16 == 0);
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                const int st = it % STAGES;
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {                     // loads first: they fly while we wait for the stage
                    const int k = kc * T2_BK + 4 * c;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (arow >= 0) {
                        const float* p = P.A + arow * P.lda + k;
                        if (vec2 && k + 3 < P.K) {
                            x = __ldg(reinterpret_cast<const float4*>(p));
                            if (P.A2) { const float4 y = __ldg(reinterpret_cast<const float4*>(P.A2 + arow * P.lda + k)); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
                        } else {
                            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (k + e < P.K) t[e] = __ldg(p + e) + (P.A2 ? __ldg(P.A2 + arow * P.lda + k + e) : 0.f);
                            x = make_float4(t[0], t[1], t[2], t[3]);
                        }
                    }
                    v[c] = x;
                }
                mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);
                unsigned char* a_hi = smem + st * STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 8; ++c) split_store(a_hi, a_hi + A_HALF, r, c, v[c]);
                fence_proxy_async_smem();
                mbar_arrive(&full_bar[st]);
            }
        }
        }
    } else if (warp == 8) {
        // ================================================= W producer ================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tn = tile % P.tiles_n;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int st = it % STAGES;
                    mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[st], 2 * B_HALF);
                    tma_load_1d(smem + st * STAGE_BYTES + 2 * A_HALF, P.wpack + ((size_t)tn * nkc + kc) * (2 * B_HALF / 4),
                                2 * B_HALF, &full_bar[st]);
                }
            }
        }
    } else if (warp == 9) {
        // ================================================= MMA issuer ================================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(T2_BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);          // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(acc * BN);
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int st = it % STAGES;
                    mbar_wait(&full_bar[st], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t ah = smem_u32(smem + st * STAGE_BYTES), al = ah + A_HALF, bh = ah + 2 * A_HALF, bl = bh + B_HALF;
#pragma unroll
                    for (int ks = 0; ks < T2_BK / 8; ++ks) {
                        const uint64_t dah = make_desc_sw128(ah + ks * 32), dal = make_desc_sw128(al + ks * 32);
                        const uint64_t dbh = make_desc_sw128(bh + ks * 32), dbl = make_desc_sw128(bl + ks * 32);
                        umma_tf32(d, dal, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty_bar[st]);                               // stage reusable when these MMAs retire
                }
                umma_commit(&tmem_full[acc]);                                  // accumulator ready for the epilogue
            }
        }
    } else {
        // ================================================= epilogue (warps 0-3) ======================================
        // tcgen05.ld gives thread `lane` the 32 consecutive columns of ITS row; written directly that is 32 scattered
        // 16-byte pieces per store instruction (32 LSU wavefronts). Each warp instead transposes the 32x32 block through
        // its private shared-memory pad so that one store instruction writes 128 contiguous bytes of one row.
        float* pad = s_pad[warp];
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int tm = tile / P.tiles_n, tn = tile % P.tiles_n;
            const int acc = tcount & 1;
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            tc_fence_after();
            const int row0 = tm * T2_BM + warp * 32;
            const int n0 = tn * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= P.N) break;
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                store_block32(P, pad, v, lane, row0, n0 + c0);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * BN);
}

// ---- streaming variant ------------------------------------------------------------------------------------------------
// Same roles, but the activation loads are decoupled from the tensor core: raw fp32 chunks land in a RAW-deep ring with
// cp.async (16 KB each; every thread copies and later splits ONLY its own 16-byte pieces, so the ring needs no barrier and
// its look-ahead never waits for an MMA), and a separate OPS-deep ring holds the split hi/lo operands. The split of chunk
// i therefore overlaps the MMAs of chunk i-1 and the HBM latency of chunks i+1 .. i+RAW-1. Pieces are mapped so that 8
// consecutive lanes cover one 128-byte row segment (4 full lines per warp instruction) and 8 warps share the split pass.
//
// LW = loader warps, MINB = CTAs per SM the register allocation is capped for. The "light" instantiations (one operand
// stage, two raw stages, 4 loader warps, <= 64 registers: ~115 KB of shared memory and 20 K registers per CTA) exist so
// that a dense layer can share an SM with the CTAs of other streams' kernels (FPS clusters, kNN, attention) instead of
// waiting for 215 KB of shared memory and 38 K registers to become free on every SM at once.
constexpr int RAW_BYTES = T2_BM * 128;

// Epilogue of one finished accumulator tile (rows tm*128 .. +127, columns tn*BN .. ) held in TMEM at `tbase` (this warp's 32
// lanes): plain bias / ReLU store, or the fused row epilogue. `next_tm` = the row tile this CTA processes next (-1: none),
// whose residual rows are prefetched into L2. Used by the streaming kernel (linear_tc3).
template <int BN, int LNW, bool FRAG>
__device__ __forceinline__ void tile_epilogue(const Tc2Params& P, float* pad, const float (*s_ln)[LNW], int warp, int lane, int tm,
                                              int tn, int next_tm, uint32_t tbase, int cb0 = 0, int cbstep = 1,
                                              float2* stat = nullptr) {
    // `warp` = the TMEM lane quadrant (rows 32 warp .. +31 of the tile); in the plain epilogue this warp stores the 32-column
    // blocks cb0, cb0 + cbstep, ... (two warps per quadrant split the blocks even / odd)
    const int row0 = tm * T2_BM + warp * 32;
    const int n0 = tn * BN;
    if (P.ln) {
        // Fused row epilogue (host guarantees N % 32 == 0 and the whole row in TMEM at tbase: one n-tile, or two side by side
        // in the 512-column layout of the deep configuration). After tcgen05.ld a thread holds 32 columns
        // of ITS row, so the LayerNorm statistics are thread-local: pass A adds bias and the (optionally gathered)
        // pre-norm residual and writes the row back to TMEM while summing it, pass B sums the squared deviations,
        // pass C normalises and hands 32x32 blocks to the transposed store (post-norm residual, ReLU, float4 rows).
        const int row = row0 + lane;
        const bool valid = row < P.M;
        {   // the residual rows of this CTA's NEXT tile: requested now (L2 prefetch), consumed one tile later - each
            // thread reads its own row, so without this the 700-cycle miss latency is exposed twice per tile
            const int nrow = next_tm * T2_BM + warp * 32 + lane;
            if (next_tm >= 0 && nrow < P.M) {
                if (P.res_pre) {
                    const float* pr = P.res_pre + (P.res_pre_index ? (long long)__ldg(P.res_pre_index + nrow) : (long long)nrow) * P.ldr;
                    for (int c = 0; c < P.N; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + c));
                }
                if (P.res_post) {
                    const float* pq = P.res_post + (long long)nrow * P.ldr;
                    for (int c = 0; c < P.N; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pq + c));
                }
            }
        }
        const float* pre = nullptr;
        if (P.res_pre && valid)
            pre = P.res_pre + (P.res_pre_index ? (long long)__ldg(P.res_pre_index + row) : (long long)row) * P.ldr;
        // one pass for both moments, shifted by the row's first value so that the subtraction sum2/N - (sum1/N)^2
        // never cancels (|mean - shift| is of the order of the standard deviation)
        // With `stat` (two warps per quadrant, cbstep == 2) each warp owns every other 32-column block, takes the moments of
        // ITS half of the row around its own shift, and the two halves are merged (equal counts) after a 64-thread barrier.
        float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll 1
        for (int c0 = 32 * cb0; c0 < P.N; c0 += 32 * cbstep) {
            float v[32];
            tmem_ld32(tbase + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(&s_ln[0][c0 + 4 * j]);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
            if (pre) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(pre + c0) + j);
                    v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                }
            }
            if (c0 == 32 * cb0) shift = v[0];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d0 = v[4 * j] - shift, d1 = v[4 * j + 1] - shift, d2 = v[4 * j + 2] - shift, d3 = v[4 * j + 3] - shift;
                a0 += d0; a1 += d1; a2 += d2; a3 += d3;
                q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
            }
            s1 += (a0 + a1) + (a2 + a3);
            s2 += (q0 + q1) + (q2 + q3);
            tmem_st32(tbase + c0, v);
        }
        float mean, sq;
        if (stat) {
            const float nw = 0.5f * (float)P.N;                  // columns per warp
            const float dmw = s1 / nw;
            stat[cb0 * 32 + lane] = make_float2(shift + dmw, fmaxf(s2 - s1 * dmw, 0.f));
            asm volatile("bar.sync %0, 64;" ::"r"(1 + warp) : "memory");       // the two warps of this quadrant
            const float2 a = stat[lane], b = stat[32 + lane];    // same expression order in both warps: identical mean / rstd
            const float d = a.x - b.x;
            mean = 0.5f * (a.x + b.x);
            sq = a.y + b.y + d * d * (0.5f * nw);
        } else {
            const float dm = s1 / (float)P.N;
            mean = shift + dm;
            sq = fmaxf(s2 - s1 * dm, 0.f);                       // sum (x - mean)^2 = sum d^2 - (sum d)^2 / N
        }
        const float rstd = rsqrtf(sq / (float)P.N + 1e-5f);
#pragma unroll 1
        for (int c0 = 32 * cb0; c0 < P.N; c0 += 32 * cbstep) {
            float v[32];
            tmem_ld32(tbase + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 g = *reinterpret_cast<const float4*>(&s_ln[1][c0 + 4 * j]);
                const float4 b = *reinterpret_cast<const float4*>(&s_ln[2][c0 + 4 * j]);
                v[4 * j] = (v[4 * j] - mean) * rstd * g.x + b.x; v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g.y + b.y;
                v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g.z + b.z; v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g.w + b.w;
            }
            store_block32(P, pad, v, lane, row0, c0, nullptr, P.res_post);
        }
    } else if (FRAG && frag_store_ok(P)) {
#pragma unroll 1
        for (int c0 = 32 * cb0; c0 < BN; c0 += 32 * cbstep) {
            if (n0 + c0 >= P.N) break;
            store_block32_frag(P, tbase + c0, lane, row0, n0 + c0);
        }
    } else {
#pragma unroll 1
        for (int c0 = 32 * cb0; c0 < BN; c0 += 32 * cbstep) {
            if (n0 + c0 >= P.N) break;
            float v[32];
            tmem_ld32(tbase + c0, v);
            store_block32(P, pad, v, lane, row0, n0 + c0);
        }
    }
}

#ifndef LN_SPLIT
#define LN_SPLIT 1
#endif
template <int BN, int OPS, int RAW, int LW, int MINB, bool GATHER>
__global__ void __launch_bounds__(128 + LW * 32 + 64 + (MINB == 1 ? 128 : 0), MINB) linear_tc3_kernel(const Tc2Params P) {
    // EW epilogue warps: the deep configuration (one CTA per SM) runs EIGHT - the round-1 ablation (profiles/r01g_ablate_gemm.txt)
    // shows the C stores and the main loop adding up instead of overlapping (0.247 ms = 0.117 + 0.130 at (640000, 192, 64)):
    // four warps walking 32x32 blocks one after the other (tcgen05.ld -> wait -> transpose -> stores) are the longest chain of
    // the tile. Warps w and w + 4 (counted among the epilogue warps) share TMEM lane quadrant w % 4 and take the even / odd
    // 32-column blocks. The fused LayerNorm epilogue needs a whole row per thread and stays on the first four.
    constexpr int EW = (MINB == 1) ? 8 : 4;
    constexpr int T3_LOADERS = LW * 32;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int B_HALF = BN * 128;
    constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_HALF;
    __shared__ __align__(8) uint64_t full_bar[OPS], empty_bar[OPS], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_pad[EW][32 * PAD_STRIDE];
    // Row groups: with the fused LayerNorm over TWO n-tiles (N <= 256 at BN = 128, deep configuration only) a CTA computes both
    // n-tiles of a row tile back to back into the two halves of one 256-column accumulator slot (TMEM: 2 slots = 512 columns),
    // and the epilogue sees the whole row. gshift = log2(tiles per row group); iteration i of this CTA is tile_of(i).
    constexpr int LNW = (BN == 128 && MINB == 1) ? 256 : BN;
    __shared__ __align__(16) float s_ln[3][LNW];       // bias, gamma, beta of the fused LayerNorm epilogue
    __shared__ float2 s_stat[EW == 8 ? 4 : 1][2][64];  // per quadrant, per tile parity: (mean, M2) of each warp's half row

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < OPS; ++i) { mbar_init(&full_bar[i], T3_LOADERS + 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EW * 32); }
        mbar_fence_init();
    }
    const int gshift = (LNW > BN && P.ln && P.tiles_n == 2) ? 1 : 0;
    const uint32_t tmem_cols = (uint32_t)(2 * BN) << gshift;
    if (warp == 0) tmem_alloc(&s_tmem, tmem_cols);
    if (P.ln && tid < 128) {
        for (int i = tid; i < LNW; i += 128) {
            const bool in = i < P.N;
            s_ln[0][i] = (in && P.bias) ? __ldg(P.bias + i) : 0.f;
            s_ln[1][i] = in ? __ldg(P.ln_gamma + i) : 0.f;
            s_ln[2][i] = in ? __ldg(P.ln_beta + i) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int total_tiles = P.tiles_m * P.tiles_n;
    const int nkc = P.nkc;
    constexpr int W_LOADER0 = 4, W_PRODUCER = 4 + T3_LOADERS / 32, W_MMA = W_PRODUCER + 1;
    const int gmask = (1 << gshift) - 1;
    auto tile_of = [&](uint32_t i) -> int {            // >= total_tiles: this CTA is done
        return (int)(((blockIdx.x + (i >> gshift) * gridDim.x) << gshift) + (i & gmask));
    };

    if (warp >= W_LOADER0 && warp < W_PRODUCER) {
        // ================================================= A loaders =================================================
        const int t = tid - 128;
        constexpr int RSTEP = LW * 4, NPIECE = T2_BM / RSTEP;           // pieces (r0 + RSTEP j, c), j = 0..NPIECE-1
        const int c = t & 7, r0 = t >> 3;
        unsigned char* raw = smem + OPS * STAGE_BYTES;
        const uint32_t raw_u32 = smem_u32(raw);
        uint32_t i_i = 0;
        int i_tile = tile_of(0), i_kc = 0;
        uint32_t i_it = 0;
        auto issue_one = [&]() {
            if (i_tile < total_tiles) {
                const uint32_t dst = raw_u32 + (i_it % RAW) * RAW_BYTES + (uint32_t)(r0 * 128 + c * 16);
                const int k = i_kc * T2_BK + 4 * c;
                const int am0 = (i_tile / P.tiles_n) * T2_BM + r0;
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) {
                    const int am = am0 + RSTEP * j;
                    const bool ok = am < P.M && k < P.K;
                    const long long arow = (GATHER && ok) ? (long long)__ldg(P.a_index + am) : (long long)am;   // GATHER: rows through a_index
                    cp_async16(dst + j * RSTEP * 128, P.A + (ok ? arow * P.lda + k : 0ll), ok ? 16u : 0u);
                }
                ++i_it;
                if (++i_kc == nkc) { i_kc = 0; i_tile = tile_of(++i_i); }
            }
            cp_async_commit();                                          // always commit: keeps the group count uniform
        };
#pragma unroll
        for (int j = 0; j < RAW - 1; ++j) issue_one();
        uint32_t it = 0;
        for (uint32_t i = 0; tile_of(i) < total_tiles; ++i) {
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                issue_one();                                            // refills the slot this thread drained last iteration
                cp_async_wait<RAW - 1>();                               // this thread's pieces of chunk `it` have landed
                const unsigned char* src = raw + (it % RAW) * RAW_BYTES + r0 * 128 + c * 16;
                float4 v[NPIECE];
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) v[j] = *reinterpret_cast<const float4*>(src + j * RSTEP * 128);
                const int st = it % OPS;
                mbar_wait(&empty_bar[st], ((it / OPS) & 1) ^ 1);        // MMAs of chunk it - OPS have retired
                unsigned char* a_hi = smem + st * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) split_store(a_hi, a_hi + A_HALF, r0 + RSTEP * j, c, v[j]);
                fence_proxy_async_smem();
                mbar_arrive(&full_bar[st]);
            }
        }
        cp_async_wait<0>();
    } else if (warp == W_PRODUCER) {
        // ================================================= W producer ================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t i = 0; tile_of(i) < total_tiles; ++i) {
                const int tn = tile_of(i) % P.tiles_n;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int st = it % OPS;
                    mbar_wait(&empty_bar[st], ((it / OPS) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[st], 2 * B_HALF);
                    tma_load_1d(smem + st * STAGE_BYTES + 2 * A_HALF, P.wpack + ((size_t)tn * nkc + kc) * (2 * B_HALF / 4),
                                2 * B_HALF, &full_bar[st]);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================================= MMA issuer ================================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(T2_BM, BN);
            uint32_t it = 0;
            for (uint32_t i = 0; tile_of(i) < total_tiles; ++i) {
                const uint32_t r = i >> gshift, sub = i & gmask;               // row group of this CTA, n-tile within it
                const int acc = r & 1;
                if (sub == 0) {
                    mbar_wait(&tmem_empty[acc], ((r >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                const uint32_t d = tmem + (uint32_t)((acc * BN) << gshift) + sub * BN;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int st = it % OPS;
                    mbar_wait(&full_bar[st], (it / OPS) & 1);
                    tc_fence_after();
                    const uint32_t ah = smem_u32(smem + st * STAGE_BYTES), al = ah + A_HALF, bh = ah + 2 * A_HALF, bl = bh + B_HALF;
#pragma unroll
                    for (int ks = 0; ks < T2_BK / 8; ++ks) {
                        const uint64_t dah = make_desc_sw128(ah + ks * 32), dal = make_desc_sw128(al + ks * 32);
                        const uint64_t dbh = make_desc_sw128(bh + ks * 32), dbl = make_desc_sw128(bl + ks * 32);
                        umma_tf32(d, dal, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty_bar[st]);
                }
                if (sub == (uint32_t)gmask) umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp < 4 || warp > W_MMA) {
        // ================================================= epilogue (warps 0-3 and, with EW = 8, the last four) ========
        const int ew = warp < 4 ? warp : 4 + (warp - W_MMA - 1);     // index among the epilogue warps
        const int quad = warp & 3;                                   // TMEM lanes 32 quad .. +31 are the ones this warp may read
        const bool second = ew >= 4;
        float* pad = s_pad[ew];
        for (uint32_t tcount = 0; tile_of(tcount << gshift) < total_tiles; ++tcount) {     // one row group per iteration
            const int tile = tile_of(tcount << gshift), next_tile = tile_of((tcount + 1) << gshift);
            const int tm = tile / P.tiles_n, tn = tile % P.tiles_n;
            const int acc = tcount & 1;
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            tc_fence_after();
            // fused LayerNorm: split between the two warps of a quadrant when the row has an even number of 32-column blocks
            const bool ln_split = LN_SPLIT && EW == 8 && P.ln && ((P.N >> 5) & 1) == 0;
            const bool halves = EW == 8 && (!P.ln || ln_split);
            if (!(P.ln && second && !ln_split))
                tile_epilogue<BN, LNW, MINB == 1>(P, pad, s_ln, quad, lane, tm, tn, next_tile < total_tiles ? next_tile / P.tiles_n : -1,
                                       tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * BN) << gshift), (halves && second) ? 1 : 0, halves ? 2 : 1,
                                  ln_split ? s_stat[EW == 8 ? quad : 0][tcount & 1] : nullptr);
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int BN, int OPS, int RAW, int LW, int MINB, bool GATHER = false>
int launch_tc3(const Tc2Params& P, cudaStream_t st) {
    constexpr int T3_THREADS = 128 + LW * 32 + 64 + (MINB == 1 ? 128 : 0);
    constexpr int smem = OPS * (2 * A_HALF + 2 * BN * 128) + RAW * RAW_BYTES + 1024;
    static bool attr_dev[ROITR_MAX_DEVICES] = {};
    static int num_sms_dev[ROITR_MAX_DEVICES] = {}, per_sm_dev[ROITR_MAX_DEVICES] = {};
    const int dv = roitr_cur_device();
    bool& attr = attr_dev[dv];
    int &num_sms = num_sms_dev[dv], &per_sm = per_sm_dev[dv];
    if (!attr) {
        ROITR_CUDA(cudaFuncSetAttribute(linear_tc3_kernel<BN, OPS, RAW, LW, MINB, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0;
        ROITR_CUDA(cudaGetDevice(&dev));
        ROITR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        ROITR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, linear_tc3_kernel<BN, OPS, RAW, LW, MINB, GATHER>, T3_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        attr = true;
    }
    const int total = (P.ln && P.tiles_n == 2) ? P.tiles_m : P.tiles_m * P.tiles_n;      // row groups (see the kernel)
    const int grid = total < num_sms * per_sm ? total : num_sms * per_sm;
    linear_tc3_kernel<BN, OPS, RAW, LW, MINB, GATHER><<<grid, T3_THREADS, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("linear_tc3_kernel");
    return ROITR_OK;
}

// ---- row-group variant ---------------------------------------------------------------------------------------------------
// For N > 128 the streaming kernel above visits a row tile once per 128-column weight tile, so the activations are fetched
// (from L2 the second time) and split into TF32 hi / lo once per weight tile - and the loader warps are what limits that
// kernel. Here the unit of work is (row tile, PAIR of weight tiles): the operand rings are separate (SA activation stages,
// SB weight stages), the k loop is outermost, and each split activation chunk feeds the MMAs of both weight tiles into the
// two halves of a 256-column accumulator slot (2 slots = the whole TMEM). Plain epilogue only (bias / ReLU, fragment-layout
// stores); the LayerNorm epilogue, gathered rows and the light configuration stay with the kernel above.
constexpr int T4_LW = 8, T4_THREADS = 128 + T4_LW * 32 + 64 + 128;
template <int SA, int SB, int RAW>
__global__ void __launch_bounds__(T4_THREADS, 1) linear_tc4_kernel(const Tc2Params P) {
    constexpr int BN = 128, EW = 8, T4_LOADERS = T4_LW * 32;
    constexpr int B_HALF = BN * 128;
    constexpr int A_STAGE = 2 * A_HALF, B_STAGE = 2 * B_HALF;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                              // SA x [hi | lo], 128 rows x 128 B each
    unsigned char* sB = smem + SA * A_STAGE;               // SB x [hi | lo]
    unsigned char* raw = sB + SB * B_STAGE;                // RAW x 16 KB of raw fp32 activations
    __shared__ __align__(8) uint64_t a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], T4_LOADERS); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EW * 32); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int ngroups = (P.tiles_n + 1) >> 1;
    const int units = P.tiles_m * ngroups;                 // unit u: row tile u / ngroups, weight tiles 2 g, 2 g + 1 (g = u % ngroups)
    const int nkc = P.nkc;
    constexpr int W_LOADER0 = 4, W_PRODUCER = 4 + T4_LW, W_MMA = W_PRODUCER + 1;

    if (warp >= W_LOADER0 && warp < W_PRODUCER) {
        // ================================================= A loaders =================================================
        const int t = tid - 128;
        constexpr int RSTEP = T4_LW * 4, NPIECE = T2_BM / RSTEP;
        const int c = t & 7, r0 = t >> 3;
        const uint32_t raw_u32 = smem_u32(raw);
        int i_unit = blockIdx.x, i_kc = 0;
        uint32_t i_it = 0;
        auto issue_one = [&]() {
            if (i_unit < units) {
                const uint32_t dst = raw_u32 + (i_it % RAW) * RAW_BYTES + (uint32_t)(r0 * 128 + c * 16);
                const int k = i_kc * T2_BK + 4 * c;
                const int am0 = (i_unit / ngroups) * T2_BM + r0;
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) {
                    const int am = am0 + RSTEP * j;
                    const bool ok = am < P.M && k < P.K;
                    cp_async16(dst + j * RSTEP * 128, P.A + (ok ? (long long)am * P.lda + k : 0ll), ok ? 16u : 0u);
                }
                ++i_it;
                if (++i_kc == nkc) { i_kc = 0; i_unit += gridDim.x; }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int j = 0; j < RAW - 1; ++j) issue_one();
        uint32_t it = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                issue_one();
                cp_async_wait<RAW - 1>();
                const unsigned char* src = raw + (it % RAW) * RAW_BYTES + r0 * 128 + c * 16;
                float4 v[NPIECE];
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) v[j] = *reinterpret_cast<const float4*>(src + j * RSTEP * 128);
                const int sa = it % SA;
                mbar_wait(&a_empty[sa], ((it / SA) & 1) ^ 1);
                unsigned char* a_hi = sA + sa * A_STAGE;
#pragma unroll
                for (int j = 0; j < NPIECE; ++j) split_store(a_hi, a_hi + A_HALF, r0 + RSTEP * j, c, v[j]);
                fence_proxy_async_smem();
                mbar_arrive(&a_full[sa]);
            }
        }
        cp_async_wait<0>();
    } else if (warp == W_PRODUCER) {
        // ================================================= W producer ================================================
        if (lane == 0) {
            uint32_t ib = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int g = unit % ngroups;
                const int nt = min(2, P.tiles_n - 2 * g);
                for (int kc = 0; kc < nkc; ++kc) {
                    for (int t = 0; t < nt; ++t, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(&b_empty[sb], ((ib / SB) & 1) ^ 1);
                        mbar_expect_tx(&b_full[sb], 2 * B_HALF);
                        tma_load_1d(sB + sb * B_STAGE, P.wpack + ((size_t)(2 * g + t) * nkc + kc) * (2 * B_HALF / 4), 2 * B_HALF, &b_full[sb]);
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================================= MMA issuer ================================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(T2_BM, BN);
            uint32_t ia = 0, ib = 0, ucount = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++ucount) {
                const int g = unit % ngroups;
                const int nt = min(2, P.tiles_n - 2 * g);
                const int slot = ucount & 1;
                mbar_wait(&tmem_empty[slot], ((ucount >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < nkc; ++kc, ++ia) {
                    const int sa = ia % SA;
                    mbar_wait(&a_full[sa], (ia / SA) & 1);
                    tc_fence_after();
                    const uint32_t ah = smem_u32(sA + sa * A_STAGE), al = ah + A_HALF;
                    for (int t = 0; t < nt; ++t, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(&b_full[sb], (ib / SB) & 1);
                        tc_fence_after();
                        const uint32_t bh = smem_u32(sB + sb * B_STAGE), bl = bh + B_HALF;
                        const uint32_t d = tmem + (uint32_t)(slot * 256 + t * BN);
#pragma unroll
                        for (int ks = 0; ks < T2_BK / 8; ++ks) {
                            const uint64_t dah = make_desc_sw128(ah + ks * 32), dal = make_desc_sw128(al + ks * 32);
                            const uint64_t dbh = make_desc_sw128(bh + ks * 32), dbl = make_desc_sw128(bl + ks * 32);
                            umma_tf32(d, dal, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                            umma_tf32(d, dah, dbl, idesc, 1u);
                            umma_tf32(d, dah, dbh, idesc, 1u);
                        }
                        umma_commit(&b_empty[sb]);                             // weight stage reusable when these MMAs retire
                    }
                    umma_commit(&a_empty[sa]);                                 // activation stage: after BOTH weight tiles
                }
                umma_commit(&tmem_full[slot]);
            }
        }
    } else {
        // ================================================= epilogue (warps 0-3 and the last four) =====================
        const int ew = warp < 4 ? warp : 4 + (warp - W_MMA - 1);
        const int quad = warp & 3;
        const int first = ew >= 4 ? 32 : 0;                  // the two warps of a quadrant take even / odd 32-column blocks
        uint32_t ucount = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++ucount) {
            const int tm = unit / ngroups, g = unit % ngroups;
            const int nt = min(2, P.tiles_n - 2 * g);
            const int slot = ucount & 1;
            mbar_wait(&tmem_full[slot], (ucount >> 1) & 1);
            tc_fence_after();
            const int row0 = tm * T2_BM + quad * 32, ncol0 = g * 2 * BN;
            const uint32_t tq = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * 256);
#pragma unroll 1
            for (int c0 = first; c0 < nt * BN; c0 += 64) {
                if (ncol0 + c0 >= P.N) break;
                store_block32_frag(P, tq + c0, lane, row0, ncol0 + c0);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int SA, int SB, int RAW>
int launch_tc4(const Tc2Params& P, cudaStream_t st) {
    constexpr int smem = SA * 2 * A_HALF + SB * 2 * 128 * 128 + RAW * RAW_BYTES + 1024;
    static bool attr_dev[ROITR_MAX_DEVICES] = {};
    static int num_sms_dev[ROITR_MAX_DEVICES] = {};
    const int dv = roitr_cur_device();
    bool& attr = attr_dev[dv];
    int& num_sms = num_sms_dev[dv];
    if (!attr) {
        ROITR_CUDA(cudaFuncSetAttribute(linear_tc4_kernel<SA, SB, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0;
        ROITR_CUDA(cudaGetDevice(&dev));
        ROITR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr = true;
    }
    const int units = P.tiles_m * ((P.tiles_n + 1) / 2);
    const int grid = units < num_sms ? units : num_sms;
    linear_tc4_kernel<SA, SB, RAW><<<grid, T4_THREADS, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("linear_tc4_kernel");
    return ROITR_OK;
}

template <int BN, int STAGES>
int launch_tc2(const Tc2Params& P, cudaStream_t st) {
    constexpr int smem = STAGES * (2 * A_HALF + 2 * BN * 128) + 1024;   // + ~21 KB static (transpose pads, bias)
    static bool attr_dev[ROITR_MAX_DEVICES] = {};
    static int num_sms_dev[ROITR_MAX_DEVICES] = {};
    const int dv = roitr_cur_device();
    bool& attr = attr_dev[dv];
    int& num_sms = num_sms_dev[dv];
    if (!attr) {
        ROITR_CUDA(cudaFuncSetAttribute(linear_tc2_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int dev = 0;
        ROITR_CUDA(cudaGetDevice(&dev));
        ROITR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr = true;
    }
    const int total = P.tiles_m * P.tiles_n;
    const int grid = total < num_sms ? total : num_sms;
    linear_tc2_kernel<BN, STAGES><<<grid, T2_THREADS, smem, st>>>(P);
    ROITR_CHECK_LAUNCH("linear_tc2_kernel");
    return ROITR_OK;
}

}  // namespace

// Configuration of the streaming kernel for the launches issued from now on (a host-side mode the engine sets around the
// level-1 layers it issues while the FPS clusters hold most SMs): 0 = deep rings, one CTA per SM (215 KB, 57 K registers);
// 3 = light footprint (one operand stage, two raw stages, 4 loader warps, <= 64 registers: ~115 KB, 20 K registers) that
// shares an SM with the CTAs of other streams' kernels. Same arithmetic, same results.
#ifndef TC4_ON
#define TC4_ON 1
#endif
static int g_tc3_variant = 0;
extern "C" int roitr_set_linear_config(int v) { g_tc3_variant = (v == 3) ? 3 : 0; return 0; }

static int linear_tc_packed_impl(int M, int N, int K, const float* A, const float* a_add, int lda, const int* a_index,
                                 const float* wpack, int bn, const float* bias, float* C, int ldc, int relu,
                                 const float* ln_gamma, const float* ln_beta, const float* res_pre, const int* res_pre_index,
                                 const float* res_post, int ldr, void* stream) {
    ROITR_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && A && wpack && C, "linear_tc_packed: bad arguments M=%d N=%d K=%d", M, N, K);
    ROITR_CHECK_ARG(lda >= K && ldc >= N, "linear_tc_packed: bad leading dimensions");
    ROITR_CHECK_ARG(bn == 64 || bn == 128, "linear_tc_packed: weights must be packed with 64- or 128-row tiles, got %d", bn);
    ROITR_CHECK_ARG((uintptr_t)wpack % 16 == 0, "linear_tc_packed: wpack must be 16-byte aligned");
    if (M == 0) return ROITR_OK;
    Tc2Params P;
    P.M = M; P.N = N; P.K = K; P.A = A; P.A2 = a_add; P.lda = lda; P.a_index = a_index; P.wpack = wpack; P.bias = bias; P.C = C;
    P.ldc = ldc; P.relu = relu;
    P.ln = ln_gamma != nullptr; P.ln_gamma = ln_gamma; P.ln_beta = ln_beta; P.res_pre = res_pre; P.res_pre_index = res_pre_index;
    P.res_post = res_post; P.ldr = ldr;
    P.tiles_m = ceil_div(M, T2_BM); P.tiles_n = ceil_div(N, bn); P.nkc = ceil_div(K, T2_BK);
    cudaStream_t st = (cudaStream_t)stream;
    // the streaming kernel takes plain or GATHERED rows (a_index: 16-byte cp.async pieces of the indexed row); the sum of two
    // inputs (a_add) needs the register-staged loaders of the coupled-ring kernel
    const bool stream_ok = !a_add && lda % 4 == 0 && K % 4 == 0 && (uintptr_t)A % 16 == 0;
    if (P.ln)
        ROITR_CHECK_ARG(stream_ok && (P.tiles_n == 1 || (P.tiles_n == 2 && bn == 128)) && N % 32 == 0 && ln_beta && ldr >= N && ldr % 4 == 0 &&
                            ((uintptr_t)res_pre | (uintptr_t)res_post) % 16 == 0,
                        "linear_ln_tc_packed: needs a plain 16-byte aligned input, N a multiple of 32 within one weight tile, or two "
                        "128-row tiles (N=%d, tile %d)", N, bn);
    if (P.ln) {
        if (P.tiles_n == 2) return launch_tc3<128, 2, 3, 8, 1>(P, st);          // the 512-column layout exists in the deep configuration only
        if (g_tc3_variant == 3) return bn == 64 ? launch_tc3<64, 1, 2, 4, 3>(P, st) : launch_tc3<128, 1, 2, 4, 3>(P, st);
        return bn == 64 ? launch_tc3<64, 2, 4, 8, 1>(P, st) : launch_tc3<128, 2, 3, 8, 1>(P, st);
    }
    if (stream_ok && a_index) return bn == 128 ? launch_tc3<128, 2, 3, 8, 1, true>(P, st) : launch_tc3<64, 2, 4, 8, 1, true>(P, st);
    if (stream_ok) {
        if (g_tc3_variant == 3) return bn == 128 ? launch_tc3<128, 1, 2, 4, 3>(P, st) : launch_tc3<64, 1, 2, 4, 3>(P, st);
        // more than one 128-column weight tile: activations loaded and split once per PAIR of weight tiles
        // (measured, profiles/r02_bench_gemm.txt: 0.213 -> 0.148 ms at (640000, 192, 64), 0.201 -> 0.168 at N = 256; with fewer than
        // ~74 row tiles the one-tile-per-CTA kernel spreads over more SMs and wins: 0.015 vs 0.023 ms at (4992, 256, 256); three
        // weight tiles - one pair and a single - are no better than the kernel above)
        if (TC4_ON && bn == 128 && P.tiles_n >= 2 && P.tiles_n != 3 && P.tiles_m >= 74 && ldc % 2 == 0 && N % 2 == 0 &&
            (uintptr_t)C % 8 == 0 && (uintptr_t)bias % 8 == 0)
            // rings: 2 activation stages, 2 weight stages, 4 raw chunks in flight (2/3/3, 2/4/2, 3/3/2 measured within 5 % of it)
            return launch_tc4<2, 2, 4>(P, st);
        return bn == 128 ? launch_tc3<128, 2, 3, 8, 1>(P, st) : launch_tc3<64, 2, 4, 8, 1>(P, st);
    }
    if (bn == 64) return launch_tc2<64, 4>(P, st);
    return launch_tc2<128, 3>(P, st);
}

extern "C" int roitr_linear_tc_packed(int M, int N, int K, const float* A, const float* a_add, int lda, const int* a_index,
                                      const float* wpack, int bn, const float* bias, float* C, int ldc, int relu,
                                      void* stream) {
    return linear_tc_packed_impl(M, N, K, A, a_add, lda, a_index, wpack, bn, bias, C, ldc, relu, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, 0, stream);
}

extern "C" int roitr_linear_ln_tc_packed(int M, int N, int K, const float* A, int lda, const float* wpack, int bn,
                                         const float* bias, const float* gamma, const float* beta, const float* res_pre,
                                         const int* res_pre_index, const float* res_post, int ldr, int relu, float* C,
                                         int ldc, void* stream) {
    ROITR_CHECK_ARG(gamma && beta, "linear_ln_tc_packed: gamma / beta required");
    return linear_tc_packed_impl(M, N, K, A, nullptr, lda, nullptr, wpack, bn, bias, C, ldc, relu, gamma, beta, res_pre,
                                 res_pre_index, res_post, ldr, stream);
}
