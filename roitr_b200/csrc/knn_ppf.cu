// Exact kNN fused with point-pair-feature construction (sm_100a).
//
// Replaces knnquery_cuda_kernel (cpp_wrappers/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-108) and the ~17
// eager ops of queryandgroup(return_idx) + gathers + calc_ppf_gpu (pointops.py:87-92, model/model.py:36-41,
// lib/utils.py:358-389) with one kernel.
//
// Design (B200): the reference runs one thread per query with a 100-slot heap in local memory (800 B stack) and a
// full scan from global memory; 79 CTAs for 20k queries. Here a WARP owns QW queries. Reference points are streamed
// through shared memory in 2048-point tiles by the TMA engine (1-D cp.async.bulk + mbarrier, double buffered; xyz
// rows are 12 B so tiles are flat byte ranges starting at multiples of 4 points = 48 B, 16-B aligned). Each lane
// evaluates one reference point per step for all QW queries (stride-3-word LDS is bank-conflict free). The running
// top-k of a query is a sorted list DISTRIBUTED OVER THE LANES (lane l holds the l-th best), so admission is one
// compare + ballot per 32 points and an insertion is a ballot/popc + one shuffle — no local memory, no divergence.
// Points are consumed in ascending index order and admission is strict '<', so among exactly equal distances the
// lower index wins; queries in which an exact-distance tie played any role (rare) are marked and replayed by
// knn_tie_fixup_kernel with the reference's sequential max-heap, so results equal the reference bit for bit, ties included.
// The epilogue drops the self column and computes the PPF tuple per (query, neighbour) lane, writing coalesced
// idx and float4 PPF.
#include <math_constants.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"
#include "knn_grid.cuh"

namespace {

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int TILE_PTS = 2048;
constexpr int TILE_BYTES = TILE_PTS * 12;  // 24 KB, multiple of 16

struct KnnParams {
    const float* xyz;
    const float* nrm;
    const float* qxyz;
    const float* qnrm;
    const int* offset;
    const int* new_offset;
    int b, m, nslots, drop;
    int* idx;
    float* dist;
    float* ppf;
    int dist_squared;
    int n_total;
    int use_tma;
};

__device__ __forceinline__ int find_segment(int q, const int* __restrict__ ends, int b) {
    int lo = 0, hi = b - 1;  // first s with q < ends[s]
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (q < __ldg(ends + mid)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// angle(a, b) / pi  =  atan2(|a x b|, a.b) / pi   (lib/utils.py:372-387). Products and sums are rounded separately
// (no FMA contraction) like the eager elementwise reference.
__device__ __forceinline__ float angle_over_pi(float ax, float ay, float az, float bx, float by, float bz) {
    // torch.sum starts from +0, so an all-(-0) dot product is +0 and atan2(0, +0) = 0 (not pi)
    float dot = __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(ax, bx)), __fmul_rn(ay, by)), __fmul_rn(az, bz));
    float cx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    float cy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    float cz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
    float cn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    return __fdiv_rn(atan2f(cn, dot), 3.14159265358979323846f);
}

// writes one (query, slot) result: index, distance, PPF tuple
__device__ __forceinline__ void emit_result(const KnnParams& P, int q, int slot, int kout, int nb, float d2, float qx,
                                            float qy, float qz) {
    const size_t o = (size_t)q * kout + slot;
    P.idx[o] = nb;
    if (P.dist) P.dist[o] = P.dist_squared ? d2 : __fsqrt_rn(d2);
    if (P.ppf) {
        const float n1x = __ldg(P.qnrm + 3 * (size_t)q), n1y = __ldg(P.qnrm + 3 * (size_t)q + 1),
                    n1z = __ldg(P.qnrm + 3 * (size_t)q + 2);
        const float px = __ldg(P.xyz + 3 * (size_t)nb), py = __ldg(P.xyz + 3 * (size_t)nb + 1),
                    pz = __ldg(P.xyz + 3 * (size_t)nb + 2);
        const float n2x = __ldg(P.nrm + 3 * (size_t)nb), n2y = __ldg(P.nrm + 3 * (size_t)nb + 1),
                    n2z = __ldg(P.nrm + 3 * (size_t)nb + 2);
        const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
        float4 f;
        f.x = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        f.y = angle_over_pi(n1x, n1y, n1z, dx, dy, dz);
        f.z = angle_over_pi(n2x, n2y, n2z, dx, dy, dz);
        f.w = angle_over_pi(n1x, n1y, n1z, n2x, n2y, n2z);
        reinterpret_cast<float4*>(P.ppf)[o] = f;
    }
}

template <int QW>
__global__ void __launch_bounds__(KNN_THREADS, 3) knn_ppf_kernel(const KnnParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile0 = reinterpret_cast<float*>(smem_raw);
    float* tile1 = reinterpret_cast<float*>(smem_raw + TILE_BYTES);
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ int s_cta[3];  // lo, hi, mixed

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = P.nslots;
    const unsigned kmask = (K >= 32) ? FULL_MASK : ((1u << K) - 1u);
    const int q_cta0 = blockIdx.x * (KNN_WARPS * QW);

    if (tid == 0) {
        int q_last = min(q_cta0 + KNN_WARPS * QW, P.m) - 1;
        int s0 = find_segment(q_cta0, P.new_offset, P.b);
        int s1 = find_segment(q_last, P.new_offset, P.b);
        s_cta[0] = s0 == 0 ? 0 : __ldg(P.offset + s0 - 1);
        s_cta[1] = __ldg(P.offset + s1);
        s_cta[2] = (s0 != s1);
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }

    // ---- per-query state (replicated across the warp except the distributed list) ----
    float qx[QW], qy[QW], qz[QW], ld[QW], tau[QW];
    int qs[QW], qe[QW], li[QW];
    bool tie[QW];  // an exact-distance tie influenced this query: its result is recomputed by knn_tie_fixup_kernel
#pragma unroll
    for (int j = 0; j < QW; ++j) {
        tie[j] = false;
        int q = q_cta0 + warp * QW + j;
        if (q < P.m) {
            int s = find_segment(q, P.new_offset, P.b);
            qs[j] = s == 0 ? 0 : __ldg(P.offset + s - 1);
            qe[j] = __ldg(P.offset + s);
            qx[j] = __ldg(P.qxyz + 3 * (size_t)q);
            qy[j] = __ldg(P.qxyz + 3 * (size_t)q + 1);
            qz[j] = __ldg(P.qxyz + 3 * (size_t)q + 2);
            ld[j] = 1e10f;  // knnquery_cuda_kernel.cu:88-91
            li[j] = qs[j];
            tau[j] = 1e10f;
        } else {
            qs[j] = qe[j] = 0;
            qx[j] = qy[j] = qz[j] = 0.f;
            ld[j] = -1.f;
            li[j] = 0;
            tau[j] = -1.f;  // nothing is ever admitted
        }
    }
    __syncthreads();
    const int lo = s_cta[0], hi = s_cta[1];
    const bool mixed = s_cta[2] != 0;
    const int base0 = lo & ~3;
    const int ntiles = (hi > base0) ? (hi - base0 + TILE_PTS - 1) / TILE_PTS : 0;

    auto tile_is_tma = [&](int t) { return P.use_tma && (base0 + (t + 1) * TILE_PTS <= P.n_total); };
    auto issue_tma = [&](int t) {
        int buf = t & 1;
        mbar_expect_tx(&full_bar[buf], TILE_BYTES);
        tma_load_1d(buf ? tile1 : tile0, P.xyz + 3 * (size_t)(base0 + t * TILE_PTS), TILE_BYTES, &full_bar[buf]);
    };

    if (tid == 0 && ntiles > 0 && tile_is_tma(0)) issue_tma(0);

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const float* tile = buf ? tile1 : tile0;
        const int base = base0 + t * TILE_PTS;
        if (tile_is_tma(t)) {
            // prefetch the next tile into the other buffer (its readers finished at the barrier ending iteration t-1)
            if (tid == 0 && t + 1 < ntiles && tile_is_tma(t + 1)) issue_tma(t + 1);
            mbar_wait(&full_bar[buf], (t >> 1) & 1);
        } else {
            if (tid == 0 && t + 1 < ntiles && tile_is_tma(t + 1)) issue_tma(t + 1);
            const int npts = min(TILE_PTS, P.n_total - base);
            float* wt = buf ? tile1 : tile0;
            const float* src = P.xyz + 3 * (size_t)base;
            for (int i = tid; i < npts * 3; i += KNN_THREADS) wt[i] = __ldg(src + i);
            __syncthreads();
        }

        const int t_lo = max(base, lo) - base;
        const int t_hi = min(base + TILE_PTS, hi) - base;
        for (int s = t_lo & ~31; s < t_hi; s += 32) {
            const int pos = s + lane;
            const bool inb = (pos >= t_lo) && (pos < t_hi);
            const float x = tile[3 * pos], y = tile[3 * pos + 1], z = tile[3 * pos + 2];
            const int gi = base + pos;
#pragma unroll
            for (int j = 0; j < QW; ++j) {
                const float d = sqdist_ref(qx[j] - x, qy[j] - y, qz[j] - z);
                bool ok = inb;
                if (mixed) ok = ok && (gi >= qs[j]) && (gi < qe[j]);
                unsigned mask = __ballot_sync(FULL_MASK, ok && (d <= tau[j]));
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float cd = __shfl_sync(FULL_MASK, d, l);
                    if (cd < tau[j]) {
                        const int ci = base + s + l;
                        const int ins = __popc(__ballot_sync(FULL_MASK, ld[j] <= cd) & kmask);
                        const float up_d = __shfl_up_sync(FULL_MASK, ld[j], 1);
                        const int up_i = __shfl_up_sync(FULL_MASK, li[j], 1);
                        if (lane > ins) { ld[j] = up_d; li[j] = up_i; }
                        else if (lane == ins) { ld[j] = cd; li[j] = ci; }
                        const float tau_old = tau[j];
                        tau[j] = __shfl_sync(FULL_MASK, ld[j], K - 1);
                        // the evicted element equals the new k-th one: which of the two the reference keeps depends on
                        // its heap history (1e10 = the unfilled-slot filler, all identical, no ambiguity)
                        if (tau[j] == tau_old && tau_old != 1e10f) tie[j] = true;
                    } else if (cd == tau[j] && cd != 1e10f) {
                        tie[j] = true;  // rejected at the boundary by strict '<' here; the heap may have kept another one
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with `buf` before it is refilled
    }

    // ---- epilogue: drop leading columns, PPF per (query, neighbour) lane ----
    const int kout = K - P.drop;
#pragma unroll
    for (int j = 0; j < QW; ++j) {
        const int q = q_cta0 + warp * QW + j;
        if (q >= P.m) continue;
        // equal distances inside the final list: the reference's order among them is its heap's, not ours
        {
            const float nxt = __shfl_down_sync(FULL_MASK, ld[j], 1);
            if (__ballot_sync(FULL_MASK, lane + 1 < K && ld[j] == nxt && ld[j] != 1e10f)) tie[j] = true;
        }
        const int slot = lane - P.drop;
        if (slot < 0 || lane >= K) continue;
        if (tie[j]) {
            if (slot == 0) P.idx[(size_t)q * kout] = -1;  // marker consumed by knn_tie_fixup_kernel
            continue;
        }
        emit_result(P, q, slot, kout, li[j], ld[j], qx[j], qy[j], qz[j]);
    }
}

// Exact replay of the reference algorithm (sequential scan in index order, max-heap with strict '<', heap sort:
// knnquery_cuda_kernel.cu:21-48,86-107) for the queries the main kernel marked as tie-affected (about one or two per
// 20k x 20k call on random data). One WARP per marked query: the 32 lanes evaluate 32 consecutive points and filter
// them against the current heap root (the root only decreases, so the filter is conservative); the few survivors are
// pushed through the heap one by one, in index order, by lane 0 - exactly the reference's sequence of heap updates.
constexpr int FIX_WARPS = 4;
__global__ void __launch_bounds__(FIX_WARPS * 32) knn_tie_fixup_kernel(const KnnParams P) {
    __shared__ float s_bd[FIX_WARPS][32];
    __shared__ int s_bi[FIX_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = P.nslots, kout = K - P.drop;
    const int q_first = (blockIdx.x * FIX_WARPS + warp) * 32;
    const int my_q = q_first + lane;
    unsigned todo = __ballot_sync(FULL_MASK, my_q < P.m && P.idx[(size_t)my_q * kout] < 0);
    float* bd = s_bd[warp];
    int* bi = s_bi[warp];
    auto sift = [&](int k) {  // lane 0 only
        int root = 0, child = 1;
        while (child < k) {
            if (child + 1 < k && bd[child + 1] > bd[child]) child++;
            if (bd[root] > bd[child]) return;
            const float td = bd[root]; bd[root] = bd[child]; bd[child] = td;
            const int ti = bi[root]; bi[root] = bi[child]; bi[child] = ti;
            root = child;
            child = 2 * root + 1;
        }
    };
    while (todo) {
        const int q = q_first + __ffs(todo) - 1;
        todo &= todo - 1;
        const int sgm = find_segment(q, P.new_offset, P.b);
        const int start = sgm == 0 ? 0 : __ldg(P.offset + sgm - 1), end = __ldg(P.offset + sgm);
        const float qx = __ldg(P.qxyz + 3 * (size_t)q), qy = __ldg(P.qxyz + 3 * (size_t)q + 1), qz = __ldg(P.qxyz + 3 * (size_t)q + 2);
        if (lane < K) { bd[lane] = 1e10f; bi[lane] = start; }
        __syncwarp();
        constexpr int U = 8;  // 8 x 32 points per outer step: the independent loads overlap their L2 latency
        for (int base = start; base < end; base += 32 * U) {
            float d2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + 32 * u + lane;
                d2[u] = CUDART_INF_F;
                if (i < end)
                    d2[u] = sqdist_ref(qx - __ldg(P.xyz + 3 * (size_t)i), qy - __ldg(P.xyz + 3 * (size_t)i + 1),
                                       qz - __ldg(P.xyz + 3 * (size_t)i + 2));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {  // still strictly in index order
                unsigned cand = __ballot_sync(FULL_MASK, d2[u] < bd[0]);
                while (cand) {
                    const int l = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const float cd = __shfl_sync(FULL_MASK, d2[u], l);
                    if (lane == 0 && cd < bd[0]) { bd[0] = cd; bi[0] = base + 32 * u + l; sift(K); }
                    __syncwarp();
                }
            }
        }
        if (lane == 0)
            for (int i = K - 1; i > 0; --i) {
                const float td = bd[0]; bd[0] = bd[i]; bd[i] = td;
                const int ti = bi[0]; bi[0] = bi[i]; bi[i] = ti;
                sift(i);
            }
        __syncwarp();
        if (lane >= P.drop && lane < K) emit_result(P, q, lane - P.drop, kout, bi[lane], bd[lane], qx, qy, qz);
        __syncwarp();
    }
}

// ---- grid-accelerated exact kNN: one warp per query, candidates from the cells of a growing cube -------------------------
// Same distance arithmetic, same (distance, index) order, same tie marking / replay as the brute-force kernel; candidates
// arrive in cell order instead of index order, so admission and insertion compare (distance, index) lexicographically.
__global__ void __launch_bounds__(KNN_THREADS, 4) knn_grid_kernel(const KnnParams P, const knngrid::SegHeader* __restrict__ hdr,
                                                               const int* __restrict__ cell_start,
                                                               const float4* __restrict__ sorted) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * KNN_WARPS + warp;
    if (q >= P.m) return;
    const int K = P.nslots, kout = K - P.drop;
    const unsigned kmask = (K >= 32) ? FULL_MASK : ((1u << K) - 1u);
    const int sgm = find_segment(q, P.new_offset, P.b);
    const int qs = sgm == 0 ? 0 : __ldg(P.offset + sgm - 1);
    const knngrid::SegHeader H = hdr[sgm];
    const float qx = __ldg(P.qxyz + 3 * (size_t)q), qy = __ldg(P.qxyz + 3 * (size_t)q + 1), qz = __ldg(P.qxyz + 3 * (size_t)q + 2);
    const int cx = knngrid::cell_coord(qx, H.ox, H.inv_h, H.nx), cy = knngrid::cell_coord(qy, H.oy, H.inv_h, H.ny),
              cz = knngrid::cell_coord(qz, H.oz, H.inv_h, H.nz);
    const int* cs = cell_start + H.cell_base;
    const float4* pts = sorted + qs;

    float ld = 1e10f, tau = 1e10f;
    int li = qs, tau_i = qs;
    bool tie = false;

    auto scan_range = [&](int beg, int end) {
        for (int p0 = beg; p0 < end; p0 += 32) {
            const int p = p0 + lane;
            float d = CUDART_INF_F;
            int ci = 0;
            if (p < end) {
                const float4 v = __ldg(pts + p);
                d = sqdist_ref(qx - v.x, qy - v.y, qz - v.z);
                ci = __float_as_int(v.w);
            }
            unsigned mask = __ballot_sync(FULL_MASK, d <= tau);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const float cd = __shfl_sync(FULL_MASK, d, l);
                const int cidx = __shfl_sync(FULL_MASK, ci, l);
                if (cd == tau && tau != 1e10f) tie = true;     // boundary tie: the reference's choice depends on its heap
                if (cd < tau || (cd == tau && cidx < tau_i)) {
                    const bool before = (ld < cd) || (ld == cd && li < cidx);
                    const int ins = __popc(__ballot_sync(FULL_MASK, before) & kmask);
                    const float up_d = __shfl_up_sync(FULL_MASK, ld, 1);
                    const int up_i = __shfl_up_sync(FULL_MASK, li, 1);
                    if (lane > ins) { ld = up_d; li = up_i; }
                    else if (lane == ins) { ld = cd; li = cidx; }
                    const float tau_old = tau;
                    tau = __shfl_sync(FULL_MASK, ld, K - 1);
                    tau_i = __shfl_sync(FULL_MASK, li, K - 1);
                    if (tau == tau_old && tau_old != 1e10f) tie = true;
                }
            }
        }
    };

    for (int r = 0;; ++r) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, H.nx - 1);
        for (int dz = -r; dz <= r; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= H.nz) continue;
            for (int dy = -r; dy <= r; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= H.ny) continue;
                const int row = (z * H.ny + y) * H.nx;
                if (abs(dz) == r || abs(dy) == r) {               // face of the cube: the whole x span
                    scan_range(__ldg(cs + row + x0), __ldg(cs + row + x1 + 1));
                } else {                                         // interior row: only the two end cells
                    if (cx - r >= 0) scan_range(__ldg(cs + row + cx - r), __ldg(cs + row + cx - r + 1));
                    if (cx + r < H.nx) scan_range(__ldg(cs + row + cx + r), __ldg(cs + row + cx + r + 1));
                }
            }
        }
        // everything outside the cube of radius r is at least `bound` away from the query
        const bool all = (cx - r <= 0) && (cx + r >= H.nx - 1) && (cy - r <= 0) && (cy + r >= H.ny - 1) && (cz - r <= 0) &&
                         (cz + r >= H.nz - 1);
        if (all) break;
        float bound = CUDART_INF_F;
        if (cx - r > 0) bound = fminf(bound, qx - (H.ox + (float)(cx - r) * H.h));
        if (cx + r < H.nx - 1) bound = fminf(bound, (H.ox + (float)(cx + r + 1) * H.h) - qx);
        if (cy - r > 0) bound = fminf(bound, qy - (H.oy + (float)(cy - r) * H.h));
        if (cy + r < H.ny - 1) bound = fminf(bound, (H.oy + (float)(cy + r + 1) * H.h) - qy);
        if (cz - r > 0) bound = fminf(bound, qz - (H.oz + (float)(cz - r) * H.h));
        if (cz + r < H.nz - 1) bound = fminf(bound, (H.oz + (float)(cz + r + 1) * H.h) - qz);
        bound -= 1e-4f * H.h;                                     // binning rounds (v - o) * inv_h: keep a safety margin
        if (tau < 1e10f && bound > 0.f && tau < bound * bound * 0.99999f) break;
    }
    {   // equal distances inside the final list: the reference's order among them is its heap's
        const float nxt = __shfl_down_sync(FULL_MASK, ld, 1);
        if (__ballot_sync(FULL_MASK, lane + 1 < K && ld == nxt && ld != 1e10f)) tie = true;
    }
    const int slot = lane - P.drop;
    if (slot < 0 || lane >= K) return;
    if (tie) { if (slot == 0) P.idx[(size_t)q * kout] = -1; return; }
    emit_result(P, q, slot, kout, li, ld, qx, qy, qz);
}

// ---- grid-accelerated exact kNN, one THREAD per query ----------------------------------------------------------------------
// The warp-per-query kernel above spends its time in dependent loads: every row of every shell is a cell_start lookup
// followed by a point fetch, ~45 serial round trips per query with 32 lanes sharing ~6 candidates. Here a thread owns a
// query and keeps its k best (distance, index) pairs as a sorted list in REGISTERS (fully unrolled compare-exchange
// insertion, K is a template parameter); memory latency is hidden by the other ~1000 resident queries of the SM instead
// of by lanes. Queries are visited in the cell order of the QUERY set's own grid when one is given (for self queries
// the same grid): the 32 queries of a warp then sit in the same or adjacent cells, walk nearly identical candidate
// ranges (L1 hits, little divergence). Same distance arithmetic, same lexicographic (distance, index) order, same tie
// marking / replay as the other kernels, so the result is bit-identical. Results are staged in shared memory and
// emitted cooperatively (one (query, slot) pair per thread: contiguous index / PPF rows, PPF maths spread evenly).
constexpr int KT_THREADS = 128;

// The per-thread search: the K nearest reference points of (qx,qy,qz) inside segment `sgm`, as a list sorted by
// (squared distance, index) in registers. `tie` is set when an exact-distance tie played any role (see above).
template <int K>
__device__ __forceinline__ void grid_search_thread(float qx, float qy, float qz, int qs, const knngrid::SegHeader& H,
                                                   const int* __restrict__ cs, const float4* __restrict__ pts, float (&bd)[K],
                                                   int (&bi)[K], bool& tie) {
    const int cx = knngrid::cell_coord(qx, H.ox, H.inv_h, H.nx), cy = knngrid::cell_coord(qy, H.oy, H.inv_h, H.ny),
              cz = knngrid::cell_coord(qz, H.oz, H.inv_h, H.nz);
#pragma unroll
    for (int j = 0; j < K; ++j) { bd[j] = 1e10f; bi[j] = qs; }
    float tau = 1e10f;
    int tau_i = qs;

    auto scan_range = [&](int beg, int end) {
        for (int p = beg; p < end; ++p) {
            const float4 v = __ldg(pts + p);
            const float cd = sqdist_ref(qx - v.x, qy - v.y, qz - v.z);
            if (cd <= tau) {
                const int ci = __float_as_int(v.w);
                if (cd == tau && tau != 1e10f) tie = true;     // boundary tie: the reference's choice depends on its heap
                if (cd < tau || ci < tau_i) {
                    bd[K - 1] = cd; bi[K - 1] = ci;
#pragma unroll
                    for (int j = K - 1; j > 0; --j) {
                        const bool sw = bd[j] < bd[j - 1] || (bd[j] == bd[j - 1] && bi[j] < bi[j - 1]);
                        const float d0 = bd[j - 1], d1 = bd[j];
                        const int i0 = bi[j - 1], i1 = bi[j];
                        bd[j - 1] = sw ? d1 : d0; bd[j] = sw ? d0 : d1;
                        bi[j - 1] = sw ? i1 : i0; bi[j] = sw ? i0 : i1;
                    }
                    const float tau_old = tau;
                    tau = bd[K - 1]; tau_i = bi[K - 1];
                    if (tau == tau_old && tau_old != 1e10f) tie = true;
                }
            }
        }
    };

    for (int r = 0;; ++r) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, H.nx - 1);
        for (int dz = -r; dz <= r; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= H.nz) continue;
            for (int dy = -r; dy <= r; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= H.ny) continue;
                const int row = (z * H.ny + y) * H.nx;
                if (abs(dz) == r || abs(dy) == r) {               // face of the cube: the whole x span
                    scan_range(__ldg(cs + row + x0), __ldg(cs + row + x1 + 1));
                } else {                                         // interior row: only the two end cells
                    if (cx - r >= 0) scan_range(__ldg(cs + row + cx - r), __ldg(cs + row + cx - r + 1));
                    if (cx + r < H.nx) scan_range(__ldg(cs + row + cx + r), __ldg(cs + row + cx + r + 1));
                }
            }
        }
        // everything outside the cube of radius r is at least `bound` away from the query
        const bool all = (cx - r <= 0) && (cx + r >= H.nx - 1) && (cy - r <= 0) && (cy + r >= H.ny - 1) && (cz - r <= 0) &&
                         (cz + r >= H.nz - 1);
        if (all) break;
        float bound = CUDART_INF_F;
        if (cx - r > 0) bound = fminf(bound, qx - (H.ox + (float)(cx - r) * H.h));
        if (cx + r < H.nx - 1) bound = fminf(bound, (H.ox + (float)(cx + r + 1) * H.h) - qx);
        if (cy - r > 0) bound = fminf(bound, qy - (H.oy + (float)(cy - r) * H.h));
        if (cy + r < H.ny - 1) bound = fminf(bound, (H.oy + (float)(cy + r + 1) * H.h) - qy);
        if (cz - r > 0) bound = fminf(bound, qz - (H.oz + (float)(cz - r) * H.h));
        if (cz + r < H.nz - 1) bound = fminf(bound, (H.oz + (float)(cz + r + 1) * H.h) - qz);
        bound -= 1e-4f * H.h;                                     // binning rounds (v - o) * inv_h: keep a safety margin
        if (tau < 1e10f && bound > 0.f && tau < bound * bound * 0.99999f) break;
    }
    // equal distances inside the final list: the reference's order among them is its heap's
#pragma unroll
    for (int j = 0; j + 1 < K; ++j)
        if (bd[j] == bd[j + 1] && bd[j] != 1e10f) tie = true;
}


template <int K>
__global__ void __launch_bounds__(KT_THREADS) knn_grid_thread_kernel(const KnnParams P, const knngrid::SegHeader* __restrict__ hdr,
                                                                      const int* __restrict__ cell_start,
                                                                      const float4* __restrict__ sorted,
                                                                      const float4* __restrict__ qorder) {
    __shared__ float s_d[KT_THREADS * K];
    __shared__ int s_i[KT_THREADS * K];
    __shared__ int s_q[KT_THREADS];
    __shared__ float s_qp[KT_THREADS * 3];
    const int tid = threadIdx.x;
    const int t = blockIdx.x * KT_THREADS + tid;
    int q = -1;
    bool tie = false;
    float bd[K];
    int bi[K];
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (t < P.m) {
        const int sgm = find_segment(t, P.new_offset, P.b);
        if (qorder) {
            const float4 v = __ldg(qorder + t);
            q = __float_as_int(v.w); qx = v.x; qy = v.y; qz = v.z;
        } else {
            q = t;
            qx = __ldg(P.qxyz + 3 * (size_t)q); qy = __ldg(P.qxyz + 3 * (size_t)q + 1); qz = __ldg(P.qxyz + 3 * (size_t)q + 2);
        }
        const int qs = sgm == 0 ? 0 : __ldg(P.offset + sgm - 1);
        const knngrid::SegHeader H = hdr[sgm];
        grid_search_thread<K>(qx, qy, qz, qs, H, cell_start + H.cell_base, sorted + qs, bd, bi, tie);
    }
    // ---- stage and emit cooperatively ----
    s_q[tid] = tie ? (-2 - q) : q;            // q >= 0: valid; -1: no query; <= -2: tie-marked query -2 - q
    s_qp[3 * tid] = qx; s_qp[3 * tid + 1] = qy; s_qp[3 * tid + 2] = qz;
#pragma unroll
    for (int j = 0; j < K; ++j) { s_d[tid * K + j] = bd[j]; s_i[tid * K + j] = bi[j]; }
    __syncthreads();
    const int kout = K - P.drop;
    for (int e = tid; e < KT_THREADS * kout; e += KT_THREADS) {
        const int ql = e / kout, slot = e - ql * kout;
        const int qq = s_q[ql];
        if (qq == -1) continue;
        if (qq <= -2) {
            if (slot == 0) P.idx[(size_t)(-2 - qq) * kout] = -1;      // marker consumed by knn_tie_fixup_kernel
            continue;
        }
        emit_result(P, qq, slot, kout, s_i[ql * K + slot + P.drop], s_d[ql * K + slot + P.drop], s_qp[3 * ql], s_qp[3 * ql + 1],
                    s_qp[3 * ql + 2]);
    }
}

// ---- one thread per query, k best in a shared-memory MAX-HEAP (17 slots) -------------------------------------------------
// The register list above costs K-1 compare-exchange steps per admission (128 instructions at K = 17, and a warp pays them
// whenever ANY of its 32 queries admits a candidate), which is why 17-slot queries used to run on the warp-per-query kernel
// (5.2 ns per query against 1.6 ns for the 9-slot thread kernel). Here the k best live in a binary max-heap keyed by
// (distance, index) in shared memory (column `tid` of a [K][KT_THREADS] array: conflict-free): an admission replaces the
// root and sifts down at most 4 levels; the root IS the admission bound tau. A heap sort at the end yields the ascending
// list. Same candidates, same lexicographic admission rule, same tie flags as the register kernel -> bit-identical results.
template <int K>
__global__ void __launch_bounds__(KT_THREADS) knn_grid_thread_heap_kernel(const KnnParams P, const knngrid::SegHeader* __restrict__ hdr,
                                                                           const int* __restrict__ cell_start,
                                                                           const float4* __restrict__ sorted,
                                                                           const float4* __restrict__ qorder) {
    __shared__ float s_d[K * KT_THREADS];       // [slot][thread]
    __shared__ int s_i[K * KT_THREADS];
    __shared__ int s_q[KT_THREADS];
    __shared__ float s_qp[KT_THREADS * 3];
    const int tid = threadIdx.x;
    const int t = blockIdx.x * KT_THREADS + tid;
    float* hd = s_d + tid;
    int* hi = s_i + tid;
    constexpr int S = KT_THREADS;
    int q = -1;
    bool tie = false;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    // (d1, i1) > (d2, i2) lexicographically
    auto gt = [](float d1, int i1, float d2, int i2) { return d1 > d2 || (d1 == d2 && i1 > i2); };
    // put (cd, ci) at the root of the heap of `n` elements and restore the heap property
    auto sift_root = [&](float cd, int ci, int n) {
        int pos = 0;
        while (true) {
            const int l = 2 * pos + 1;
            if (l >= n) break;
            int c = l;
            float dc = hd[l * S];
            int ic = hi[l * S];
            if (l + 1 < n) {
                const float dr = hd[(l + 1) * S];
                const int ir = hi[(l + 1) * S];
                if (gt(dr, ir, dc, ic)) { c = l + 1; dc = dr; ic = ir; }
            }
            if (!gt(dc, ic, cd, ci)) break;
            hd[pos * S] = dc; hi[pos * S] = ic;
            pos = c;
        }
        hd[pos * S] = cd; hi[pos * S] = ci;
    };
    if (t < P.m) {
        const int sgm = find_segment(t, P.new_offset, P.b);
        if (qorder) {
            const float4 v = __ldg(qorder + t);
            q = __float_as_int(v.w); qx = v.x; qy = v.y; qz = v.z;
        } else {
            q = t;
            qx = __ldg(P.qxyz + 3 * (size_t)q); qy = __ldg(P.qxyz + 3 * (size_t)q + 1); qz = __ldg(P.qxyz + 3 * (size_t)q + 2);
        }
        const int qs = sgm == 0 ? 0 : __ldg(P.offset + sgm - 1);
        const knngrid::SegHeader H = hdr[sgm];
        const int* cs = cell_start + H.cell_base;
        const float4* pts = sorted + qs;
#pragma unroll
        for (int j = 0; j < K; ++j) { hd[j * S] = 1e10f; hi[j * S] = qs; }
        float tau = 1e10f;
        int tau_i = qs;
        const int cx = knngrid::cell_coord(qx, H.ox, H.inv_h, H.nx), cy = knngrid::cell_coord(qy, H.oy, H.inv_h, H.ny),
                  cz = knngrid::cell_coord(qz, H.oz, H.inv_h, H.nz);
        auto scan_range = [&](int beg, int end) {
            for (int p = beg; p < end; ++p) {
                const float4 v = __ldg(pts + p);
                const float cd = sqdist_ref(qx - v.x, qy - v.y, qz - v.z);
                if (cd <= tau) {
                    const int ci = __float_as_int(v.w);
                    if (cd == tau && tau != 1e10f) tie = true;     // boundary tie: the reference's choice depends on its heap
                    if (cd < tau || ci < tau_i) {
                        sift_root(cd, ci, K);
                        const float tau_old = tau;
                        tau = hd[0]; tau_i = hi[0];
                        if (tau == tau_old && tau_old != 1e10f) tie = true;
                    }
                }
            }
        };
        for (int r = 0;; ++r) {          // the cube walk of grid_search_thread
            const int x0 = max(cx - r, 0), x1 = min(cx + r, H.nx - 1);
            for (int dz = -r; dz <= r; ++dz) {
                const int z = cz + dz;
                if (z < 0 || z >= H.nz) continue;
                for (int dy = -r; dy <= r; ++dy) {
                    const int y = cy + dy;
                    if (y < 0 || y >= H.ny) continue;
                    const int row = (z * H.ny + y) * H.nx;
                    if (abs(dz) == r || abs(dy) == r) {
                        scan_range(__ldg(cs + row + x0), __ldg(cs + row + x1 + 1));
                    } else {
                        if (cx - r >= 0) scan_range(__ldg(cs + row + cx - r), __ldg(cs + row + cx - r + 1));
                        if (cx + r < H.nx) scan_range(__ldg(cs + row + cx + r), __ldg(cs + row + cx + r + 1));
                    }
                }
            }
            const bool all = (cx - r <= 0) && (cx + r >= H.nx - 1) && (cy - r <= 0) && (cy + r >= H.ny - 1) && (cz - r <= 0) &&
                             (cz + r >= H.nz - 1);
            if (all) break;
            float bound = CUDART_INF_F;
            if (cx - r > 0) bound = fminf(bound, qx - (H.ox + (float)(cx - r) * H.h));
            if (cx + r < H.nx - 1) bound = fminf(bound, (H.ox + (float)(cx + r + 1) * H.h) - qx);
            if (cy - r > 0) bound = fminf(bound, qy - (H.oy + (float)(cy - r) * H.h));
            if (cy + r < H.ny - 1) bound = fminf(bound, (H.oy + (float)(cy + r + 1) * H.h) - qy);
            if (cz - r > 0) bound = fminf(bound, qz - (H.oz + (float)(cz - r) * H.h));
            if (cz + r < H.nz - 1) bound = fminf(bound, (H.oz + (float)(cz + r + 1) * H.h) - qz);
            bound -= 1e-4f * H.h;
            if (tau < 1e10f && bound > 0.f && tau < bound * bound * 0.99999f) break;
        }
        // heap sort: ascending (distance, index) in slots 0 .. K-1
        for (int n = K - 1; n > 0; --n) {
            const float rd = hd[0], ld = hd[n * S];
            const int ri = hi[0], li = hi[n * S];
            hd[n * S] = rd; hi[n * S] = ri;
            sift_root(ld, li, n);
        }
        // equal distances inside the final list: the reference's order among them is its heap's
        for (int j = 0; j + 1 < K; ++j)
            if (hd[j * S] == hd[(j + 1) * S] && hd[j * S] != 1e10f) tie = true;
    }
    // ---- emit cooperatively (one (query, slot) per thread: contiguous index / PPF rows) ----
    s_q[tid] = tie ? (-2 - q) : q;
    s_qp[3 * tid] = qx; s_qp[3 * tid + 1] = qy; s_qp[3 * tid + 2] = qz;
    __syncthreads();
    const int kout = K - P.drop;
    for (int e = tid; e < KT_THREADS * kout; e += KT_THREADS) {
        const int ql = e / kout, slot = e - ql * kout;
        const int qq = s_q[ql];
        if (qq == -1) continue;
        if (qq <= -2) {
            if (slot == 0) P.idx[(size_t)(-2 - qq) * kout] = -1;      // marker consumed by knn_tie_fixup_kernel
            continue;
        }
        emit_result(P, qq, slot, kout, s_i[(slot + P.drop) * S + ql], s_d[(slot + P.drop) * S + ql], s_qp[3 * ql], s_qp[3 * ql + 1],
                    s_qp[3 * ql + 2]);
    }
}

// ---- surface normals: kNN + covariance + smallest-eigenvalue eigenvector, one thread per point ---------------------------
// The step before the hot path: Open3D's estimate_normals(KDTreeSearchParamKNN(knn)) followed by
// dataset/common.py:312-320 normal_redirect, as called in dataset/tdmatch.py:120-127. Per point: the knn nearest points of
// its own cloud (the point itself included), the covariance of those points from first and second cumulants (what Open3D's
// ComputeCovariance does), the eigenvector of the smallest eigenvalue by the non-iterative symmetric 3x3 solver Open3D
// uses for fast_normal_computation (Eberly, "A Robust Eigensolver for 3x3 Symmetric Matrices"), all in fp64; then the
// sign is chosen so that the normal points towards the view point. Fewer than 3 neighbours: (0, 0, 1) like Open3D.
__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
// eigenvector of the symmetric matrix A for eigenvalue ev: the largest cross product of two rows of A - ev I
__device__ __forceinline__ void eigenvector0(const double (&A)[6], double ev, double (&out)[3]) {
    const double r0[3] = {A[0] - ev, A[1], A[2]}, r1[3] = {A[1], A[3] - ev, A[4]}, r2[3] = {A[2], A[4], A[5] - ev};
    double c01[3], c02[3], c12[3];
    cross3(r0, r1, c01); cross3(r0, r2, c02); cross3(r1, r2, c12);
    const double d0 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    const double d1 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    const double d2 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    const double dm = fmax(d0, fmax(d1, d2));
    const double* best = (d0 >= d1 && d0 >= d2) ? c01 : (d1 >= d2 ? c02 : c12);
    const double inv = dm > 0.0 ? rsqrt(dm) : 0.0;
    out[0] = best[0] * inv; out[1] = best[1] * inv; out[2] = best[2] * inv;
}
// second eigenvector, orthogonal to evec0, for eigenvalue ev (Eberly's ComputeEigenvector1)
__device__ __forceinline__ void eigenvector1(const double (&A)[6], const double (&e0)[3], double ev, double (&out)[3]) {
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) {
        const double inv = 1.0 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
        U[0] = -e0[2] * inv; U[1] = 0.0; U[2] = e0[0] * inv;
    } else {
        const double inv = 1.0 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
        U[0] = 0.0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
    }
    cross3(e0, U, V);
    const double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[3] * U[1] + A[4] * U[2], A[2] * U[0] + A[4] * U[1] + A[5] * U[2]};
    const double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[3] * V[1] + A[4] * V[2], A[2] * V[0] + A[4] * V[1] + A[5] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - ev;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - ev;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0.0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m01 * U[i] - m00 * V[i];
        } else { for (int i = 0; i < 3; ++i) out[i] = U[i]; }
    } else {
        if (fmax(a11, a01) > 0.0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m11 * U[i] - m01 * V[i];
        } else { for (int i = 0; i < 3; ++i) out[i] = U[i]; }
    }
}
// unit eigenvector of the SMALLEST eigenvalue of the symmetric matrix {a00,a01,a02,a11,a12,a22}; false if A == 0
__device__ bool smallest_eigenvector(const double (&Ain)[6], double (&n)[3]) {
    double mx = 0.0;
    for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(Ain[i]));
    if (!(mx > 0.0)) return false;
    double A[6];
    for (int i = 0; i < 6; ++i) A[i] = Ain[i] / mx;
    const double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0.0) {
        const double q = (A[0] + A[3] + A[5]) / 3.0;
        const double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * norm) / 6.0);
        const double c00 = b11 * b22 - A[4] * A[4], c01 = A[1] * b22 - A[4] * A[2], c02 = A[1] * A[4] - b11 * A[2];
        const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        const double half = fmin(fmax(0.5 * det, -1.0), 1.0);
        const double angle = acos(half) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = 2.0 * cos(angle), beta0 = 2.0 * cos(angle + two_thirds_pi), beta1 = -(beta0 + beta2);
        const double ev0 = q + p * beta0, ev1 = q + p * beta1, ev2 = q + p * beta2;   // ev0 <= ev1 <= ev2
        if (half >= 0.0) {      // ev2 is the best separated eigenvalue: start from its eigenvector
            double e2[3], e1[3];
            eigenvector0(A, ev2, e2);
            eigenvector1(A, e2, ev1, e1);
            cross3(e1, e2, n);
        } else {
            eigenvector0(A, ev0, n);
        }
    } else {                    // diagonal matrix
        const int k = (A[0] <= A[3] && A[0] <= A[5]) ? 0 : (A[3] <= A[5] ? 1 : 2);
        n[0] = k == 0; n[1] = k == 1; n[2] = k == 2;
    }
    const double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(l > 0.0)) return false;
    n[0] /= l; n[1] /= l; n[2] /= l;
    return true;
}

template <int K>
__global__ void __launch_bounds__(KT_THREADS) normals_kernel(int b, int n, const float* __restrict__ xyz, const int* __restrict__ offset,
                                                            const knngrid::SegHeader* __restrict__ hdr,
                                                            const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                                                            float vx, float vy, float vz, float* __restrict__ normals) {
    const int t = blockIdx.x * KT_THREADS + threadIdx.x;
    if (t >= n) return;
    const int sgm = find_segment(t, offset, b);
    const float4 me = __ldg(sorted + t);                            // points are visited in cell order
    const int q = __float_as_int(me.w);
    const int qs = sgm == 0 ? 0 : __ldg(offset + sgm - 1);
    const knngrid::SegHeader H = hdr[sgm];
    float bd[K];
    int bi[K];
    bool tie = false;
    grid_search_thread<K>(me.x, me.y, me.z, qs, H, cell_start + H.cell_base, sorted + qs, bd, bi, tie);
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        if (bd[j] >= 1e10f) break;                                   // unfilled slots (segment smaller than K)
        const float* p = xyz + 3 * (size_t)bi[j];
        const double x = (double)__ldg(p), y = (double)__ldg(p + 1), z = (double)__ldg(p + 2);
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z; c[6] += y * y; c[7] += y * z; c[8] += z * z;
        ++cnt;
    }
    double nrm[3] = {0.0, 0.0, 1.0};
    if (cnt >= 3) {
        const double inv = 1.0 / (double)cnt;
        for (int i = 0; i < 9; ++i) c[i] *= inv;
        const double A[6] = {c[3] - c[0] * c[0], c[4] - c[0] * c[1], c[5] - c[0] * c[2], c[6] - c[1] * c[1], c[7] - c[1] * c[2],
                             c[8] - c[2] * c[2]};
        double e[3];
        if (smallest_eigenvector(A, e)) { nrm[0] = e[0]; nrm[1] = e[1]; nrm[2] = e[2]; }
    }
    // normal_redirect (dataset/common.py:312-320): towards the view point
    const double dot = ((double)vx - me.x) * nrm[0] + ((double)vy - me.y) * nrm[1] + ((double)vz - me.z) * nrm[2];
    const double sgn = dot < 0.0 ? -1.0 : 1.0;
    normals[3 * (size_t)q] = (float)(sgn * nrm[0]);
    normals[3 * (size_t)q + 1] = (float)(sgn * nrm[1]);
    normals[3 * (size_t)q + 2] = (float)(sgn * nrm[2]);
}

                              // for k <= 9 slots (measured: 2-4x faster there, slower for 17 slots), 2 = thread-per-query everywhere
constexpr float GRID_TARGET = 0.5f;   // average points per grid cell over the bounding box (scripts/tune_grid.py on B200: best for every query type of the forward)

int launch_knn(const KnnParams& P, cudaStream_t st) {
    if (P.m == 0) return ROITR_OK;
    const int smem = 2 * TILE_BYTES;
    // many queries: 4 per warp (amortises the shared-memory reads); few queries: 1 per warp (more CTAs in flight)
    const bool wide = P.m >= 4 * KNN_WARPS * 148 * 2;
    if (wide) {
        static bool attr4_dev[ROITR_MAX_DEVICES] = {};
        bool& attr4 = attr4_dev[roitr_cur_device()];
        if (!attr4) {
            ROITR_CUDA(cudaFuncSetAttribute(knn_ppf_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr4 = true;
        }
        knn_ppf_kernel<4><<<ceil_div(P.m, KNN_WARPS * 4), KNN_THREADS, smem, st>>>(P);
    } else {
        static bool attr1_dev[ROITR_MAX_DEVICES] = {};
        bool& attr1 = attr1_dev[roitr_cur_device()];
        if (!attr1) {
            ROITR_CUDA(cudaFuncSetAttribute(knn_ppf_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr1 = true;
        }
        knn_ppf_kernel<1><<<ceil_div(P.m, KNN_WARPS), KNN_THREADS, smem, st>>>(P);
    }
    ROITR_CHECK_LAUNCH("knn_ppf_kernel");
    knn_tie_fixup_kernel<<<ceil_div(P.m, FIX_WARPS * 32), FIX_WARPS * 32, 0, st>>>(P);
    ROITR_CHECK_LAUNCH("knn_tie_fixup_kernel");
    return ROITR_OK;
}

}  // namespace

// n_total (rows of xyz) is needed on the host to clip the TMA tiles. The reference ABI does not carry it
// (knnquery_cuda_kernel.h:13): the *_n entry points take it from the caller (no host sync; what the engine uses), the
// reference-shaped ones read offset[b-1] back (one 4-byte D2H + stream sync, like the reference's own .item() calls).
static int knn_common(int b, int m, int nslots, int drop, const float* xyz, const float* nrm, const float* qxyz,
                      const float* qnrm, const int* offset, const int* new_offset, int* idx, float* dist, float* ppf,
                      int dist_squared, int n_total, void* stream) {
    ROITR_CHECK_ARG(b >= 1 && m >= 0, "knn: bad b=%d m=%d", b, m);
    ROITR_CHECK_ARG(nslots >= 1 && nslots <= 32, "knn: nsample(+drop) must be in [1,32], got %d", nslots);
    ROITR_CHECK_ARG(drop >= 0 && drop < nslots, "knn: bad drop_first=%d", drop);
    ROITR_CHECK_ARG(xyz && qxyz && offset && new_offset && idx, "knn: null pointer");
    ROITR_CHECK_ARG(!ppf || (nrm && qnrm), "knn: ppf output needs normals");
    ROITR_CHECK_ARG(!ppf || ((uintptr_t)ppf % 16 == 0), "knn: ppf must be 16-byte aligned");
    KnnParams P;
    P.xyz = xyz; P.nrm = nrm; P.qxyz = qxyz; P.qnrm = qnrm; P.offset = offset; P.new_offset = new_offset;
    P.b = b; P.m = m; P.nslots = nslots; P.drop = drop; P.idx = idx; P.dist = dist; P.ppf = ppf;
    P.dist_squared = dist_squared;
    P.use_tma = (n_total > 0) && ((uintptr_t)xyz % 16 == 0);
    P.n_total = n_total > 0 ? n_total : 0x7fffffff;
    return launch_knn(P, (cudaStream_t)stream);
}


extern "C" int roitr_knnquery_n(int b, int m, int nsample, int n_total, const float* xyz, const float* new_xyz,
                                const int* offset, const int* new_offset, int* idx, float* dist2, void* stream) {
    return knn_common(b, m, nsample, 0, xyz, nullptr, new_xyz, nullptr, offset, new_offset, idx, dist2, nullptr, 1,
                      n_total, stream);
}

extern "C" int roitr_knnquery(int b, int m, int nsample, const float* xyz, const float* new_xyz, const int* offset,
                              const int* new_offset, int* idx, float* dist2, void* stream) {
    // reference ABI: the row count of xyz is not passed. Read it from offset[b-1] (one 4-byte D2H on `stream`).
    int n_total = 0;
    ROITR_CHECK_ARG(offset != nullptr && b >= 1, "knnquery: bad offset/b");
    ROITR_CUDA(cudaMemcpyAsync(&n_total, offset + (b - 1), sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ROITR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return roitr_knnquery_n(b, m, nsample, n_total, xyz, new_xyz, offset, new_offset, idx, dist2, stream);
}

extern "C" int roitr_knn_ppf_n(int b, int m, int k_out, int drop_first, int n_total, const float* xyz,
                               const float* normals, const float* new_xyz, const float* new_normals, const int* offset,
                               const int* new_offset, int* idx, float* dist, float* ppf, void* stream) {
    return knn_common(b, m, k_out + drop_first, drop_first, xyz, normals, new_xyz, new_normals, offset, new_offset, idx,
                      dist, ppf, 0, n_total, stream);
}

extern "C" int roitr_knn_ppf(int b, int m, int k_out, int drop_first, const float* xyz, const float* normals,
                             const float* new_xyz, const float* new_normals, const int* offset, const int* new_offset,
                             int* idx, float* dist, float* ppf, void* stream) {
    int n_total = 0;
    ROITR_CHECK_ARG(offset != nullptr && b >= 1, "knn_ppf: bad offset/b");
    ROITR_CUDA(cudaMemcpyAsync(&n_total, offset + (b - 1), sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ROITR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return roitr_knn_ppf_n(b, m, k_out, drop_first, n_total, xyz, normals, new_xyz, new_normals, offset, new_offset, idx,
                           dist, ppf, stream);
}


// ---- grid API -----------------------------------------------------------------------------------------------------------
static size_t grid_hdr_bytes(int b) { return ((size_t)b * sizeof(knngrid::SegHeader) + 255) / 256 * 256; }
static size_t grid_cells_bytes(int b) { return ((size_t)b * (knngrid::MAX_CELLS + 1) * sizeof(int) + 255) / 256 * 256; }

extern "C" long long roitr_knn_grid_workspace_bytes(int b, int n) {
    return (long long)(grid_hdr_bytes(b) + 2 * grid_cells_bytes(b) + (size_t)n * sizeof(float4));
}

/* byte offset, inside a grid workspace of b segments, of the cell-sorted (x, y, z, index) float4 array */
extern "C" long long roitr_knn_grid_sorted_offset(int b) { return (long long)(grid_hdr_bytes(b) + 2 * grid_cells_bytes(b)); }

extern "C" int roitr_knn_grid_build_target(int b, int n, const float* xyz, const int* offset, float target_per_cell, void* workspace,
                                           void* stream) {
    ROITR_CHECK_ARG(b >= 1 && n >= 0 && xyz && offset && workspace && target_per_cell > 0.f, "knn_grid_build: bad arguments");
    ROITR_CHECK_ARG((uintptr_t)workspace % 256 == 0, "knn_grid_build: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)workspace;
    auto* hdr = (knngrid::SegHeader*)w;
    int* cell_start = (int*)(w + grid_hdr_bytes(b));
    int* cursor = (int*)(w + grid_hdr_bytes(b) + grid_cells_bytes(b));
    float4* sorted = (float4*)(w + grid_hdr_bytes(b) + 2 * grid_cells_bytes(b));
    ROITR_CUDA(cudaMemsetAsync(cursor, 0, grid_cells_bytes(b), st));
    knngrid::grid_header_kernel<<<b, 1024, 0, st>>>(b, xyz, offset, hdr, target_per_cell);
    if (n > 0) knngrid::grid_bin_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, b, xyz, offset, hdr, cell_start, cursor, sorted, 0);
    knngrid::grid_scan_kernel<<<b, 1024, 0, st>>>(hdr, cursor, cell_start);
    if (n > 0) knngrid::grid_bin_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, b, xyz, offset, hdr, cell_start, cursor, sorted, 1);
    ROITR_CHECK_LAUNCH("knn_grid_build");
    return ROITR_OK;
}

extern "C" int roitr_knn_grid_build(int b, int n, const float* xyz, const int* offset, void* workspace, void* stream) {
    return roitr_knn_grid_build_target(b, n, xyz, offset, GRID_TARGET, workspace, stream);
}

extern "C" int roitr_knn_ppf_grid_q(int b, int m, int k_out, int drop_first, int n_total, const float* xyz,
                                    const float* normals, const float* new_xyz, const float* new_normals, const int* offset,
                                    const int* new_offset, const void* workspace, const void* query_workspace, int* idx,
                                    float* dist, float* ppf, void* stream) {
    const int nslots = k_out + drop_first;
    ROITR_CHECK_ARG(b >= 1 && m >= 0 && nslots >= 1 && nslots <= 32 && drop_first >= 0 && drop_first < nslots, "knn_ppf_grid: bad sizes");
    ROITR_CHECK_ARG(xyz && new_xyz && offset && new_offset && idx && workspace, "knn_ppf_grid: null pointer");
    ROITR_CHECK_ARG(!ppf || (normals && new_normals && (uintptr_t)ppf % 16 == 0), "knn_ppf_grid: ppf needs normals / alignment");
    if (m == 0) return ROITR_OK;
    KnnParams P;
    P.xyz = xyz; P.nrm = normals; P.qxyz = new_xyz; P.qnrm = new_normals; P.offset = offset; P.new_offset = new_offset;
    P.b = b; P.m = m; P.nslots = nslots; P.drop = drop_first; P.idx = idx; P.dist = dist; P.ppf = ppf; P.dist_squared = 0;
    P.use_tma = 0; P.n_total = n_total;
    const unsigned char* w = (const unsigned char*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    const auto* hdr = (const knngrid::SegHeader*)w;
    const int* cell_start = (const int*)(w + grid_hdr_bytes(b));
    const float4* sorted = (const float4*)(w + grid_hdr_bytes(b) + 2 * grid_cells_bytes(b));
    // visiting order of the queries: the cell order of the query set's own grid (self queries: the same grid)
    const unsigned char* qw = (const unsigned char*)query_workspace;
    if (!qw && new_xyz == xyz && m == n_total) qw = w;
    const float4* qorder = qw ? (const float4*)(qw + grid_hdr_bytes(b) + 2 * grid_cells_bytes(b)) : nullptr;
    const int grid = ceil_div(m, KT_THREADS);
    // thread-per-query: sorted list in registers for 1 / 3 / 9 slots, a shared-memory max-heap for 17 (the unrolled insertion
    // of a 17-entry register list diverges too much); any other slot count: the warp-per-query kernel (top-k over the lanes)
    bool done = true;
    if (nslots == 1) knn_grid_thread_kernel<1><<<grid, KT_THREADS, 0, st>>>(P, hdr, cell_start, sorted, qorder);
    else if (nslots == 3) knn_grid_thread_kernel<3><<<grid, KT_THREADS, 0, st>>>(P, hdr, cell_start, sorted, qorder);
    else if (nslots == 9) knn_grid_thread_heap_kernel<9><<<grid, KT_THREADS, 0, st>>>(P, hdr, cell_start, sorted, qorder);
    else if (nslots == 17) knn_grid_thread_heap_kernel<17><<<grid, KT_THREADS, 0, st>>>(P, hdr, cell_start, sorted, qorder);
    else done = false;
    if (!done) knn_grid_kernel<<<ceil_div(m, KNN_WARPS), KNN_THREADS, 0, st>>>(P, hdr, cell_start, sorted);
    ROITR_CHECK_LAUNCH("knn_grid_kernel");
    knn_tie_fixup_kernel<<<ceil_div(m, FIX_WARPS * 32), FIX_WARPS * 32, 0, st>>>(P);
    ROITR_CHECK_LAUNCH("knn_tie_fixup_kernel");
    return ROITR_OK;
}

extern "C" int roitr_knn_ppf_grid(int b, int m, int k_out, int drop_first, int n_total, const float* xyz,
                                  const float* normals, const float* new_xyz, const float* new_normals, const int* offset,
                                  const int* new_offset, const void* workspace, int* idx, float* dist, float* ppf,
                                  void* stream) {
    return roitr_knn_ppf_grid_q(b, m, k_out, drop_first, n_total, xyz, normals, new_xyz, new_normals, offset, new_offset,
                                workspace, nullptr, idx, dist, ppf, stream);
}

extern "C" int roitr_estimate_normals(int b, int n, int knn, const float* xyz, const int* offset, const void* workspace,
                                      const float* view_point, float* normals, void* stream) {
    ROITR_CHECK_ARG(b >= 1 && n >= 0 && xyz && offset && workspace && view_point && normals, "estimate_normals: bad arguments");
    ROITR_CHECK_ARG(knn == 9 || knn == 17 || knn == 33, "estimate_normals: knn must be 9, 17 or 33 (Open3D call site: 33), got %d", knn);
    if (n == 0) return ROITR_OK;
    const unsigned char* w = (const unsigned char*)workspace;
    const auto* hdr = (const knngrid::SegHeader*)w;
    const int* cell_start = (const int*)(w + grid_hdr_bytes(b));
    const float4* sorted = (const float4*)(w + grid_hdr_bytes(b) + 2 * grid_cells_bytes(b));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(n, KT_THREADS);
    const float vx = view_point[0], vy = view_point[1], vz = view_point[2];
    if (knn == 9) normals_kernel<9><<<grid, KT_THREADS, 0, st>>>(b, n, xyz, offset, hdr, cell_start, sorted, vx, vy, vz, normals);
    else if (knn == 17) normals_kernel<17><<<grid, KT_THREADS, 0, st>>>(b, n, xyz, offset, hdr, cell_start, sorted, vx, vy, vz, normals);
    else normals_kernel<33><<<grid, KT_THREADS, 0, st>>>(b, n, xyz, offset, hdr, cell_start, sorted, vx, vy, vz, normals);
    ROITR_CHECK_LAUNCH("normals_kernel");
    return ROITR_OK;
}
