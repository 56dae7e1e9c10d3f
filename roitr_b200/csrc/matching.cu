// Coarse-to-fine matching head (sm_100a): point-to-node partition, coarse superpoint matching, fused fine scoring +
// log-domain Sinkhorn optimal transport + mutual top-k, and the ordered compaction that replaces torch.nonzero.
//
// Replaces point_to_node_partition (lib/utils.py:428-471), CoarseMatching.forward (model/modules.py:141-178),
// the einsum + LearnableLogOptimalTransport.forward (model/RIGA_v2.py:150-153, modules.py:21-68) and
// FineMatching.forward (modules.py:242-324).
#include <math_constants.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace {

// square_distance(a, b) of lib/utils.py:139-156 for one pair given the matmul term xy: clamp((-2 xy + |a|^2) + |b|^2, 1e-12)
__device__ __forceinline__ float sqdist_matmul_form(float xy, float a2, float b2) {
    return fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, xy), a2), b2), 1e-12f);
}
__device__ __forceinline__ float sumsq3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// ------------------------------------------------------------------------------------------------ partition
// All kernels of the head take a leading PAIR (cloud) dimension: blockIdx.y (or .z) selects one of `B` equally sized
// problems whose arrays lie back to back, so one launch serves every pair of a batch (B = 1: the single-pair forward).
//
// owner[p] = argmin_nodes d(node, p) (first minimum), dmin[p] = that distance, count[node]++.
__global__ void point_owner_kernel(int N, int M, const float* __restrict__ pts, const float* __restrict__ nodes,
                                   int* __restrict__ owner, float* __restrict__ dmin, int* __restrict__ count) {
    extern __shared__ float sn[];  // M x 4: x, y, z, |n|^2
    const size_t cloud = blockIdx.y;
    pts += cloud * N * 3; nodes += cloud * M * 3; owner += cloud * N; dmin += cloud * N; count += cloud * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const float x = __ldg(nodes + 3 * i), y = __ldg(nodes + 3 * i + 1), z = __ldg(nodes + 3 * i + 2);
        sn[4 * i] = x; sn[4 * i + 1] = y; sn[4 * i + 2] = z; sn[4 * i + 3] = sumsq3(x, y, z);
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float px = __ldg(pts + 3 * p), py = __ldg(pts + 3 * p + 1), pz = __ldg(pts + 3 * p + 2);
    const float p2 = sumsq3(px, py, pz);
    float best = CUDART_INF_F;
    int bi = 0;
    for (int i = 0; i < M; ++i) {
        const float4 nd = *reinterpret_cast<const float4*>(sn + 4 * i);
        const float xy = fmaf(nd.z, pz, fmaf(nd.y, py, __fmul_rn(nd.x, px)));
        const float d = sqdist_matmul_form(xy, nd.w, p2);
        if (d < best) { best = d; bi = i; }
    }
    owner[p] = bi;
    dmin[p] = best;
    atomicAdd(count + bi, 1);
}

// start[node] = exclusive prefix sum of count over the nodes of one cloud (one CTA per cloud; M is a few hundred)
__global__ void __launch_bounds__(1024) node_start_kernel(int M, const int* __restrict__ count, int* __restrict__ start) {
    __shared__ int buf[1024];
    __shared__ int carry;
    const size_t cloud = blockIdx.x;
    count += cloud * M; start += cloud * M;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < M; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < M ? count[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < M) start[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += buf[1023];
        __syncthreads();
    }
}

// counting-sort scatter: every point drops its (distance bits, index) key into its owner's bucket (slot order inside a
// bucket is arbitrary; the per-node sort below orders it)
__global__ void node_bucket_kernel(int N, int M, const int* __restrict__ owner, const float* __restrict__ dmin,
                                   const int* __restrict__ start, int* __restrict__ cursor,
                                   unsigned long long* __restrict__ bucket) {
    const size_t cloud = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int o = __ldg(owner + cloud * N + p);
    const int slot = __ldg(start + cloud * M + o) + atomicAdd(cursor + cloud * M + o, 1);
    bucket[cloud * N + slot] = ((unsigned long long)__float_as_uint(__ldg(dmin + cloud * N + p)) << 32) | (unsigned)p;
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int n, int tid, int nthreads) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += nthreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// per node: its own points (its bucket) sorted by (distance, index), first `limit` kept (topk(k=limit, largest=False),
// :459); remaining slots = N (the pad row) with mask false. The reference scans a masked (M,N) matrix per node; here a
// node touches only the ~N/M points it owns.
constexpr int PART_CAP = 4096;
__global__ void __launch_bounds__(256) node_knn_kernel(int N, int M, int limit, const int* __restrict__ count,
                                                       const int* __restrict__ start,
                                                       const unsigned long long* __restrict__ bucket,
                                                       int* __restrict__ knn_idx, unsigned char* __restrict__ knn_mask,
                                                       unsigned char* __restrict__ node_mask) {
    __shared__ unsigned long long keys[PART_CAP];
    const size_t cloud = blockIdx.y;
    const int node = blockIdx.x, tid = threadIdx.x;
    const int cnt = __ldg(count + cloud * M + node);
    const unsigned long long* mine = bucket + cloud * N + __ldg(start + cloud * M + node);
    knn_idx += (cloud * M + node) * (size_t)limit;
    knn_mask += (cloud * M + node) * (size_t)limit;
    if (tid == 0) node_mask[cloud * M + node] = cnt > 0;
    if (cnt <= PART_CAP) {
        int n2 = 64;
        while (n2 < cnt) n2 <<= 1;
        for (int i = tid; i < n2; i += blockDim.x) keys[i] = i < cnt ? mine[i] : ~0ull;
        __syncthreads();
        bitonic_sort_u64(keys, n2, tid, blockDim.x);
        for (int j = tid; j < limit; j += blockDim.x) {
            const bool ok = j < cnt;
            knn_idx[j] = ok ? (int)(keys[j] & 0xffffffffu) : N;
            knn_mask[j] = ok;
        }
    } else {
        // pathological node owning more than PART_CAP points: `limit` rounds of "smallest key greater than the last"
        __shared__ unsigned long long red[256];
        unsigned long long last = 0ull;
        bool first = true;
        for (int j = 0; j < limit; ++j) {
            unsigned long long best = ~0ull;
            for (int i = tid; i < cnt; i += blockDim.x) {
                const unsigned long long key = mine[i];
                if ((first || key > last) && key < best) best = key;
            }
            red[tid] = best;
            __syncthreads();
            for (int s2 = 128; s2 > 0; s2 >>= 1) {
                if (tid < s2 && red[tid + s2] < red[tid]) red[tid] = red[tid + s2];
                __syncthreads();
            }
            last = red[0];
            first = false;
            __syncthreads();
            if (tid == 0) { knn_idx[j] = (int)(last & 0xffffffffu); knn_mask[j] = 1; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ ordered compaction
// out = flat indices of the non-zero flags in ascending order (torch.nonzero order). Three passes.
constexpr int CMP_CHUNK = 2048;
// blockIdx.y = segment: `n` flags per segment, ceil(n / CMP_CHUNK) chunk counters per segment, `capacity` outputs per segment
__global__ void compact_count_kernel(long long n, const unsigned char* __restrict__ flags, int* __restrict__ chunk_count) {
    flags += (size_t)blockIdx.y * n;
    chunk_count += (size_t)blockIdx.y * gridDim.x;
    const long long base = (long long)blockIdx.x * CMP_CHUNK;
    int c = 0;
    for (int i = threadIdx.x; i < CMP_CHUNK; i += blockDim.x)
        if (base + i < n && flags[base + i]) ++c;
    c = __reduce_add_sync(FULL_MASK, c);
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += ws[w];
        chunk_count[blockIdx.x] = t;
    }
}
__global__ void compact_scan_kernel(int nchunks, int* __restrict__ chunk_count, int* __restrict__ total) {
    // one CTA per segment: exclusive scan (nchunks is small: n / 2048)
    chunk_count += (size_t)blockIdx.x * nchunks;
    __shared__ int carry;
    __shared__ int buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nchunks ? chunk_count[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nchunks) chunk_count[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry;
}
__global__ void compact_write_kernel(long long n, const unsigned char* __restrict__ flags,
                                     const int* __restrict__ chunk_offset, int* __restrict__ out, int capacity) {
    // 256 threads, each owns 8 consecutive flags of the chunk -> order preserved
    flags += (size_t)blockIdx.y * n;
    chunk_offset += (size_t)blockIdx.y * gridDim.x;
    out += (size_t)blockIdx.y * (capacity > 0 ? capacity : 1);
    const long long base = (long long)blockIdx.x * CMP_CHUNK + threadIdx.x * 8;
    int c = 0;
    unsigned bits = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (base + i < n && flags[base + i]) { bits |= 1u << i; ++c; }
    __shared__ int sc[256];
    sc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int t = threadIdx.x >= o ? sc[threadIdx.x - o] : 0;
        __syncthreads();
        sc[threadIdx.x] += t;
        __syncthreads();
    }
    int pos = chunk_offset[blockIdx.x] + sc[threadIdx.x] - c;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (bits & (1u << i)) {
            if (pos < capacity) out[pos] = (int)(base + i);
            ++pos;
        }
}

// ------------------------------------------------------------------------------------------------ coarse matching
__global__ void row_sqnorm_kernel(int M, int C, const float* __restrict__ f, float* __restrict__ out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= M) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = __ldg(f + (size_t)row * C + c); s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}
// s_ij = exp(-sqdist) on valid (i,j), else 0; one CTA per row (blockIdx.y = pair), also the row sum (modules.py:163)
__global__ void coarse_exp_kernel(int Mr, int Ms, const float* __restrict__ xy, const float* __restrict__ r2,
                                  const float* __restrict__ s2, const unsigned char* __restrict__ rmask,
                                  const unsigned char* __restrict__ smask, float* __restrict__ S,
                                  float* __restrict__ rowsum) {
    const int i = blockIdx.x;
    const size_t b = blockIdx.y;
    xy += b * Mr * Ms; S += b * Mr * Ms; r2 += b * Mr; s2 += b * Ms; rmask += b * Mr; smask += b * Ms; rowsum += b * Mr;
    __shared__ float red[8];
    float acc = 0.f;
    const bool rv = rmask[i];
    for (int j = threadIdx.x; j < Ms; j += blockDim.x) {
        float v = 0.f;
        if (rv && smask[j]) v = expf(-sqdist_matmul_form(__ldg(xy + (size_t)i * Ms + j), __ldg(r2 + i), __ldg(s2 + j)));
        S[(size_t)i * Ms + j] = v;
        acc += v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        rowsum[i] = t;
    }
}
__global__ void col_sum_kernel(int Mr, int Ms, const float* __restrict__ S, float* __restrict__ colsum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Ms) return;
    S += (size_t)blockIdx.y * Mr * Ms;
    float t = 0.f;
    for (int i = 0; i < Mr; ++i) t += __ldg(S + (size_t)i * Ms + j);
    colsum[(size_t)blockIdx.y * Ms + j] = t;
}

// Flat top-k (largest, sorted descending; ties by ascending flat index) of n floats >= 0 (entries < 0 are excluded), one
// CTA of 1024 threads per pair (blockIdx.x): 4 x 8-bit radix-select passes for the k-th value, collect, bitonic sort.
// With `dn` the kernel first applies the dual normalisation of modules.py:166-169 in place (row / column sums from
// coarse_exp / col_sum; invalid entries become -1 so they are never selected), which used to be a launch of its own.
struct DualNorm {
    const float* rowsum; const float* colsum; const unsigned char* rmask; const unsigned char* smask;
    int Mr, Ms, dual;
};
constexpr int TOPK_MAX = 1024;
__global__ void __launch_bounds__(1024) flat_topk_kernel(int n, int k, float* __restrict__ v, int row_len,
                                                         int* __restrict__ out_row, int* __restrict__ out_col,
                                                         float* __restrict__ out_val, int* __restrict__ out_count,
                                                         int out_stride, DualNorm dn) {
    {
        const size_t b = blockIdx.x;
        v += b * n; out_row += b * out_stride; out_col += b * out_stride; out_val += b * out_stride; out_count += b;
        if (dn.rowsum) { dn.rowsum += b * dn.Mr; dn.colsum += b * dn.Ms; dn.rmask += b * dn.Mr; dn.smask += b * dn.Ms; }
    }
    if (dn.rowsum) {
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int i = e / dn.Ms, j = e % dn.Ms;
            if (!(dn.rmask[i] && dn.smask[j])) { v[e] = -1.f; continue; }
            const float sv = v[e];
            if (dn.dual) v[e] = __fmul_rn(__fdiv_rn(sv, __fadd_rn(dn.rowsum[i], 1e-8f)), __fdiv_rn(sv, __fadd_rn(dn.colsum[j], 1e-8f)));
        }
        __syncthreads();
    }
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining;
    __shared__ int s_valid, s_cnt;
    __shared__ unsigned long long keys[TOPK_MAX];
    const int tid = threadIdx.x;
    // number of candidates
    if (tid == 0) { s_valid = 0; s_cnt = 0; }
    __syncthreads();
    int c = 0;
    for (int i = tid; i < n; i += blockDim.x) c += (v[i] >= 0.f);
    atomicAdd(&s_valid, c);
    __syncthreads();
    const int kk = min(k, s_valid);
    if (kk == 0) { if (tid == 0) *out_count = 0; return; }
    // radix select the kk-th largest bit pattern
    if (tid == 0) { s_prefix = 0; s_remaining = (unsigned)kk; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < n; i += blockDim.x) {
            const float f = v[i];
            if (f < 0.f) continue;
            const unsigned b = __float_as_uint(f);
            if ((b & himask) == (prefix & himask)) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rem = s_remaining;
            int d = 255;
            for (; d > 0; --d) {
                if (hist[d] >= rem) break;
                rem -= hist[d];
            }
            s_prefix = prefix | ((unsigned)d << shift);
            s_remaining = rem;
        }
        __syncthreads();
    }
    const unsigned pivot = s_prefix;      // bit pattern of the kk-th largest value
    const unsigned need_eq = s_remaining; // how many entries equal to the pivot are taken (lowest flat index first)
    // entries strictly greater than the pivot
    for (int i = tid; i < n; i += blockDim.x) {
        const float f = v[i];
        if (f >= 0.f && __float_as_uint(f) > pivot) {
            const int s = atomicAdd(&s_cnt, 1);
            keys[s] = ((unsigned long long)(~__float_as_uint(f)) << 32) | (unsigned)i;  // ascending sort == descending value
        }
    }
    __syncthreads();
    // entries equal to the pivot: lowest flat indices first. Collect them in parallel (ties are rare), sort by index.
    __shared__ unsigned eq_idx[TOPK_MAX];
    __shared__ int s_eq;
    if (tid == 0) s_eq = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const float f = v[i];
        if (f >= 0.f && __float_as_uint(f) == pivot) {
            const int s = atomicAdd(&s_eq, 1);
            if (s < TOPK_MAX) eq_idx[s] = (unsigned)i;
        }
    }
    __syncthreads();
    const int base_cnt = s_cnt;
    if (s_eq <= TOPK_MAX) {
        // small odd-even transposition sort by index (s_eq is almost always 1)
        const int ne = s_eq;
        for (int round = 0; round < ne; ++round) {
            for (int i = 2 * tid + (round & 1); i + 1 < ne; i += 2 * blockDim.x)
                if (eq_idx[i] > eq_idx[i + 1]) { const unsigned t = eq_idx[i]; eq_idx[i] = eq_idx[i + 1]; eq_idx[i + 1] = t; }
            __syncthreads();
        }
        for (int t = tid; t < (int)need_eq; t += blockDim.x)
            keys[base_cnt + t] = ((unsigned long long)(~pivot) << 32) | eq_idx[t];
    } else if (tid == 0) {  // massive ties (e.g. constant scores): sequential walk in index order
        unsigned taken = 0;
        for (int i = 0; i < n && taken < need_eq; ++i)
            if (v[i] >= 0.f && __float_as_uint(v[i]) == pivot) {
                keys[base_cnt + taken] = ((unsigned long long)(~pivot) << 32) | (unsigned)i;
                ++taken;
            }
    }
    __syncthreads();
    int n2 = 2;
    while (n2 < kk) n2 <<= 1;
    for (int i = kk + tid; i < n2; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(keys, n2, tid, blockDim.x);
    for (int j = tid; j < kk; j += blockDim.x) {
        const int flat = (int)(keys[j] & 0xffffffffu);
        out_row[j] = flat / row_len;
        out_col[j] = flat % row_len;
        out_val[j] = __uint_as_float(~(unsigned)(keys[j] >> 32));
    }
    if (tid == 0) *out_count = kk;
}

// ------------------------------------------------------------------------------------------------ fine matching
// One CTA per superpoint correspondence p: gather the two 64-point patches' descriptors, S = Ft Fs^T / sqrt(C),
// 100 log-Sinkhorn iterations with the 65x65 matrix resident in shared memory, write the (65,65) log-assignment,
// then exp / mutual top-k / threshold / validity -> match flags.
struct FineParams {
    const float* tgt_feat; const float* src_feat;      // (Nt, C), (Ns, C) fine descriptors
    int Nt, Ns, C;
    const int* tgt_knn; const int* src_knn;            // (Mt, 64), (Ms, 64) int32, pad = Nt / Ns
    const unsigned char* tgt_kmask; const unsigned char* src_kmask;
    const int* corr_t; const int* corr_s; const int* corr_count;   // (Pmax) node pairs, device count
    const float* alpha;
    float* scores;               // (Pmax, 65, 65)
    unsigned char* flags;        // (Pmax, 64, 64)
    int num_iter, topk, mutual;
    float threshold, sqrt_c;
    int Pmax, Mt, Ms;            // batched launch (blockIdx.y = pair): per-pair strides of every array above
};

constexpr int FP = 64, FP1 = 65;

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int FT = 288, FW = FT / 32;      // 8 warps own rows/columns 0..63 (4 threads each), warp 8 owns the dustbin row/column

#ifndef FINE_MINB
#define FINE_MINB 4
#endif
__global__ void __launch_bounds__(FT, FINE_MINB) fine_patch_kernel(FineParams P) {
    __shared__ float Z[FP1 * FP1];
    __shared__ __align__(16) float At[32][FP + 4];
    __shared__ __align__(16) float Bs[32][FP + 4];
    __shared__ __align__(16) float u[FP1 + 3], v[FP1 + 3];
    __shared__ float log_mu[FP1], log_nu[FP1];
    __shared__ int t_idx[FP], s_idx[FP];
    __shared__ unsigned char t_ok[FP], s_ok[FP];
    __shared__ __align__(16) unsigned char rowf[FP * FP];   // top-k flags; before that, the scale vectors of the Sinkhorn loop
    __shared__ float s_norm;

    const int p = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // blockIdx.y = pair of the batch: every per-pair array advances by its own size
        const size_t b = blockIdx.y;
        P.corr_count += b; P.corr_t += b * P.Pmax; P.corr_s += b * P.Pmax;
        P.tgt_knn += b * P.Mt * FP; P.src_knn += b * P.Ms * FP; P.tgt_kmask += b * P.Mt * FP; P.src_kmask += b * P.Ms * FP;
        P.tgt_feat += b * P.Nt * P.C; P.src_feat += b * P.Ns * P.C;
        P.scores += b * P.Pmax * FP1 * FP1; P.flags += b * P.Pmax * FP * FP;
    }
    if (p >= __ldg(P.corr_count)) return;
    const int nt = __ldg(P.corr_t + p), ns = __ldg(P.corr_s + p);
    if (tid < FP) {
        t_idx[tid] = __ldg(P.tgt_knn + (size_t)nt * FP + tid);
        t_ok[tid] = P.tgt_kmask[(size_t)nt * FP + tid];
    } else if (tid < 2 * FP) {
        s_idx[tid - FP] = __ldg(P.src_knn + (size_t)ns * FP + tid - FP);
        s_ok[tid - FP] = P.src_kmask[(size_t)ns * FP + tid - FP];
    }
    __syncthreads();

    // ---- 64 x 64 x C scores (rows = tgt patch points, cols = src patch points), threads 0..255 ----
    const bool mm = tid < 256;
    const int tx = tid & 15, ty = (tid >> 4) & 15;
    // The contraction runs over C = 256 / 512 terms whose sum reaches |z| sqrt(C) ~ 10^4 for peaky descriptors: one long fp32
    // FMA chain then carries ~1e-4 of rounding noise into z (measured against a float64 evaluation, tests/parity.py - more
    // than the reference's blocked sgemm does). Each 32-term chunk is therefore summed on its own (`part`) and added to a
    // COMPENSATED running total (Knuth TwoSum: `acc` + the exact rounding errors collected in `comp`), which makes z
    // correctly rounded to within an ulp for 7 extra FADDs per output per chunk.
    float acc[4][4], comp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; comp[i][j] = 0.f; }
    const int lrow = (tid >> 2) & 63, lk = (tid & 3) * 8;
    const int ti = t_idx[lrow], si = s_idx[lrow];
    for (int k0 = 0; k0 < P.C; k0 += 32) {
        if (mm) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (ti < P.Nt) a = __ldg(reinterpret_cast<const float4*>(P.tgt_feat + (size_t)ti * P.C + k0 + lk) + h);
                if (si < P.Ns) b = __ldg(reinterpret_cast<const float4*>(P.src_feat + (size_t)si * P.C + k0 + lk) + h);
                At[lk + 4 * h][lrow] = a.x; At[lk + 4 * h + 1][lrow] = a.y; At[lk + 4 * h + 2][lrow] = a.z; At[lk + 4 * h + 3][lrow] = a.w;
                Bs[lk + 4 * h][lrow] = b.x; Bs[lk + 4 * h + 1][lrow] = b.y; Bs[lk + 4 * h + 2][lrow] = b.z; Bs[lk + 4 * h + 3][lrow] = b.w;
            }
        }
        __syncthreads();
        if (mm) {
            float part[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&At[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {       // TwoSum(acc, part): sum + exact error, in round-to-nearest adds
                    const float a0 = acc[i][j], b0 = part[i][j];
                    const float sum = __fadd_rn(a0, b0);
                    const float bb = __fsub_rn(sum, a0);
                    const float err = __fadd_rn(__fsub_rn(a0, __fsub_rn(sum, bb)), __fsub_rn(b0, bb));
                    acc[i][j] = sum;
                    comp[i][j] = __fadd_rn(comp[i][j], err);
                }
        }
        __syncthreads();
    }
    if (mm) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], comp[i][j]);
    }
    // ---- padded, masked score matrix (modules.py:36-46) and marginals (:48-60) ----
    const float alpha = __ldg(P.alpha);
    const float NEG = -1e6f;
    if (mm) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = ty * 4 + i, c = tx * 4 + j;
                Z[r * FP1 + c] = (t_ok[r] && s_ok[c]) ? __fdiv_rn(acc[i][j], P.sqrt_c) : NEG;
            }
    }
    if (tid < FP) {
        Z[tid * FP1 + FP] = t_ok[tid] ? alpha : NEG;   // dustbin column
        Z[FP * FP1 + tid] = s_ok[tid] ? alpha : NEG;   // dustbin row
    }
    if (tid == 0) {
        Z[FP * FP1 + FP] = alpha;
        int nr = 0, nc = 0;
        for (int i = 0; i < FP; ++i) { nr += t_ok[i]; nc += s_ok[i]; }
        const float norm = -logf((float)nr + (float)nc);
        s_norm = norm;
        log_mu[FP] = logf((float)nc) + norm;
        log_nu[FP] = logf((float)nr) + norm;
    }
    __syncthreads();
    if (tid < FP) {
        log_mu[tid] = t_ok[tid] ? s_norm : NEG;
        log_nu[tid] = s_ok[tid] ? s_norm : NEG;
    }
    if (tid < FP1 + 3) { u[tid] = 0.f; v[tid] = 0.f; }
    __syncthreads();
    // ---- log-domain Sinkhorn (modules.py:21-26) ----
    // 200 dependent logsumexp sweeps per patch pair; the kernel's time is the instruction count of this loop, so a
    // logsumexp is split over only FOUR threads: thread (i, q) of warps 0-7 holds elements 16q .. 16q+15 (+ the dustbin
    // element for q = 3) of row i AND of column i of Z in registers for the whole loop, evaluates its 16-17 terms
    // serially (independent FADD / FMNMX / FFMA / EX2 chains, no shuffles) and finishes with a 2-step butterfly; warp 8
    // handles the dustbin row / column across its 32 lanes. Only u and v (65 floats each) go through shared memory, read
    // as float4 broadcasts, one __syncthreads per sweep. exp(x - m) is ex2(fma(x, log2 e, -m log2 e)) and log is lg2 * ln 2
    // (hardware paths, relative error 2^-21, well inside the parity tolerance; the iteration is a contraction so errors
    // do not accumulate).
    {
        constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
        const int ri = (tid >> 2) & 63, q = tid & 3;
        // the 16 terms live in register PAIRS: the adds and FMAs of the sweep are packed (add / fma.rn.f32x2: the same IEEE
        // operation per component, so the results are those of the scalar loop bit for bit, at half the instructions)
        float2 zr[8], zc[8];
        float zr16 = -CUDART_INF_F, zc16 = -CUDART_INF_F;      // the dustbin term (q = 3); warp 8: its third element
        if (warp < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                zr[j] = make_float2(Z[ri * FP1 + q * 16 + 2 * j], Z[ri * FP1 + q * 16 + 2 * j + 1]);
                zc[j] = make_float2(Z[(q * 16 + 2 * j) * FP1 + ri], Z[(q * 16 + 2 * j + 1) * FP1 + ri]);
            }
            if (q == 3) { zr16 = Z[ri * FP1 + 64]; zc16 = Z[64 * FP1 + ri]; }
        } else {
            zr[0] = make_float2(Z[64 * FP1 + lane], Z[64 * FP1 + lane + 32]);
            zc[0] = make_float2(Z[lane * FP1 + 64], Z[(lane + 32) * FP1 + 64]);
            if (lane == 0) { zr16 = Z[64 * FP1 + 64]; zc16 = Z[64 * FP1 + 64]; }
        }
        // returns whether the value this thread wrote differs from the one it replaced (fixed-point detection below)
        auto sweep = [&](const float2 (&z)[8], const float z16, const float* __restrict__ add, const float* __restrict__ marg,
                         float* __restrict__ dst) -> int {
            int changed = 0;
            if (warp < 8) {
                float2 x[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 a = *reinterpret_cast<const float4*>(add + q * 16 + 4 * j4);
                    x[2 * j4] = __fadd2_rn(z[2 * j4], make_float2(a.x, a.y));
                    x[2 * j4 + 1] = __fadd2_rn(z[2 * j4 + 1], make_float2(a.z, a.w));
                }
                const float x16 = z16 + add[64];
                float m0 = fmaxf(x[0].x, x[0].y), m1 = fmaxf(x[1].x, x[1].y), m2 = fmaxf(x[2].x, x[2].y), m3 = fmaxf(x[3].x, x[3].y);
                m0 = fmaxf(m0, fmaxf(x[4].x, x[4].y)); m1 = fmaxf(m1, fmaxf(x[5].x, x[5].y));
                m2 = fmaxf(m2, fmaxf(x[6].x, x[6].y)); m3 = fmaxf(m3, fmaxf(x[7].x, x[7].y));
                float m = fmaxf(fmaxf(m0, m1), fmaxf(fmaxf(m2, m3), x16));
                m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 1));
                m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 2));
                // nml = rn(-m log2 e) carries a rounding error eps = nml + m log2 e (up to 3e-5 for |m| ~ 500) that would scale
                // every term by 2^eps, i.e. shift the logsumexp by eps ln 2: the FMA recovers eps exactly and it is taken out
                // after the log (one FMA per logsumexp instead of a subtraction per term)
                const float nml = -m * L2E;
                const float eps = fmaf(m, L2E, nml);
                const float2 l2e2 = make_float2(L2E, L2E), nml2 = make_float2(nml, nml);
                float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float2 t01 = __ffma2_rn(x[j], l2e2, nml2), t23 = __ffma2_rn(x[j + 1], l2e2, nml2);
                    s01 = __fadd2_rn(s01, make_float2(fast_ex2(t01.x), fast_ex2(t01.y)));
                    s23 = __fadd2_rn(s23, make_float2(fast_ex2(t23.x), fast_ex2(t23.y)));
                }
                const float s0 = s01.x + fast_ex2(fmaf(x16, L2E, nml));
                float sm = (s0 + s01.y) + (s23.x + s23.y);
                sm += __shfl_xor_sync(FULL_MASK, sm, 1);
                sm += __shfl_xor_sync(FULL_MASK, sm, 2);
                if (q == 0) {
                    const float nv = marg[ri] - fmaf(__log2f(sm) - eps, LN2, m);
                    changed = nv != dst[ri];
                    dst[ri] = nv;
                }
            } else {
                const float x0 = z[0].x + add[lane], x1 = z[0].y + add[lane + 32], x2 = z16 + add[64];
                const float m = warp_max(fmaxf(fmaxf(x0, x1), x2));
                const float nml = -m * L2E;
                const float eps = fmaf(m, L2E, nml);
                const float sm = warp_sum(fast_ex2(fmaf(x0, L2E, nml)) + fast_ex2(fmaf(x1, L2E, nml)) + fast_ex2(fmaf(x2, L2E, nml)));
                if (lane == 0) {
                    const float nv = marg[64] - fmaf(__log2f(sm) - eps, LN2, m);
                    changed = nv != dst[64];
                    dst[64] = nv;
                }
            }
            return changed;
        };
        // The iteration is deterministic: once a full iteration leaves v unchanged BIT FOR BIT, (u, v) is a fixed point of the
        // fp32 map and every remaining iteration would reproduce it, so stopping there gives exactly the result of all
        // num_iter iterations (the reference always runs 100, modules.py:21-26). The test is one compare per row folded into
        // the barrier the sweep needs anyway.
        for (int it = 0; it < P.num_iter; ++it) {
            sweep(zr, zr16, v, log_mu, u);      // u = log_mu - LSE_j(Z + v)
            __syncthreads();
            const int changed = sweep(zc, zc16, u, log_nu, v);      // v = log_nu - LSE_i(Z + u)
            if (!__syncthreads_or(changed)) break;
        }
    }
    // ---- output (P,65,65) log-assignment; keep it in Z for the matching step ----
    float* out = P.scores + (size_t)p * FP1 * FP1;
    for (int e = tid; e < FP1 * FP1; e += FT) {
        const int r = e / FP1, c = e % FP1;
        const float val = Z[e] + u[r] + v[c] - s_norm;
        Z[e] = val;
        out[e] = val;
    }
    __syncthreads();
    // ---- FineMatching (modules.py:242-274): exp, top-k along rows and columns, threshold, mutual, validity ----
    for (int e = tid; e < FP * FP; e += FT) {
        const int r = e >> 6, c = e & 63;
        Z[r * FP1 + c] = expf(Z[r * FP1 + c]);
        rowf[e] = 0;
    }
    __syncthreads();
    // row-wise top-k: warp per row, each lane holds columns lane and lane+32; ties -> lower index
    for (int r = warp; r < FP; r += FW) {
        float a0 = Z[r * FP1 + lane], a1 = Z[r * FP1 + lane + 32];
        for (int t = 0; t < P.topk; ++t) {
            float bv = a0 >= a1 ? a0 : a1;
            int bi = a0 >= a1 ? lane : lane + 32;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bi == lane) a0 = -CUDART_INF_F;
            if (bi == lane + 32) a1 = -CUDART_INF_F;
            if (lane == 0 && bv > P.threshold) rowf[r * FP + bi] = 1;
        }
    }
    __syncthreads();
    unsigned char* fl = P.flags + (size_t)p * FP * FP;
    for (int c = warp; c < FP; c += FW) {
        float a0 = Z[lane * FP1 + c], a1 = Z[(lane + 32) * FP1 + c];
        unsigned colsel0 = 0, colsel1 = 0;  // whether my rows were selected by the column top-k
        for (int t = 0; t < P.topk; ++t) {
            float bv = a0 >= a1 ? a0 : a1;
            int bi = a0 >= a1 ? lane : lane + 32;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bi == lane) { a0 = -CUDART_INF_F; colsel0 = bv > P.threshold; }
            if (bi == lane + 32) { a1 = -CUDART_INF_F; colsel1 = bv > P.threshold; }
        }
        const bool r0 = rowf[lane * FP + c], r1 = rowf[(lane + 32) * FP + c];
        const bool m0 = (P.mutual ? (r0 && colsel0) : (r0 || colsel0)) && t_ok[lane] && s_ok[c];
        const bool m1 = (P.mutual ? (r1 && colsel1) : (r1 || colsel1)) && t_ok[lane + 32] && s_ok[c];
        fl[lane * FP + c] = m0;
        fl[(lane + 32) * FP + c] = m1;
    }
}

// gather the final correspondences from the compacted flat indices (modules.py:276-283); blockIdx.y = pair
struct FineGatherStrides { int Pmax, Mt, Ms, Nt1, Ns1; };     // Nt1 / Ns1: rows of the per-pair (padded) point arrays
__global__ void fine_gather_kernel(const int* __restrict__ flat, const int* __restrict__ count, int capacity,
                                   const float* __restrict__ scores, const int* __restrict__ corr_t,
                                   const int* __restrict__ corr_s, const int* __restrict__ tgt_knn,
                                   const int* __restrict__ src_knn, const float* __restrict__ tgt_pts,
                                   const float* __restrict__ src_pts, float* __restrict__ out_t,
                                   float* __restrict__ out_s, float* __restrict__ out_score, FineGatherStrides S) {
    {
        const size_t b = blockIdx.y;
        flat += b * capacity; count += b; scores += b * S.Pmax * FP1 * FP1; corr_t += b * S.Pmax; corr_s += b * S.Pmax;
        tgt_knn += b * S.Mt * FP; src_knn += b * S.Ms * FP; tgt_pts += b * S.Nt1 * 3; src_pts += b * S.Ns1 * 3;
        out_t += b * capacity * 3; out_s += b * capacity * 3; out_score += b * capacity;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(__ldg(count), capacity);
    if (i >= n) return;
    const int f = __ldg(flat + i);
    const int p = f >> 12, r = (f >> 6) & 63, c = f & 63;
    const int pt = __ldg(tgt_knn + (size_t)__ldg(corr_t + p) * FP + r);
    const int ps = __ldg(src_knn + (size_t)__ldg(corr_s + p) * FP + c);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        out_t[3 * (size_t)i + d] = __ldg(tgt_pts + 3 * (size_t)pt + d);
        out_s[3 * (size_t)i + d] = __ldg(src_pts + 3 * (size_t)ps + d);
    }
    out_score[i] = expf(__ldg(scores + (size_t)p * FP1 * FP1 + r * FP1 + c));
}

}  // namespace

extern "C" long long roitr_point_to_node_workspace_bytes(int B, int N, int M) {
    return (long long)B * N * 8 + (long long)B * M * 8;      // buckets (u64 per point) + start / cursor (int per node each)
}

extern "C" int roitr_point_to_node_batched(int B, int N, int M, int limit, const float* pts, const float* nodes, int* owner,
                                           float* dmin, int* count, void* workspace, int* knn_idx, unsigned char* knn_mask,
                                           unsigned char* node_mask, void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && N >= 1 && M >= 1 && limit >= 1 && limit <= PART_CAP, "point_to_node: bad sizes");
    ROITR_CHECK_ARG(pts && nodes && owner && dmin && count && workspace && knn_idx && knn_mask && node_mask, "point_to_node: null");
    ROITR_CHECK_ARG((size_t)M * 16 <= 200 * 1024, "point_to_node: too many nodes (%d)", M);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* bucket = (unsigned long long*)workspace;
    int* start = (int*)(bucket + (size_t)B * N);
    int* cursor = start + (size_t)B * M;
    ROITR_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)B * M, st));
    ROITR_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)B * M, st));
    const size_t smem = (size_t)M * 16;
    static size_t configured_dev[ROITR_MAX_DEVICES] = {};
    size_t& configured = configured_dev[roitr_cur_device()];
    if (smem > 48 * 1024 && smem > configured) {
        ROITR_CUDA(cudaFuncSetAttribute(point_owner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    point_owner_kernel<<<dim3(ceil_div(N, 256), B), 256, smem, st>>>(N, M, pts, nodes, owner, dmin, count);
    ROITR_CHECK_LAUNCH("point_owner_kernel");
    node_start_kernel<<<B, 1024, 0, st>>>(M, count, start);
    node_bucket_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(N, M, owner, dmin, start, cursor, bucket);
    node_knn_kernel<<<dim3(M, B), 256, 0, st>>>(N, M, limit, count, start, bucket, knn_idx, knn_mask, node_mask);
    ROITR_CHECK_LAUNCH("node_knn_kernel");
    return ROITR_OK;
}

extern "C" int roitr_compact_flags_batched(int B, long long n, const unsigned char* flags, int* chunk_scratch, int* out,
                                           int capacity, int* count, void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && n >= 0 && flags && chunk_scratch && out && count, "compact_flags: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunks = (int)ceil_div_ll(n, CMP_CHUNK);
    if (nchunks == 0) { ROITR_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * B, st)); return ROITR_OK; }
    compact_count_kernel<<<dim3(nchunks, B), 256, 0, st>>>(n, flags, chunk_scratch);
    compact_scan_kernel<<<B, 1024, 0, st>>>(nchunks, chunk_scratch, count);
    compact_write_kernel<<<dim3(nchunks, B), 256, 0, st>>>(n, flags, chunk_scratch, out, capacity);
    ROITR_CHECK_LAUNCH("compact_flags");
    return ROITR_OK;
}

extern "C" int roitr_compact_flags(long long n, const unsigned char* flags, int* chunk_scratch, int* out, int capacity,
                                   int* count, void* stream) {
    return roitr_compact_flags_batched(1, n, flags, chunk_scratch, out, capacity, count, stream);
}

extern "C" long long roitr_compact_scratch_ints(long long n) { return ceil_div_ll(n, CMP_CHUNK) + 1; }

extern "C" int roitr_coarse_matching_batched(int B, int Mr, int Ms, int C, int k, int dual, const float* ref_feats,
                                             const float* src_feats, const unsigned char* ref_mask,
                                             const unsigned char* src_mask, const float* xy, float* work, int* out_ref,
                                             int* out_src, float* out_score, int* out_count, void* stream) {
    // xy = ref_feats @ src_feats^T per pair (B, Mr, Ms), computed by the caller; work: B * (Mr*Ms + 2*(Mr+Ms)) floats
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && Mr >= 1 && Ms >= 1 && k >= 1 && k <= TOPK_MAX, "coarse_matching: bad sizes (k <= %d)", TOPK_MAX);
    ROITR_CHECK_ARG(ref_feats && src_feats && ref_mask && src_mask && xy && work && out_ref && out_src && out_score && out_count,
                    "coarse_matching: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float* S = work;
    float* r2 = S + (size_t)B * Mr * Ms;
    float* s2 = r2 + (size_t)B * Mr;
    float* rowsum = s2 + (size_t)B * Ms;
    float* colsum = rowsum + (size_t)B * Mr;
    row_sqnorm_kernel<<<ceil_div(B * Mr * 32, 256), 256, 0, st>>>(B * Mr, C, ref_feats, r2);
    row_sqnorm_kernel<<<ceil_div(B * Ms * 32, 256), 256, 0, st>>>(B * Ms, C, src_feats, s2);
    coarse_exp_kernel<<<dim3(Mr, B), 256, 0, st>>>(Mr, Ms, xy, r2, s2, ref_mask, src_mask, S, rowsum);
    col_sum_kernel<<<dim3(ceil_div(Ms, 128), B), 128, 0, st>>>(Mr, Ms, S, colsum);
    DualNorm dn;
    dn.rowsum = rowsum; dn.colsum = colsum; dn.rmask = ref_mask; dn.smask = src_mask; dn.Mr = Mr; dn.Ms = Ms; dn.dual = dual;
    flat_topk_kernel<<<B, 1024, 0, st>>>(Mr * Ms, k, S, Ms, out_ref, out_src, out_score, out_count, k, dn);
    ROITR_CHECK_LAUNCH("coarse_matching");
    return ROITR_OK;
}

extern "C" int roitr_fine_matching_batched(int B, int Pmax, int Mt, int Ms, int Nt, int Ns, int C, const float* tgt_feat,
                                           const float* src_feat, const int* tgt_knn, const int* src_knn,
                                           const unsigned char* tgt_kmask, const unsigned char* src_kmask, const int* corr_t,
                                           const int* corr_s, const int* corr_count, const float* alpha, int num_iter, int topk,
                                           int mutual, float threshold, float* scores, unsigned char* flags, void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && Pmax >= 1 && C % 32 == 0 && topk >= 1 && topk <= 32, "fine_matching: bad sizes");
    ROITR_CHECK_ARG(tgt_feat && src_feat && tgt_knn && src_knn && tgt_kmask && src_kmask && corr_t && corr_s &&
                    corr_count && alpha && scores && flags, "fine_matching: null pointer");
    ROITR_CHECK_ARG(((uintptr_t)tgt_feat | (uintptr_t)src_feat) % 16 == 0 && ((size_t)Nt * C) % 4 == 0 && ((size_t)Ns * C) % 4 == 0,
                    "fine_matching: alignment");
    FineParams P;
    P.tgt_feat = tgt_feat; P.src_feat = src_feat; P.Nt = Nt; P.Ns = Ns; P.C = C; P.tgt_knn = tgt_knn; P.src_knn = src_knn;
    P.tgt_kmask = tgt_kmask; P.src_kmask = src_kmask; P.corr_t = corr_t; P.corr_s = corr_s; P.corr_count = corr_count;
    P.alpha = alpha; P.scores = scores; P.flags = flags; P.num_iter = num_iter; P.topk = topk; P.mutual = mutual;
    P.threshold = threshold; P.sqrt_c = sqrtf((float)C); P.Pmax = Pmax; P.Mt = Mt; P.Ms = Ms;
    cudaStream_t st = (cudaStream_t)stream;
    ROITR_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * Pmax * FP * FP, st));
    fine_patch_kernel<<<dim3(Pmax, B), FT, 0, st>>>(P);
    ROITR_CHECK_LAUNCH("fine_patch_kernel");
    return ROITR_OK;
}

extern "C" int roitr_fine_gather_batched(int B, int capacity, int Pmax, int Mt, int Ms, int Nt1, int Ns1, const int* flat,
                                         const int* count, const float* scores, const int* corr_t, const int* corr_s,
                                         const int* tgt_knn, const int* src_knn, const float* tgt_pts_padded,
                                         const float* src_pts_padded, float* out_t, float* out_s, float* out_score,
                                         void* stream) {
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && capacity >= 0 && flat && count && scores && out_t && out_s && out_score, "fine_gather: bad arguments");
    if (capacity == 0) return ROITR_OK;
    FineGatherStrides S;
    S.Pmax = Pmax; S.Mt = Mt; S.Ms = Ms; S.Nt1 = Nt1; S.Ns1 = Ns1;
    fine_gather_kernel<<<dim3(ceil_div(capacity, 256), B), 256, 0, (cudaStream_t)stream>>>(
        flat, count, capacity, scores, corr_t, corr_s, tgt_knn, src_knn, tgt_pts_padded, src_pts_padded, out_t, out_s,
        out_score, S);
    ROITR_CHECK_LAUNCH("fine_gather_kernel");
    return ROITR_OK;
}

// ------------------------------------------------------------------------------------------------ adaptive coarse matching
// AdaptiveSuperPointMatching.forward (model/modules.py:81-123, 4DMatch head): sim = sqrt(clamp(2 - 2 a.b, 1e-12)) on the
// valid superpoints; every pair with sim <= threshold in row-major (torch.nonzero) order, or - when fewer than
// min(min_num, #valid pairs) qualify - the min_num smallest, ascending. Both candidate lists are produced on the device
// (ordered compaction + flat top-k) and a last kernel picks one from the device-side counts: no host synchronisation.
namespace {

__global__ void adaptive_sim_kernel(int Ma, int Mb, const float* __restrict__ xy, const unsigned char* __restrict__ amask,
                                    const unsigned char* __restrict__ bmask, float thr, float* __restrict__ sim,
                                    float* __restrict__ key, unsigned char* __restrict__ flag) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)Ma * Mb;
    if (e >= n) return;
    const size_t b = blockIdx.y;
    xy += b * n; sim += b * n; key += b * n; flag += b * n; amask += b * Ma; bmask += b * Mb;
    const int i = (int)(e / Mb), j = (int)(e % Mb);
    if (!(amask[i] && bmask[j])) { sim[e] = CUDART_INF_F; key[e] = -1.f; flag[e] = 0; return; }
    // square_distance(normalized=True): 2 - 2 xy, clamp 1e-12 (lib/utils.py:149,155); then sqrt (modules.py:101)
    const float d = __fsqrt_rn(fmaxf(__fsub_rn(2.0f, __fmul_rn(2.0f, __ldg(xy + e))), 1e-12f));
    sim[e] = d;
    key[e] = __uint_as_float(0x7f7fffffu - __float_as_uint(d));   // exact order-reversing map: largest key = smallest sim
    flag[e] = d <= thr;
}

__global__ void adaptive_select_kernel(int Ma, int Mb, int cap, int min_num, const int* __restrict__ count_c,
                                       const int* __restrict__ flat_c, const int* __restrict__ count_k,
                                       const int* __restrict__ row_k, const int* __restrict__ col_k,
                                       const float* __restrict__ sim, int* __restrict__ out_a, int* __restrict__ out_b,
                                       float* __restrict__ out_score, int* __restrict__ out_count) {
    {
        const size_t b = blockIdx.y;
        count_c += b; count_k += b; flat_c += b * cap; row_k += b * min_num; col_k += b * min_num;
        sim += b * Ma * Mb; out_a += b * cap; out_b += b * cap; out_score += b * cap; out_count += b;
    }
    const int nc = __ldg(count_c), nk = __ldg(count_k);
    const bool use_topk = nc < nk;                                  // masks.sum() < min_num (modules.py:105)
    const int n = use_topk ? nk : min(nc, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int a, b;
        if (use_topk) { a = __ldg(row_k + i); b = __ldg(col_k + i); }
        else { const int f = __ldg(flat_c + i); a = f / Mb; b = f % Mb; }
        out_a[i] = a; out_b[i] = b;
        out_score[i] = expf(-__ldg(sim + (size_t)a * Mb + b));      // corr_scores = exp(-corr_distances) (modules.py:112)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = n;
}

}  // namespace

extern "C" int roitr_coarse_matching_adaptive_batched(int B, int Ma, int Mb, int min_num, float threshold,
                                                      const unsigned char* a_mask, const unsigned char* b_mask, const float* xy,
                                                      float* work, int* iwork, int cap, int* out_a, int* out_b,
                                                      float* out_score, int* out_count, void* stream) {
    // xy = a_feats @ b_feats^T per pair (B, Ma, Mb). work: B * (2*Ma*Mb floats + Ma*Mb bytes rounded up to floats).
    // iwork: B * (cap + 3*min_num + 4 + roitr_compact_scratch_ints(Ma*Mb)) ints.
    ROITR_CHECK_ARG(B >= 1 && B <= 65535 && Ma >= 1 && Mb >= 1 && min_num >= 1 && min_num <= TOPK_MAX && cap >= min_num,
                    "coarse_matching_adaptive: bad sizes");
    ROITR_CHECK_ARG(a_mask && b_mask && xy && work && iwork && out_a && out_b && out_score && out_count, "coarse_matching_adaptive: null");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)Ma * Mb;
    const int nchunks = (int)ceil_div_ll(n, CMP_CHUNK);
    float* sim = work;
    float* key = sim + (size_t)B * n;
    unsigned char* flag = reinterpret_cast<unsigned char*>(key + (size_t)B * n);
    int* flat_c = iwork;
    int* row_k = flat_c + (size_t)B * cap;
    int* col_k = row_k + (size_t)B * min_num;
    float* val_k = reinterpret_cast<float*>(col_k + (size_t)B * min_num);
    int* count_c = reinterpret_cast<int*>(val_k + (size_t)B * min_num);
    int* count_k = count_c + B;
    int* scratch = count_k + B;
    adaptive_sim_kernel<<<dim3((unsigned)ceil_div_ll(n, 256), B), 256, 0, st>>>(Ma, Mb, xy, a_mask, b_mask, threshold, sim, key, flag);
    compact_count_kernel<<<dim3(nchunks, B), 256, 0, st>>>(n, flag, scratch);
    compact_scan_kernel<<<B, 1024, 0, st>>>(nchunks, scratch, count_c);
    compact_write_kernel<<<dim3(nchunks, B), 256, 0, st>>>(n, flag, scratch, flat_c, cap);
    DualNorm none;
    none.rowsum = nullptr; none.colsum = nullptr; none.rmask = nullptr; none.smask = nullptr; none.Mr = 0; none.Ms = 0; none.dual = 0;
    flat_topk_kernel<<<B, 1024, 0, st>>>((int)n, min_num, key, Mb, row_k, col_k, val_k, count_k, min_num, none);
    adaptive_select_kernel<<<dim3(ceil_div(cap, 256) > 1024 ? 1024 : ceil_div(cap, 256), B), 256, 0, st>>>(
        Ma, Mb, cap, min_num, count_c, flat_c, count_k, row_k, col_k, sim, out_a, out_b, out_score, out_count);
    ROITR_CHECK_LAUNCH("coarse_matching_adaptive");
    return ROITR_OK;
}
