// Furthest point sampling on a thread-block cluster (sm_100a).
//
// Replaces furthestsampling_cuda_kernel (cpp_wrappers/pointops/src/sampling/sampling_cuda_kernel.cu:14-129).
//
// The reference runs ONE block per cloud (grid = b = 1 in RoITr): every iteration re-reads all points and the running
// distance array `tmp` from global memory and finishes with an 11-level shared-memory tree (11 __syncthreads). FPS is a
// chain of m dependent iterations, so what matters is the latency of one iteration.
//
// Here a cloud is owned by a CLUSTER of CL CTAs. Points, running min-distances and tie-break keys live in REGISTERS for
// the whole kernel (PPT points per thread); an iteration is
//   per-thread scan (PPT fused distance updates) -> 2x redux.sync per warp (max of the float bits, min of the key)
//   -> one __syncthreads -> every warp re-reduces the <=16 warp candidates -> one DSMEM all-to-all of the CTA candidates
//   + one cluster barrier -> every thread picks the winner (with its xyz carried along, so no global read on the chain).
//
// Bit-exactness with the reference, including exact-distance ties: the distance is the reference's FMA association
// (sqdist_ref) with d = x_k - x_last; the reference's winner among equal maxima is the lowest k inside a thread
// (strict '>' scanning k ascending, .cu:57-58) and then the lower SLOT at every level of its power-of-two tree
// ('v2 > v1 ? i2 : i1', .cu:5-10), i.e. the smallest bit-reversed thread id. Both are encoded in a 31-bit key
//   key(k) = bitrev_{log2 BS}((k-start) % BS) << 21 | (k-start) / BS ,   BS = reference block size for n_max
// and the winner is (max distance, min key). k is recovered from the key, so candidates are (dist, key, x, y, z).
#include <cooperative_groups.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int MAX_CL = 8;

struct __align__(16) Cand {
    unsigned bits;  // float bits of the running distance (>= 0, so unsigned order == float order)
    unsigned key;   // tie-break key, smaller wins
    float x, y, z;
    float pad0, pad1, pad2;
};

struct FpsParams {
    const float* xyz;
    const int* offset;
    const int* new_offset;
    int* idx;
    float* new_xyz;
    int b;
    int bs_shared_log2;  // >= 0: log2 of the batch-wide reference block size; -1: per-segment
};

// src/cuda_utils.h:11-14: min(2^(int)(log(n)/log(2)), 1024). The host's double-precision log ratio equals the integer
// floor(log2 n) for every n in [1, 70000) (checked against glibc in tests/test_oracle_native.py), so use the integer form.
__host__ __device__ __forceinline__ int ref_block_log2(int n) {
    int p = 0;
    while ((2 << p) <= n && p < 10) ++p;
    return p;
}

__device__ __forceinline__ bool better(unsigned b1, unsigned k1, unsigned b2, unsigned k2) {
    return (b1 > b2) || (b1 == b2 && k1 < k2);
}

template <int CL, int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_cluster_kernel(const FpsParams P) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (CL > 1) ? (int)cluster.block_rank() : 0;
    const int cloud = blockIdx.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    __shared__ Cand s_warp[2][FPS_WARPS];
    __shared__ Cand s_cta[2][MAX_CL];

    const int start_n = cloud == 0 ? 0 : __ldg(P.offset + cloud - 1);
    const int end_n = __ldg(P.offset + cloud);
    const int start_m = cloud == 0 ? 0 : __ldg(P.new_offset + cloud - 1);
    const int end_m = __ldg(P.new_offset + cloud);
    const int n = end_n - start_n;
    const int bs_log2 = P.bs_shared_log2 >= 0 ? P.bs_shared_log2 : ref_block_log2(n);
    const int bs_mask = (1 << bs_log2) - 1;

    // ---- load this thread's points (round-robin over the whole cluster: coalesced) ----
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
    unsigned pk[PPT];
    const int g = rank * FPS_THREADS + tid;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int r = g + i * (CL * FPS_THREADS);  // index relative to the segment
        if (r < n) {
            const float* p = P.xyz + 3 * (size_t)(start_n + r);
            px[i] = __ldg(p); py[i] = __ldg(p + 1); pz[i] = __ldg(p + 2);
            pd[i] = 1e10f;  // pointops.py:22
            const unsigned t_ref = (unsigned)(r & bs_mask);
            const unsigned rev = bs_log2 ? (__brev(t_ref) >> (32 - bs_log2)) : 0u;
            pk[i] = (rev << 21) | (unsigned)(r >> bs_log2);
        } else {
            px[i] = py[i] = pz[i] = 0.f;
            pd[i] = -1.f;       // never wins: a real point has distance >= 0 (float bits of -1 are handled below)
            pk[i] = 0xffffffffu;
        }
    }

    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (n > 0) {
        lx = __ldg(P.xyz + 3 * (size_t)start_n);
        ly = __ldg(P.xyz + 3 * (size_t)start_n + 1);
        lz = __ldg(P.xyz + 3 * (size_t)start_n + 2);
    }
    if (rank == 0 && tid == 0 && end_m > start_m) {
        P.idx[start_m] = start_n;  // sampling_cuda_kernel.cu:39
        if (P.new_xyz) {
            P.new_xyz[3 * (size_t)start_m] = lx; P.new_xyz[3 * (size_t)start_m + 1] = ly;
            P.new_xyz[3 * (size_t)start_m + 2] = lz;
        }
    }
    if (CL > 1) cluster.sync();  // all CTAs of the cluster are resident before any DSMEM traffic

    for (int j = start_m + 1; j < end_m; ++j) {
        const int par = j & 1;
        // ---- per-thread scan: update running distances, keep (max dist, min key) with its coordinates ----
        unsigned bb = 0u, bk = 0xffffffffu;
        float bx = 0.f, by = 0.f, bz = 0.f;
        bool any = false;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const float d = sqdist_ref(px[i] - lx, py[i] - ly, pz[i] - lz);
            const float d2 = fminf(d, pd[i]);
            const bool valid = pk[i] != 0xffffffffu;
            pd[i] = valid ? d2 : pd[i];
            const unsigned bits = __float_as_uint(d2);
            const bool take = valid && (!any || better(bits, pk[i], bb, bk));
            bb = take ? bits : bb; bk = take ? pk[i] : bk;
            bx = take ? px[i] : bx; by = take ? py[i] : by; bz = take ? pz[i] : bz;
            any = any || valid;
        }
        // ---- warp: two redux.sync ----
        const unsigned wmax = __reduce_max_sync(FULL_MASK, any ? bb : 0u);
        const unsigned wkey = __reduce_min_sync(FULL_MASK, (any && bb == wmax) ? bk : 0xffffffffu);
        if (any && bb == wmax && bk == wkey) {
            Cand c; c.bits = bb; c.key = bk; c.x = bx; c.y = by; c.z = bz; c.pad0 = c.pad1 = c.pad2 = 0.f;
            s_warp[par][warp] = c;
        } else if (wkey == 0xffffffffu && lane == 0) {
            Cand c; c.bits = 0u; c.key = 0xffffffffu; c.x = c.y = c.z = 0.f; c.pad0 = c.pad1 = c.pad2 = 0.f;
            s_warp[par][warp] = c;  // warp with no valid point
        }
        __syncthreads();
        // ---- CTA: every warp reduces the FPS_WARPS candidates redundantly (no second barrier) ----
        unsigned cb = 0u, ck = 0xffffffffu;
        if (lane < FPS_WARPS) { cb = s_warp[par][lane].bits; ck = s_warp[par][lane].key; }
        const unsigned cmax = __reduce_max_sync(FULL_MASK, cb);
        const unsigned ckey = __reduce_min_sync(FULL_MASK, (cb == cmax) ? ck : 0xffffffffu);
        const unsigned wl = __ballot_sync(FULL_MASK, lane < FPS_WARPS && cb == cmax && ck == ckey);
        const int wsrc = __ffs(wl) - 1;  // >= 0: at least one warp holds a valid point or all are sentinels

        unsigned fbits, fkey;
        if (CL == 1) {
            const Cand w = s_warp[par][wsrc < 0 ? 0 : wsrc];
            fbits = w.bits; fkey = w.key; lx = w.x; ly = w.y; lz = w.z;
        } else {
            if (warp == 0 && lane < CL) {
                // all-to-all: lane r writes this CTA's candidate into CTA r's slot [par][rank]
                const Cand w = s_warp[par][wsrc < 0 ? 0 : wsrc];
                Cand* remote = cluster.map_shared_rank(&s_cta[par][rank], lane);
                *remote = w;
            }
            cluster.sync();
            unsigned b0 = 0u, k0 = 0xffffffffu;
            int best = 0;
#pragma unroll
            for (int r = 0; r < CL; ++r) {
                const unsigned rb = s_cta[par][r].bits, rk = s_cta[par][r].key;
                if (r == 0 || better(rb, rk, b0, k0)) { b0 = rb; k0 = rk; best = r; }
            }
            fbits = b0; fkey = k0;
            lx = s_cta[par][best].x; ly = s_cta[par][best].y; lz = s_cta[par][best].z;
        }
        (void)fbits;
        if (rank == 0 && tid == 0) {
            // recover k from the key: key = bitrev(t) << 21 | q  with k - start = q * BS + t
            const unsigned rev = fkey >> 21, qd = fkey & 0x1fffffu;
            const unsigned t_ref = bs_log2 ? (__brev(rev) >> (32 - bs_log2)) : 0u;
            P.idx[j] = start_n + (int)(qd << bs_log2) + (int)t_ref;
            if (P.new_xyz) {
                P.new_xyz[3 * (size_t)j] = lx; P.new_xyz[3 * (size_t)j + 1] = ly; P.new_xyz[3 * (size_t)j + 2] = lz;
            }
        }
    }
    if (CL > 1) cluster.sync();  // no CTA exits while a peer may still write into its shared memory
}

template <int CL, int PPT>
int launch_fps(const FpsParams& P, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.b * CL);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ROITR_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CL, PPT>, P));
    return ROITR_OK;
}

}  // namespace

extern "C" int roitr_furthestsampling_cfg(int b, int n_max, int n_seg_max, const float* xyz, const int* offset,
                                          const int* new_offset, int* idx, float* new_xyz, int cluster_hint,
                                          void* stream) {
    ROITR_CHECK_ARG(b >= 1 && xyz && offset && new_offset && idx, "furthestsampling: bad arguments");
    ROITR_CHECK_ARG(n_seg_max >= 1 && n_seg_max <= MAX_CL * FPS_THREADS * 16,
                    "furthestsampling: segment of %d points exceeds the supported %d", n_seg_max,
                    MAX_CL * FPS_THREADS * 16);
    FpsParams P;
    P.xyz = xyz; P.offset = offset; P.new_offset = new_offset; P.idx = idx; P.new_xyz = new_xyz; P.b = b;
    P.bs_shared_log2 = n_max > 0 ? ref_block_log2(n_max) : -1;
    // capacity needed: CL * 512 * PPT >= n_seg_max. Latency mode (few clouds) prefers big clusters / small PPT,
    // throughput mode (many clouds) prefers CL=1. Auto: keep roughly <= 148 CTAs in flight.
    int cl = cluster_hint;
    if (cl != 1 && cl != 2 && cl != 4 && cl != 8) {
        cl = 8;
        while (cl > 1 && b * cl > 148) cl >>= 1;
    }
    while (cl < MAX_CL && (long long)cl * FPS_THREADS * 16 < n_seg_max) cl <<= 1;
    const int per_thread = ceil_div(n_seg_max, cl * FPS_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
#define FPS_DISPATCH(CLV)                                                  \
    if (cl == CLV) {                                                       \
        if (per_thread <= 2) return launch_fps<CLV, 2>(P, st);             \
        if (per_thread <= 4) return launch_fps<CLV, 4>(P, st);             \
        if (per_thread <= 8) return launch_fps<CLV, 8>(P, st);             \
        return launch_fps<CLV, 16>(P, st);                                 \
    }
    FPS_DISPATCH(1)
    FPS_DISPATCH(2)
    FPS_DISPATCH(4)
    FPS_DISPATCH(8)
#undef FPS_DISPATCH
    roitr_set_error("furthestsampling: no kernel for cluster=%d", cl);
    return ROITR_ERR_UNSUPPORTED;
}

extern "C" int roitr_furthestsampling(int b, int n_max, const float* xyz, const int* offset, const int* new_offset,
                                      float* tmp, int* idx, float* new_xyz, int cluster_hint, void* stream) {
    (void)tmp;
    // Reference ABI: segment lengths are only known on the device. n_max (> 0) bounds every segment, as in the
    // reference launcher; with n_max == 0 read the offsets back (b ints, one sync) to size the launch.
    int n_seg_max = n_max;
    if (n_max <= 0) {
        ROITR_CHECK_ARG(b <= 4096, "furthestsampling: b too large");
        int ends[4096];
        ROITR_CUDA(cudaMemcpyAsync(ends, offset, sizeof(int) * b, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        ROITR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        n_seg_max = 0;
        for (int i = 0, prev = 0; i < b; ++i) { n_seg_max = ends[i] - prev > n_seg_max ? ends[i] - prev : n_seg_max; prev = ends[i]; }
    }
    return roitr_furthestsampling_cfg(b, n_max, n_seg_max, xyz, offset, new_offset, idx, new_xyz, cluster_hint, stream);
}
