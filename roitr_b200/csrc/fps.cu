// Furthest point sampling on a thread-block cluster (sm_100a).
//
// Replaces furthestsampling_cuda_kernel (cpp_wrappers/pointops/src/sampling/sampling_cuda_kernel.cu:14-129).
//
// The reference runs ONE block per cloud (grid = b = 1 in RoITr): every iteration re-reads all points and the running
// distance array `tmp` from global memory and finishes with an 11-level shared-memory tree (11 __syncthreads). FPS is a
// chain of m dependent iterations, so what matters is the latency of one iteration.
//
// Here a cloud is owned by a CLUSTER of CL CTAs. Coordinates, running min-distances and tie-break keys live in REGISTERS
// for the whole kernel (PPT points per thread; a second copy of the CTA's coordinates sits in shared memory only so the
// winner's xyz can be looked up by index). One iteration:
//   1. per thread: PPT independent fused distance updates + an FMNMX tree                       (no select chains)
//   2. per warp: redux.sync.max of the float bits; the lanes that hold the maximum resolve the tie-break key; redux.min
//   3. one __syncthreads; every warp re-reduces the <=16 warp candidates redundantly (no second barrier)
//   4. cluster exchange WITHOUT a cluster barrier: warp 0 pushes the CTA candidate (dist, key, xyz = 20 B) into every
//      peer's shared memory with st.async, which also signals the peer's mbarrier (complete_tx); each CTA waits on its
//      own mbarrier (try_wait), then picks the winner with two more redux. No fence, no barrier.cluster in the loop.
//      (ncu on the first version, which used cooperative_groups cluster.sync(): ERRBAR + UCGABAR_ARV/WAIT + membar were
//      ~50 % of all stall samples; profiles/r01_fps_v1_stalls.txt.)
//
// Bit-exactness with the reference, including exact-distance ties: the distance is the reference's FMA association
// (sqdist_ref) with d = x_k - x_last; the reference's winner among equal maxima is the lowest k inside a thread
// (strict '>' scanning k ascending, .cu:57-58) and then the lower SLOT at every level of its power-of-two tree
// ('v2 > v1 ? i2 : i1', .cu:5-10), i.e. the smallest bit-reversed thread id. Both are encoded in a 31-bit key
//   key(k) = bitrev_{log2 BS}((k-start) % BS) << 21 | (k-start) / BS ,   BS = reference block size for n_max
// and the winner is (max distance, min key); k is recovered from the key.
#include <cooperative_groups.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int MAX_CL = 8;
constexpr unsigned NO_KEY = 0xffffffffu;
constexpr int XCHG_BYTES = 20;  // bits, key, x, y, z

struct __align__(16) Cand {
    int bits;       // float bits of the running distance; signed compare (valid distances >= 0, the -1 sentinel is < 0)
    unsigned key;   // tie-break key, smaller wins
    float x, y, z;
    float pad[3];
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

__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// 16-byte + 4-byte remote stores that also complete `bytes` on the remote mbarrier (async proxy; no fence needed:
// the mbarrier phase completion orders the data for the waiter).
__device__ __forceinline__ void st_async_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t dst, uint32_t a, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(a), "r"(bar)
                 : "memory");
}

template <int N>
__device__ __forceinline__ float tree_max(const float (&v)[N]) {
    float t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = v[i];
#pragma unroll
    for (int s = 1; s < N; s <<= 1)
#pragma unroll
        for (int i = 0; i + s < N; i += 2 * s) t[i] = fmaxf(t[i], t[i + s]);
    return t[0];
}

// (A 64-register build, two CTAs per SM, was measured: the spills land in the iteration loop, 4.3 -> 5.3 ms per step.)
template <int CL, int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_cluster_kernel(const FpsParams P) {
    extern __shared__ __align__(16) float s_pts[];  // [PPT * FPS_THREADS][3]: this CTA's coordinates by local slot
    __shared__ Cand s_cta[2][MAX_CL];
    __shared__ int2 s_warp[2][FPS_WARPS];
    __shared__ __align__(8) uint64_t s_bar[2];

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (CL > 1) ? (int)cluster.block_rank() : 0;
    const int cloud = blockIdx.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

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
            pd[i] = -1.f;  // min(d, -1) stays -1: never the maximum
            pk[i] = NO_KEY;
        }
        float* sp = s_pts + 3 * (size_t)(i * FPS_THREADS + tid);
        sp[0] = px[i]; sp[1] = py[i]; sp[2] = pz[i];
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
    if (CL > 1) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
        }
        cluster.sync();  // peers' mbarriers are initialised and all CTAs are resident before any DSMEM traffic
    } else {
        __syncthreads();
    }

    // recover (relative index, local slot) from a key: key = bitrev(t) << 21 | q with r = q * BS + t
    auto key_to_rel = [&](unsigned key) {
        const unsigned rev = key >> 21, qd = key & 0x1fffffu;
        const unsigned t_ref = bs_log2 ? (__brev(rev) >> (32 - bs_log2)) : 0u;
        return (int)(qd << bs_log2) + (int)t_ref;
    };

    int it = 0;
    for (int j = start_m + 1; j < end_m; ++j, ++it) {
        const int par = it & 1;
        if (CL > 1 && tid == 0) mbar_expect_tx(&s_bar[par], CL * XCHG_BYTES);  // arm this iteration's exchange
        // ---- 1. per-thread update + max ----
#pragma unroll
        for (int i = 0; i < PPT; ++i) pd[i] = fminf(sqdist_ref(px[i] - lx, py[i] - ly, pz[i] - lz), pd[i]);
        const int tb = __float_as_int(tree_max<PPT>(pd));
        // ---- 2. warp candidate ----
        const int wmax = __reduce_max_sync(FULL_MASK, tb);
        unsigned key = NO_KEY;
        if (tb == wmax) {
#pragma unroll
            for (int i = 0; i < PPT; ++i) key = (__float_as_int(pd[i]) == wmax) ? min(key, pk[i]) : key;
        }
        const unsigned wkey = __reduce_min_sync(FULL_MASK, key);
        if (lane == 0) s_warp[par][warp] = make_int2(wmax, (int)wkey);
        __syncthreads();
        // ---- 3. CTA candidate (every warp, redundantly) ----
        int cb = (int)0x80000000;
        unsigned ck = NO_KEY;
        if (lane < FPS_WARPS) { const int2 w = s_warp[par][lane]; cb = w.x; ck = (unsigned)w.y; }
        const int cmax = __reduce_max_sync(FULL_MASK, cb);
        const unsigned ckey = __reduce_min_sync(FULL_MASK, cb == cmax ? ck : NO_KEY);

        unsigned fkey;
        if (CL == 1) {
            fkey = ckey;
            if (ckey != NO_KEY) {
                const int rel = key_to_rel(ckey);  // CL == 1: local slot = (rel / 512) * 512 + rel % 512 = rel
                lx = s_pts[3 * rel]; ly = s_pts[3 * rel + 1]; lz = s_pts[3 * rel + 2];
            }
        } else {
            // ---- 4. cluster exchange: st.async + mbarrier ----
            if (warp == 0 && lane < CL) {
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (ckey != NO_KEY) {
                    const int rel = key_to_rel(ckey);
                    const int slot = (rel / (CL * FPS_THREADS)) * FPS_THREADS + (rel % FPS_THREADS);
                    cx = s_pts[3 * slot]; cy = s_pts[3 * slot + 1]; cz = s_pts[3 * slot + 2];
                }
                const uint32_t dst = map_to_rank(smem_u32(&s_cta[par][rank]), (uint32_t)lane);
                const uint32_t rbar = map_to_rank(smem_u32(&s_bar[par]), (uint32_t)lane);
                st_async_v4(dst, (uint32_t)cmax, ckey, __float_as_uint(cx), __float_as_uint(cy), rbar);
                st_async_b32(dst + 16, __float_as_uint(cz), rbar);
            }
            mbar_wait(&s_bar[par], (it >> 1) & 1);
            int rb = (int)0x80000000;
            unsigned rk = NO_KEY;
            if (lane < CL) { rb = s_cta[par][lane].bits; rk = s_cta[par][lane].key; }
            const int fmax = __reduce_max_sync(FULL_MASK, rb);
            fkey = __reduce_min_sync(FULL_MASK, rb == fmax ? rk : NO_KEY);
            const int src = __ffs(__ballot_sync(FULL_MASK, lane < CL && rb == fmax && rk == fkey)) - 1;
            const Cand& w = s_cta[par][src < 0 ? 0 : src];
            lx = w.x; ly = w.y; lz = w.z;
        }
        if (rank == 0 && tid == 0) {
            P.idx[j] = start_n + (fkey != NO_KEY ? key_to_rel(fkey) : 0);
            if (P.new_xyz) {
                P.new_xyz[3 * (size_t)j] = lx; P.new_xyz[3 * (size_t)j + 1] = ly; P.new_xyz[3 * (size_t)j + 2] = lz;
            }
        }
    }
    if (CL > 1) cluster.sync();  // no CTA exits while a peer may still write into its shared memory
}

template <int CL, int PPT>
int launch_fps(const FpsParams& P, cudaStream_t st) {
    const int smem = PPT * FPS_THREADS * 12;
    static bool attr_set_dev[ROITR_MAX_DEVICES] = {};
    bool& attr_set = attr_set_dev[roitr_cur_device()];
    if (!attr_set) {  // dynamic + static shared memory can exceed the 48 KB default from PPT = 8 on
        ROITR_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<CL, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.b * CL);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem;
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
    // capacity needed: CL * 512 * PPT >= n_seg_max. Latency mode (few clouds) prefers big clusters / few points per
    // thread, throughput mode (many clouds) prefers small clusters.
    int cl = cluster_hint;
    if (cl != 1 && cl != 2 && cl != 4 && cl != 8) {
        if (b >= 8) {
            cl = 1;      // many clouds: the smallest cluster that holds a segment (raised below) - measured on B200 with 32 clouds:
                         // 5000 points 0.57 ms on 1 CTA vs 0.64 on 4, and 32 instead of 128 SMs held (699 vs 688 pairs/s)
        } else {
            cl = 8;      // few clouds: latency mode, the largest cluster that still fits the GPU
            while (cl > 1 && b * cl > 148) cl >>= 1;
        }
    }
    while (cl < MAX_CL && (long long)cl * FPS_THREADS * 16 < n_seg_max) cl <<= 1;
    const int per_thread = ceil_div(n_seg_max, cl * FPS_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
#define FPS_DISPATCH(CLV)                                                  \
    if (cl == CLV) {                                                       \
        if (per_thread <= 2) return launch_fps<CLV, 2>(P, st);             \
        if (per_thread <= 3) return launch_fps<CLV, 3>(P, st);             \
        if (per_thread <= 4) return launch_fps<CLV, 4>(P, st);             \
        if (per_thread <= 5) return launch_fps<CLV, 5>(P, st);             \
        if (per_thread <= 6) return launch_fps<CLV, 6>(P, st);             \
        if (per_thread <= 8) return launch_fps<CLV, 8>(P, st);             \
        if (per_thread <= 10) return launch_fps<CLV, 10>(P, st);           \
        if (per_thread <= 12) return launch_fps<CLV, 12>(P, st);           \
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
