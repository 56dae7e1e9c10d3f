// Shared device/host helpers for libroitr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ROITR_OK 0
#define ROITR_ERR_ARG (-1)
#define ROITR_ERR_UNSUPPORTED (-2)

void roitr_set_error(const char* fmt, ...);

#define ROITR_CHECK_ARG(cond, ...)            \
    do {                                      \
        if (!(cond)) {                        \
            roitr_set_error(__VA_ARGS__);     \
            return ROITR_ERR_ARG;             \
        }                                     \
    } while (0)

// Launch-error check: no device sync (async errors surface at the caller's next sync, like any CUDA library).
#define ROITR_CHECK_LAUNCH(name)                                                         \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            roitr_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

#define ROITR_CUDA(call)                                                                 \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            roitr_set_error("%s: %s", #call, cudaGetErrorString(e__));                   \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

// Launch-side caches (cudaFuncSetAttribute done, SM count, occupancy) are per DEVICE: a process may drive several GPUs.
#define ROITR_MAX_DEVICES 64
static inline int roitr_cur_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d < 0 ? 0 : d) % ROITR_MAX_DEVICES;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
#define FULL_MASK 0xffffffffu

// Squared distance with the exact association the reference kernels compile to (nvcc -O2, --fmad=true):
//   FMUL t = dy*dy ; FFMA t = dx*dx + t ; FFMA t = dz*dz + t
// (SASS of knnquery_cuda_kernel.cu:96 and sampling_cuda_kernel.cu:55, see oracle/pointops_ref.c).
// Written with explicit intrinsics so that no compiler flag or surrounding code can re-associate it.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; SASS: UBLKCP). dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif
