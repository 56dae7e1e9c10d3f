// Uniform-grid acceleration structure for the exact kNN (sm_100a). Included by knn_ppf.cu.
//
// Brute force evaluates m*n distances (4e8 for the 20k x 20k call) and is FP32-issue bound. The grid keeps the result
// EXACTLY the same - the same sqdist_ref arithmetic on every candidate, the same (distance, index) order, the same tie
// replay - but only visits the cells of a growing cube around the query until the k-th distance is provably smaller than
// the distance to anything outside the cube.
//
// Build (per reference set, all segments of the batch at once):  bbox -> cell size / dims -> count -> scan -> scatter.
// Points are stored cell-sorted as float4 (x, y, z, index bits) so a query streams candidates with coalesced 16-byte loads.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace knngrid {

constexpr int MAX_DIM = 40;                       // cells per axis (<= 64000 cells per segment)
constexpr int MAX_CELLS = MAX_DIM * MAX_DIM * MAX_DIM;
// target_per_cell (kernel argument): average occupancy over the bounding box; clustered data has far more in dense cells

struct SegHeader {                                // one per segment (cloud)
    float ox, oy, oz, h, inv_h;
    int nx, ny, nz;
    int cell_base;                                // offset of this segment's cell_start[] (ncell + 1 entries)
    int pad[3];
};

// ---- build ----------------------------------------------------------------------------------------------------------
__global__ void grid_header_kernel(int b, const float* __restrict__ xyz, const int* __restrict__ offset,
                                   SegHeader* __restrict__ hdr, float target_per_cell) {
    const int s = blockIdx.x;
    const int start = s == 0 ? 0 : __ldg(offset + s - 1), end = __ldg(offset + s);
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = start + threadIdx.x; i < end; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = __ldg(xyz + 3 * (size_t)i + a); mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    __shared__ float smn[3][32], smx[3][32];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = mn[a], hi = mx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(FULL_MASK, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(FULL_MASK, hi, o)); }
        if ((threadIdx.x & 31) == 0) { smn[a][threadIdx.x >> 5] = lo; smx[a][threadIdx.x >> 5] = hi; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5, n = end - start;
        float lo[3], ext[3];
        for (int a = 0; a < 3; ++a) {
            float l = CUDART_INF_F, h = -CUDART_INF_F;
            for (int w = 0; w < nw; ++w) { l = fminf(l, smn[a][w]); h = fmaxf(h, smx[a][w]); }
            if (n == 0) { l = 0.f; h = 0.f; }
            lo[a] = l; ext[a] = fmaxf(h - l, 1e-6f);
        }
        float h = cbrtf(ext[0] * ext[1] * ext[2] * target_per_cell / fmaxf((float)n, 1.f));
        h = fmaxf(h, fmaxf(ext[0], fmaxf(ext[1], ext[2])) / (float)MAX_DIM * 1.0001f);   // respect the dimension cap
        SegHeader H;
        H.ox = lo[0]; H.oy = lo[1]; H.oz = lo[2]; H.h = h; H.inv_h = 1.0f / h;
        H.nx = min(MAX_DIM, (int)(ext[0] / h) + 1); H.ny = min(MAX_DIM, (int)(ext[1] / h) + 1); H.nz = min(MAX_DIM, (int)(ext[2] / h) + 1);
        H.cell_base = s * (MAX_CELLS + 1);
        H.pad[0] = H.pad[1] = H.pad[2] = 0;
        hdr[s] = H;
    }
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int n) {
    const int c = (int)floorf((v - o) * inv_h);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
__device__ __forceinline__ int seg_of(int i, const int* __restrict__ ends, int b) {
    int lo = 0, hi = b - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (i < __ldg(ends + mid)) hi = mid; else lo = mid + 1; }
    return lo;
}

// pass 0: count per cell; pass 1 (after the scan): scatter into cell-sorted order
__global__ void grid_bin_kernel(int n, int b, const float* __restrict__ xyz, const int* __restrict__ offset,
                                const SegHeader* __restrict__ hdr, int* __restrict__ cell_start,
                                int* __restrict__ cursor, float4* __restrict__ sorted, int pass) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = seg_of(i, offset, b);
    const SegHeader H = hdr[s];
    const float x = __ldg(xyz + 3 * (size_t)i), y = __ldg(xyz + 3 * (size_t)i + 1), z = __ldg(xyz + 3 * (size_t)i + 2);
    const int c = (cell_coord(z, H.oz, H.inv_h, H.nz) * H.ny + cell_coord(y, H.oy, H.inv_h, H.ny)) * H.nx + cell_coord(x, H.ox, H.inv_h, H.nx);
    if (pass == 0) {
        atomicAdd(cursor + H.cell_base + c, 1);
    } else {
        const int seg_start = s == 0 ? 0 : __ldg(offset + s - 1);
        const int pos = seg_start + __ldg(cell_start + H.cell_base + c) + atomicAdd(cursor + H.cell_base + c, 1);
        sorted[pos] = make_float4(x, y, z, __int_as_float(i));
    }
}

// exclusive scan of the per-cell counts of each segment (cursor -> cell_start, cursor zeroed for the scatter pass): one CTA
// per segment, 8 consecutive cells per thread (serial), shuffle scan of the thread sums, one shared-memory pass over the 32
// warp totals: 3 barriers per 8192 cells (the first version ran a 20-barrier Hillis-Steele pass per 1024 cells: 51 us for
// the 40 000 cells of a 20 000-point cloud, on the search lane every kNN of the step waits for).
__global__ void __launch_bounds__(1024) grid_scan_kernel(const SegHeader* __restrict__ hdr, int* __restrict__ cursor, int* __restrict__ cell_start) {
    constexpr int IPT = 8;
    const SegHeader H = hdr[blockIdx.x];
    const int ncell = H.nx * H.ny * H.nz;
    __shared__ int wsum[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < ncell; base += 1024 * IPT) {
        const int i0 = base + tid * IPT;
        int v[IPT], sum = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            v[j] = (i0 + j < ncell) ? cursor[H.cell_base + i0 + j] : 0;
            sum += v[j];
        }
        int inc = sum;                                   // inclusive scan of the thread sums within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;                       // exclusive offset of each warp
            if (lane == 31) s_total = winc;
        }
        __syncthreads();
        int run = carry + wsum[warp] + inc - sum;        // exclusive prefix of this thread's first cell
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            if (i0 + j < ncell) { cell_start[H.cell_base + i0 + j] = run; cursor[H.cell_base + i0 + j] = 0; }
            run += v[j];
        }
        carry += s_total;
        __syncthreads();                                 // wsum / s_total are rewritten by the next pass
    }
    if (tid == 0) cell_start[H.cell_base + ncell] = carry;
}

}  // namespace knngrid
