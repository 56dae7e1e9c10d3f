/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or called by the product path
 * (roitr_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * CPU restatement of the reference's two native hot-path kernels, following the reference
 * algorithm step by step so that results are bit-identical to what the CUDA kernels produce:
 *
 *   oracle_knnquery          <- cpp_wrappers/pointops/src/knnquery/knnquery_cuda_kernel.cu:21-108
 *   oracle_furthestsampling  <- cpp_wrappers/pointops/src/sampling/sampling_cuda_kernel.cu:14-129
 *                               (+ src/cuda_utils.h:11-14 for the block size, which fixes tie order)
 *
 * Arithmetic pinned from the SASS of the reference kernels compiled with nvcc 12.9 -O2 for
 * sm_100a (oracle/build_ref.sh; default --fmad=true). For
 *     (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)
 * ptxas emits  FMUL t = dy*dy ; FFMA t = dx*dx + t ; FFMA t = dz*dz + t
 * i.e. d2 = fmaf(dz, dz, fmaf(dx, dx, dy*dy)). Built with -ffp-contract=off so only the
 * explicit fmaf calls below fuse.
 *
 * Parity pin: validated against oracle/_ref (the reference .cu files compiled unmodified) on the
 * GPU by tests/test_reference_kernels_gpu.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float sqdist_ref(float dx, float dy, float dz) {
    return __builtin_fmaf(dz, dz, __builtin_fmaf(dx, dx, dy * dy));
}

/* ---- max-heap helpers: knnquery_cuda_kernel.cu:21-48 ---- */
static inline void sift_down(float *d, int *ix, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && d[child + 1] > d[child]) child++;
        if (d[root] > d[child]) return;
        float td = d[root]; d[root] = d[child]; d[child] = td;
        int ti = ix[root]; ix[root] = ix[child]; ix[child] = ti;
        root = child;
        child = 2 * root + 1;
    }
}

static inline void heap_to_sorted(float *d, int *ix, int k) {
    for (int i = k - 1; i > 0; i--) {
        float td = d[0]; d[0] = d[i]; d[i] = td;
        int ti = ix[0]; ix[0] = ix[i]; ix[i] = ti;
        sift_down(d, ix, i);
    }
}

/* segment lookup: knnquery_cuda_kernel.cu:51-62 */
static inline int segment_of(int i, const int *ends) {
    int s = 0;
    while (i >= ends[s]) s++;
    return s;
}

/*
 * xyz (n,3), new_xyz (m,3) row-major f32; offset/new_offset: cumulative segment ends (b,).
 * idx (m,nsample) int32, dist2 (m,nsample) f32 (squared distances, ascending).
 * nsample <= 100 like the reference's fixed local arrays (knnquery_cuda_kernel.cu:86-87).
 */
int oracle_knnquery(int m, int nsample, const float *xyz, const float *new_xyz,
                    const int *offset, const int *new_offset, int *idx, float *dist2) {
    if (nsample > 100 || nsample < 1) return 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int q = 0; q < m; q++) {
        float bd[100];
        int bi[100];
        int seg = segment_of(q, new_offset);
        int start = seg == 0 ? 0 : offset[seg - 1];
        int end = offset[seg];
        float qx = new_xyz[3 * q], qy = new_xyz[3 * q + 1], qz = new_xyz[3 * q + 2];
        for (int i = 0; i < nsample; i++) { bd[i] = 1e10f; bi[i] = start; }
        for (int i = start; i < end; i++) {
            float d2 = sqdist_ref(qx - xyz[3 * i], qy - xyz[3 * i + 1], qz - xyz[3 * i + 2]);
            if (d2 < bd[0]) {           /* strict <, kernel.cu:97 */
                bd[0] = d2; bi[0] = i;
                sift_down(bd, bi, nsample);
            }
        }
        heap_to_sorted(bd, bi, nsample);
        for (int i = 0; i < nsample; i++) {
            idx[(size_t)q * nsample + i] = bi[i];
            dist2[(size_t)q * nsample + i] = bd[i];
        }
    }
    return 0;
}

/* cuda_utils.h:11-14 (the double-precision log ratio is kept on purpose) */
static int ref_block_size(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

int oracle_fps_block_size(int n_max) { return ref_block_size(n_max); }

/* host threads used by the OpenMP loops (results do not depend on it); returns the previous maximum */
int oracle_set_threads(int n) { int old = omp_get_max_threads(); if (n > 0) omp_set_num_threads(n); return old; }

/*
 * One "block" of bs emulated threads per batch element; tmp (n,) must arrive filled with 1e10
 * (pointops.py:22). Strided in-thread scan with strict '>' then the power-of-two tree with
 * "v2 > v1 ? i2 : i1" (sampling_cuda_kernel.cu:5-10,49-123).
 */
int oracle_furthestsampling(int b, int n_max, const float *xyz, const int *offset,
                            const int *new_offset, float *tmp, int *idx) {
    const int bs = ref_block_size(n_max);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bid = 0; bid < b; bid++) {
        float *dv = (float *)malloc(sizeof(float) * (size_t)bs);
        int *di = (int *)malloc(sizeof(int) * (size_t)bs);
        int start_n = bid == 0 ? 0 : offset[bid - 1], end_n = offset[bid];
        int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
        int old = start_n;
        if (end_m > start_m) idx[start_m] = start_n;
        for (int j = start_m + 1; j < end_m; j++) {
            float x1 = xyz[3 * old], y1 = xyz[3 * old + 1], z1 = xyz[3 * old + 2];
            for (int t = 0; t < bs; t++) { dv[t] = -1.f; di[t] = start_n; }
            /* iterate k ascending; thread id = (k-start_n) % bs sees its ks ascending too */
            int t = 0;
            for (int k = start_n; k < end_n; k++) {
                float d = sqdist_ref(xyz[3 * k] - x1, xyz[3 * k + 1] - y1, xyz[3 * k + 2] - z1);
                float d2 = d < tmp[k] ? d : tmp[k];   /* min(d, tmp[k]) */
                tmp[k] = d2;
                if (d2 > dv[t]) { dv[t] = d2; di[t] = k; }
                if (++t == bs) t = 0;
            }
            for (int s = bs >> 1; s >= 1; s >>= 1)
                for (int u = 0; u < s; u++) {
                    float v1 = dv[u], v2 = dv[u + s];
                    int i1 = di[u], i2 = di[u + s];
                    dv[u] = v1 > v2 ? v1 : v2;
                    di[u] = v2 > v1 ? i2 : i1;
                }
            old = di[0];
            idx[j] = old;
        }
        free(dv); free(di);
    }
    return 0;
}
