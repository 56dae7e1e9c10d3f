// Correspondence RANSAC (rigid pose from putative correspondences), the estimator the reference's evaluator runs on the
// hot path's output.
//
// Replaces ransac_pose_estimation_correspondences (registration/benchmark_utils.py:165-209, called from
// registration/evaluate_registration_c2f.py:88), i.e. Open3D's registration_ransac_based_on_correspondence with
//   TransformationEstimationPointToPoint(False), ransac_n = 3,
//   checkers [CorrespondenceCheckerBasedOnEdgeLength(0.9), CorrespondenceCheckerBasedOnDistance(distance_threshold)],
//   RANSACConvergenceCriteria(50000, 1000)   (confidence 1000 clamps to 1.0 => log(1 - 1) = -inf => no early exit: exactly
//                                             max_iteration hypotheses are drawn).
// Open3D (third party, open3d==0.13.0 in the reference's requirements.txt) is not in the tree; its published algorithm is
// restated (oracle/ransac_ref.py does the same in numpy, "parity unpinned" for the Open3D part):
//   per iteration: draw ransac_n correspondences uniformly WITH replacement; T = Umeyama (no scaling) of the sample;
//   reject unless every pair of sampled correspondences satisfies |s_i-s_j| >= 0.9|t_i-t_j| and |t_i-t_j| >= 0.9|s_i-s_j|
//   and every sampled correspondence satisfies |T s - t| <= thr; score T over ALL correspondences: inliers |T s - t| < thr,
//   fitness = inliers / n, rmse = sqrt(sum d^2 / inliers); keep the best by (fitness desc, rmse asc).
// Open3D's result depends on its process-global RNG and on OpenMP scheduling, so it is not reproducible bit for bit by
// construction; here the hypothesis sequence is a pure function of (seed, iteration) - a counter-based hash the oracle
// shares - and ties are broken by the lower iteration, so results are deterministic and oracle-comparable.
//
// Design: all arithmetic in fp64 like Open3D (Eigen double). grid = (iteration chunks, pairs). A CTA walks its chunk in
// rounds of RS_THREADS hypotheses: (1) one thread per hypothesis - sample, edge-length check (needs no transform, so it
// runs first and most outlier-heavy samples stop here), 3-point Umeyama as the dominant eigenvector of Horn's 4x4 matrix
// (horn.cuh), distance check; survivors go to a shared-memory list; (2) one WARP per survivor scores it over the pair's
// correspondences, which sit in shared memory (24 B each), lanes striding, shuffle reduction. Per-CTA bests go to global
// memory and a second tiny kernel reduces them per pair and writes the 4x4 transform. The edge-length test runs before the
// transform is computed (Open3D computes the transform first); the accepted set is identical.
#include "../../include/roitr_b200.h"
#include "common.cuh"
#include "horn.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 4;                      // hypotheses per CTA = RS_THREADS * RS_ROUNDS
constexpr int RS_CHUNK = RS_THREADS * RS_ROUNDS;
constexpr int RS_MAX_CORR = 8192;                 // correspondences per pair held in shared memory (196 KB)

struct RsBest {
    int inliers;
    int itr;
    double err2;
    double T[12];                                 // R row-major, then t
};

__device__ __forceinline__ bool rs_better(int c1, double e1, int i1, int c0, double e0, int i0) {
    if (c1 != c0) return c1 > c0;
    if (e1 != e0) return e1 < e0;
    return i1 < i0;
}

// counter-based sample index in [0, n): PCG-RXS-M-XS-32 output function over a Weyl-style state (oracle/ransac_ref.py)
__device__ __host__ __forceinline__ uint32_t rs_hash(uint32_t seed, uint32_t counter) {
    uint32_t h = counter * 747796405u + seed * 2891336453u + 1u;
    h = ((h >> ((h >> 28) + 4u)) ^ h) * 277803737u;
    return (h >> 22) ^ h;
}

__global__ void __launch_bounds__(RS_THREADS) ransac_kernel(int iters, int ransac_n_unused, const float* __restrict__ src,
                                                            const float* __restrict__ tgt, const int* __restrict__ offset,
                                                            double thr, double edge_sim, uint32_t seed, RsBest* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pair = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int begin = pair == 0 ? 0 : offset[pair - 1], n = offset[pair] - begin;
    float* sp = reinterpret_cast<float*>(smem_raw);            // n x 3 source, n x 3 target
    float* tp = sp + 3 * (size_t)n;
    __shared__ double s_T[RS_THREADS][12];
    __shared__ int s_itr[RS_THREADS];
    __shared__ int s_count;
    __shared__ RsBest s_best[RS_THREADS / 32];
    for (int i = tid; i < 3 * n; i += RS_THREADS) {
        sp[i] = __ldg(src + 3 * (size_t)begin + i);
        tp[i] = __ldg(tgt + 3 * (size_t)begin + i);
    }
    RsBest best;                                               // per warp, meaningful in lane 0
    best.inliers = -1; best.itr = 0x7fffffff; best.err2 = 0.0;
    for (int k = 0; k < 12; ++k) best.T[k] = 0.0;
    const uint32_t pair_seed = seed + 0x9e3779b9u * (uint32_t)pair;
    for (int round = 0; round < RS_ROUNDS; ++round) {
        if (tid == 0) s_count = 0;
        __syncthreads();
        const int itr = blockIdx.x * RS_CHUNK + round * RS_THREADS + tid;
        if (itr < iters && n >= 3) {
            int id[3];
            double s[3][3], t[3][3];
            for (int j = 0; j < 3; ++j) {
                id[j] = (int)(((uint64_t)rs_hash(pair_seed, (uint32_t)itr * 3u + (uint32_t)j) * (uint64_t)n) >> 32);
                for (int d = 0; d < 3; ++d) { s[j][d] = (double)sp[3 * id[j] + d]; t[j][d] = (double)tp[3 * id[j] + d]; }
            }
            bool ok = true;
            for (int i = 0; i < 3 && ok; ++i)               // CorrespondenceCheckerBasedOnEdgeLength
                for (int j = i + 1; j < 3; ++j) {
                    double ds = 0.0, dt = 0.0;
                    for (int d = 0; d < 3; ++d) { const double a = s[i][d] - s[j][d], b = t[i][d] - t[j][d]; ds += a * a; dt += b * b; }
                    ds = sqrt(ds); dt = sqrt(dt);
                    if (!(ds >= dt * edge_sim && dt >= ds * edge_sim)) { ok = false; break; }
                }
            if (ok) {
                double ms[3], mt[3], H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, R[9], T[12];
                for (int d = 0; d < 3; ++d) { ms[d] = (s[0][d] + s[1][d] + s[2][d]) / 3.0; mt[d] = (t[0][d] + t[1][d] + t[2][d]) / 3.0; }
                for (int j = 0; j < 3; ++j)
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) H[3 * r + c] += (s[j][r] - ms[r]) * (t[j][c] - mt[c]);
                horn_rotation(H, R);
                for (int k = 0; k < 9; ++k) T[k] = R[k];
                for (int r = 0; r < 3; ++r) T[9 + r] = mt[r] - (R[3 * r] * ms[0] + R[3 * r + 1] * ms[1] + R[3 * r + 2] * ms[2]);
                for (int j = 0; j < 3 && ok; ++j) {          // CorrespondenceCheckerBasedOnDistance
                    double d2 = 0.0;
                    for (int r = 0; r < 3; ++r) {
                        const double v = T[3 * r] * s[j][0] + T[3 * r + 1] * s[j][1] + T[3 * r + 2] * s[j][2] + T[9 + r] - t[j][r];
                        d2 += v * v;
                    }
                    if (!(sqrt(d2) <= thr)) ok = false;
                }
                if (ok) {
                    const int slot = atomicAdd(&s_count, 1);
                    s_itr[slot] = itr;
                    for (int k = 0; k < 12; ++k) s_T[slot][k] = T[k];
                }
            }
        }
        __syncthreads();
        const int survivors = s_count;
        for (int h = warp; h < survivors; h += RS_THREADS / 32) {   // one warp scores one surviving hypothesis
            double T[12];
            for (int k = 0; k < 12; ++k) T[k] = s_T[h][k];
            int cnt = 0;
            double e2 = 0.0;
            for (int i = lane; i < n; i += 32) {
                const double x = (double)sp[3 * i], y = (double)sp[3 * i + 1], z = (double)sp[3 * i + 2];
                const double dx = T[0] * x + T[1] * y + T[2] * z + T[9] - (double)tp[3 * i];
                const double dy = T[3] * x + T[4] * y + T[5] * z + T[10] - (double)tp[3 * i + 1];
                const double dz = T[6] * x + T[7] * y + T[8] * z + T[11] - (double)tp[3 * i + 2];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (sqrt(d2) < thr) { ++cnt; e2 += d2; }
            }
            // fixed-order tree reduction: the same (count, err2) whatever warp scored the hypothesis
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
                e2 += __shfl_xor_sync(FULL_MASK, e2, o);
            }
            const int hi = s_itr[h];
            if (rs_better(cnt, e2, hi, best.inliers, best.err2, best.itr)) {
                best.inliers = cnt; best.err2 = e2; best.itr = hi;
                for (int k = 0; k < 12; ++k) best.T[k] = T[k];
            }
        }
        __syncthreads();
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
        RsBest b = s_best[0];
        for (int w = 1; w < RS_THREADS / 32; ++w)
            if (rs_better(s_best[w].inliers, s_best[w].err2, s_best[w].itr, b.inliers, b.err2, b.itr)) b = s_best[w];
        partial[(size_t)pair * gridDim.x + blockIdx.x] = b;
    }
}

__global__ void ransac_reduce_kernel(int chunks, const RsBest* __restrict__ partial, const int* __restrict__ offset,
                                     double* __restrict__ transform, double* __restrict__ fitness, double* __restrict__ rmse,
                                     int* __restrict__ best_itr) {
    const int pair = blockIdx.x;
    if (threadIdx.x != 0) return;
    RsBest b = partial[(size_t)pair * chunks];
    for (int c = 1; c < chunks; ++c) {
        const RsBest& q = partial[(size_t)pair * chunks + c];
        if (rs_better(q.inliers, q.err2, q.itr, b.inliers, b.err2, b.itr)) b = q;
    }
    const int n = offset[pair] - (pair == 0 ? 0 : offset[pair - 1]);
    double* T = transform + 16 * (size_t)pair;
    if (b.inliers <= 0) {          // RegistrationResult(): identity, fitness 0, rmse 0
        for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.0 : 0.0;
        fitness[pair] = 0.0; rmse[pair] = 0.0; best_itr[pair] = -1;
        return;
    }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = b.T[3 * r + c];
        T[4 * r + 3] = b.T[9 + r];
    }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    fitness[pair] = (double)b.inliers / (double)n;
    rmse[pair] = sqrt(b.err2 / (double)b.inliers);
    best_itr[pair] = b.itr;
}

}  // namespace

extern "C" long long roitr_ransac_workspace_bytes(int pairs, int iterations) {
    return (long long)pairs * ceil_div(iterations > 0 ? iterations : 1, RS_CHUNK) * (long long)sizeof(RsBest);
}

extern "C" int roitr_ransac_correspondences(int pairs, int iterations, const float* src, const float* tgt, const int* offset,
                                            int max_corr, double distance_threshold, double edge_similarity, unsigned seed,
                                            void* workspace, double* transform, double* fitness, double* rmse, int* best_itr,
                                            void* stream) {
    ROITR_CHECK_ARG(pairs >= 1 && iterations >= 1 && src && tgt && offset && workspace && transform && fitness && rmse && best_itr,
                    "ransac_correspondences: bad arguments");
    ROITR_CHECK_ARG(max_corr >= 0 && max_corr <= RS_MAX_CORR, "ransac_correspondences: at most %d correspondences per pair (got %d)",
                    RS_MAX_CORR, max_corr);
    ROITR_CHECK_ARG(distance_threshold > 0.0, "ransac_correspondences: distance_threshold must be positive");
    const int chunks = ceil_div(iterations, RS_CHUNK);
    const size_t smem = (size_t)max_corr * 24 + 16;
    static size_t configured_dev[ROITR_MAX_DEVICES] = {};
    size_t& configured = configured_dev[roitr_cur_device()];
    if (smem > 16 * 1024 && smem > configured) {
        ROITR_CUDA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ransac_kernel<<<dim3(chunks, pairs), RS_THREADS, smem, st>>>(iterations, 3, src, tgt, offset, distance_threshold, edge_similarity,
                                                                 seed, (RsBest*)workspace);
    ROITR_CHECK_LAUNCH("ransac_kernel");
    ransac_reduce_kernel<<<pairs, 32, 0, st>>>(chunks, (const RsBest*)workspace, offset, transform, fitness, rmse, best_itr);
    ROITR_CHECK_LAUNCH("ransac_reduce_kernel");
    return ROITR_OK;
}
