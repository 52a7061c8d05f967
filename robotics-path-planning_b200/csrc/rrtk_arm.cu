// rrtk_arm.cu -- C-space occupancy grid of a planar N-link arm among circles
// (get_occupancy_grid arm02:79-110, NLinkArm.update_points :257-262, detect_collision :46-76).
//
// One thread per joint-space cell (i, j); the forward kinematics of the cell (prefix-sum angles with the
// reference's 2-angle rule, correctly rounded cos/sin) is computed ONCE and tested against every obstacle
// set (S sets x O circles staged in shared memory), so trigonometry is amortised over S.
// Per (link, circle): a division-free FP64 filter on squared distances decides all clear cases; only
// circles within a 1e-9 relative band of tangency run the reference's exact operation sequence
// (norm / projection / clamp with numpy's fma dot), so the grid equals the reference's bit for bit.
// Output is uint8 [S][rows][M], written 32 contiguous bytes per warp and set.
#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int ARM_MAX_LINKS = 16;
constexpr int ARM_THREADS = 128;

struct ArmParams {
    int M, row0, n_rows, n_links, S, O;
    double link[ARM_MAX_LINKS];
};

// numpy's 2-vector dot on the reference platform: fma(a1, b1, a0 * b0)
__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

// detect_collision (arm02:46-76), the reference's exact operation sequence
__device__ __noinline__ bool detect_collision_exact(double ax, double ay, double bx, double by, double cx,
                                                    double cy, double r) {
    double l0 = bx - ax, l1 = by - ay;
    double mag = sqrt(dot2(l0, l1, l0, l1));
    double v0 = cx - ax, v1 = cy - ay;
    double proj = dot2(v0, v1, l0 / mag, l1 / mag);
    double p0, p1;
    if (proj <= 0) { p0 = ax; p1 = ay; }
    else if (proj >= mag) { p0 = bx; p1 = by; }
    else { p0 = ax + l0 * proj / mag; p1 = ay + l1 * proj / mag; }
    double w0 = p0 - cx, w1 = p1 - cy;
    double dist = sqrt(dot2(w0, w1, w0, w1));
    return !(dist > r);
}

// MAXL = compile-time bound on the number of links (the per-cell joint positions live in registers)
template <int MAXL>
__global__ void __launch_bounds__(ARM_THREADS)
arm_grid_kernel(ArmParams p, const double *__restrict__ theta, const double *__restrict__ obstacles,
                uint8_t *__restrict__ grid) {
    extern __shared__ double s_obs[];  // [S * O][4]: x, y, r, r * r
    const int n_circ = p.S * p.O;
    for (int t = threadIdx.x; t < n_circ; t += blockDim.x) {
        double x = obstacles[3 * t], y = obstacles[3 * t + 1], r = obstacles[3 * t + 2];
        s_obs[4 * t] = x; s_obs[4 * t + 1] = y; s_obs[4 * t + 2] = r; s_obs[4 * t + 3] = r * r;
    }
    __syncthreads();
    const long long cells = (long long)p.n_rows * p.M;
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < cells;
         cell += (long long)gridDim.x * blockDim.x) {
        const int ir = (int)(cell / p.M), j = (int)(cell % p.M);
        const int i = p.row0 + ir;
        // forward kinematics (arm02:257-262): joint k uses theta1 (k = 1) or theta1 + theta2 (k >= 2)
        const double a1 = theta[i], a2 = theta[i] + theta[j];
        const double c1 = crm_cos(a1), s1 = crm_sin(a1), c2 = crm_cos(a2), s2 = crm_sin(a2);
        double px[MAXL + 1], py[MAXL + 1], l2[MAXL];
        px[0] = 0.0; py[0] = 0.0;
#pragma unroll
        for (int k = 1; k <= MAXL; k++) {
            if (k <= p.n_links) {
                px[k] = px[k - 1] + p.link[k - 1] * (k == 1 ? c1 : c2);
                py[k] = py[k - 1] + p.link[k - 1] * (k == 1 ? s1 : s2);
                double l0 = px[k] - px[k - 1], l1 = py[k] - py[k - 1];
                l2[k - 1] = l0 * l0 + l1 * l1;
            }
        }
        for (int s = 0; s < p.S; s++) {
            const double *ob = s_obs + (size_t)s * p.O * 4;
            bool hit = false;
#pragma unroll
            for (int k = 0; k < MAXL; k++) {
                if (k < p.n_links && !hit) {
                    const double ax = px[k], ay = py[k], bx = px[k + 1], by = py[k + 1];
                    const double l0 = bx - ax, l1 = by - ay, L2 = l2[k];
                    for (int o = 0; o < p.O && !hit; o++) {
                        const double cx = ob[4 * o], cy = ob[4 * o + 1], r2 = ob[4 * o + 3];
                        // filter: squared distance to the segment, scaled by |l|^2, no division / sqrt
                        const double v0 = cx - ax, v1 = cy - ay;
                        const double dotp = v0 * l0 + v1 * l1;
                        const double vv = v0 * v0 + v1 * v1;
                        double sdist;  // = dist^2 * L2
                        if (dotp <= 0.0) sdist = vv * L2;
                        else if (dotp >= L2) {
                            const double w0 = cx - bx, w1 = cy - by;
                            sdist = (w0 * w0 + w1 * w1) * L2;
                        } else sdist = vv * L2 - dotp * dotp;
                        const double thr = r2 * L2;
                        const double band = 1e-9 * (vv * L2 + thr);
                        if (sdist > thr + band) continue;                 // certainly clear
                        if (sdist < thr - band && L2 > 0.0) { hit = true; break; }  // certainly touching
                        hit = detect_collision_exact(ax, ay, bx, by, cx, cy, ob[4 * o + 2]);
                    }
                }
            }
            grid[((size_t)s * p.n_rows + ir) * p.M + j] = hit ? 1 : 0;
        }
    }
}

int launch_arm_grid(int M, const double *theta, int row0, int n_rows, int n_links, const double *link_host,
                    const double *obstacles, int S, int O, uint8_t *grid, cudaStream_t s) {
    ArmParams p;
    p.M = M; p.row0 = row0; p.n_rows = n_rows; p.n_links = n_links; p.S = S; p.O = O;
    for (int k = 0; k < ARM_MAX_LINKS; k++) p.link[k] = k < n_links ? link_host[k] : 0.0;
    size_t smem = (size_t)S * O * 4 * sizeof(double);
    if (smem > 200 * 1024) return set_error(RRTK_ERR_INVALID, "S * O circles do not fit in shared memory (max 6400)");
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ARM_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        long long cells = (long long)n_rows * M;
        long long want = (cells + ARM_THREADS - 1) / ARM_THREADS;
        long long grid_dim = (long long)sms * per_sm;  // persistent, a multiple of the SM count
        if (grid_dim > want) grid_dim = want;
        if (grid_dim < 1) grid_dim = 1;
        kernel<<<(unsigned)grid_dim, ARM_THREADS, smem, s>>>(p, theta, obstacles, grid);
        return cudaSuccess;
    };
    cudaError_t e;
    if (n_links <= 2) e = launch(arm_grid_kernel<2>);
    else if (n_links <= 5) e = launch(arm_grid_kernel<5>);
    else if (n_links <= 8) e = launch(arm_grid_kernel<8>);
    else e = launch(arm_grid_kernel<ARM_MAX_LINKS>);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(arm_grid_kernel)");
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "arm_grid_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
