// rrtk_arm.cu -- C-space occupancy grid of a planar N-link arm among circles
// (get_occupancy_grid arm02:79-110, NLinkArm.update_points :257-262, detect_collision :46-76).
//
// One thread per joint-space cell (i, j), one CTA per 128 columns of a row at a time; the forward kinematics of the cell
// (prefix-sum angles with the reference's 2-angle rule, correctly rounded cos/sin) is computed ONCE and tested against
// every obstacle set (S sets x O circles staged in shared memory), so trigonometry is amortised over S.
// Per (link, circle): a division-free FP64 filter on squared distances decides all clear cases; only
// circles within a 1e-9 relative band of tangency run the reference's exact operation sequence
// (norm / projection / clamp with numpy's fma dot), so the grid equals the reference's bit for bit.
// Link 1 is tested once per row, links 2..n as one straight stretch first (see arm_grid_kernel).
// Output is uint8 [S][rows][M], written 32 contiguous bytes per warp and set.
#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int ARM_MAX_LINKS = 16;
constexpr int ARM_THREADS = 128;

struct ArmParams {
    int M, row0, n_rows, n_links, S, O;
    int stretch;   // links 2..n are tested as one straight stretch first (all of them longer than 0)
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

// One (link, circle) test: the division-free filter on squared distances scaled by |l|^2, the reference's sequence inside
// its band of 1e-9 (relative).  The filter takes fma's (it only has to be right to the band).
//   1 touching, 0 clear -- exactly what detect_collision returns
__device__ __forceinline__ int link_circle_filter(double ax, double ay, double bx, double by, double L2, double cx, double cy,
                                                  double r2) {
    const double l0 = bx - ax, l1 = by - ay;
    const double v0 = cx - ax, v1 = cy - ay;
    const double dotp = fma(v1, l1, v0 * l0);
    const double vvL2 = fma(v1, v1, v0 * v0) * L2;
    double sdist;  // = dist^2 * L2
    if (dotp <= 0.0) sdist = vvL2;
    else if (dotp >= L2) {
        const double w0 = cx - bx, w1 = cy - by;
        sdist = fma(w1, w1, w0 * w0) * L2;
    } else sdist = fma(-dotp, dotp, vvL2);
    const double thr = r2 * L2;
    const double t = sdist - thr, band = 1e-9 * (vvL2 + thr);
    if (t > band) return 0;                       // certainly clear
    if (t < -band && L2 > 0.0) return 1;          // certainly touching
    return -1;
}
__device__ __noinline__ bool link_circle_hit(double ax, double ay, double bx, double by, double cx, double cy, double r, double r2) {
    const double l0 = bx - ax, l1 = by - ay;
    const int v = link_circle_filter(ax, ay, bx, by, l0 * l0 + l1 * l1, cx, cy, r2);
    return v >= 0 ? v != 0 : detect_collision_exact(ax, ay, bx, by, cx, cy, r);
}

// MAXL = compile-time bound on the number of links (the per-cell joint positions live in registers).
// The reference's update_points gives joint 1 the angle theta1 and EVERY later joint theta1 + theta2 (arm02:257-262 sums
// joint_angles[:i] of a two-entry list), so
//   * link 1 depends on the row only: its S x O tests run once per row tile (thread s takes obstacle set s) and land in
//     shared memory as one flag per set;
//   * links 2..n lie end to end on one straight line from joint 1: a circle clear of the whole stretch [p1, pn] beyond the
//     band is clear of every link, one that reaches into it beyond the band touches the link its nearest point lies on
//     (the joints computed by the reference's running sums are within ~1e-15 of that line, the band is 1e-9), and only a
//     circle inside the band of the stretch runs the per-link tests.  (Not with a zero-length link among them: the
//     reference's 0 / 0 makes that link "touch" everything -- the host clears `stretch` then.)
// One CTA works on 128 consecutive columns of one row at a time.
template <int MAXL>
__global__ void __launch_bounds__(ARM_THREADS)
arm_grid_kernel(ArmParams p, const double *__restrict__ theta, const double *__restrict__ obstacles,
                uint8_t *__restrict__ grid) {
    extern __shared__ double s_obs[];  // [S * O][4]: x, y, r, r * r; then S row flags
    const int n_circ = p.S * p.O;
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_obs + 4 * (size_t)n_circ);
    for (int t = threadIdx.x; t < n_circ; t += blockDim.x) {
        double x = obstacles[3 * t], y = obstacles[3 * t + 1], r = obstacles[3 * t + 2];
        s_obs[4 * t] = x; s_obs[4 * t + 1] = y; s_obs[4 * t + 2] = r; s_obs[4 * t + 3] = r * r;
    }
    const int tiles_per_row = (p.M + ARM_THREADS - 1) / ARM_THREADS;
    const long long tiles = (long long)p.n_rows * tiles_per_row;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int ir = (int)(tile / tiles_per_row), j = (int)(tile % tiles_per_row) * ARM_THREADS + threadIdx.x;
        const int i = p.row0 + ir;
        // forward kinematics (arm02:257-262): joint k uses theta1 (k = 1) or theta1 + theta2 (k >= 2)
        const double a1 = theta[i];
        const double c1 = crm_cos(a1), s1 = crm_sin(a1);
        double px[MAXL + 1], py[MAXL + 1];
        px[0] = 0.0; py[0] = 0.0;
        px[1] = px[0] + p.link[0] * c1;
        py[1] = py[0] + p.link[0] * s1;
        __syncthreads();   // the circles are staged / the previous tile's flags have been read
        for (int s = threadIdx.x; s < p.S; s += blockDim.x) {   // link 1 against obstacle set s
            const double *ob = s_obs + (size_t)s * p.O * 4;
            bool hit = false;
            for (int o = 0; o < p.O && !hit; o++)
                hit = link_circle_hit(px[0], py[0], px[1], py[1], ob[4 * o], ob[4 * o + 1], ob[4 * o + 2], ob[4 * o + 3]);
            s_row[s] = hit ? 1 : 0;
        }
        __syncthreads();
        if (j >= p.M) continue;
        const double a2 = a1 + theta[j];
        const double c2 = crm_cos(a2), s2 = crm_sin(a2);
#pragma unroll
        for (int k = 2; k <= MAXL; k++) {
            if (k <= p.n_links) {
                px[k] = px[k - 1] + p.link[k - 1] * c2;
                py[k] = py[k - 1] + p.link[k - 1] * s2;
            }
        }
        // the stretch of links 2..n
        double ex = px[1], ey = py[1];
#pragma unroll
        for (int k = 2; k <= MAXL; k++)
            if (k == p.n_links) { ex = px[k]; ey = py[k]; }
        const double u0 = ex - px[1], u1 = ey - py[1], U2 = fma(u1, u1, u0 * u0);
        for (int s = 0; s < p.S; s++) {
            const double *ob = s_obs + (size_t)s * p.O * 4;
            bool hit = s_row[s] != 0;
            if (p.n_links >= 2) {
                for (int o = 0; o < p.O && !hit; o++) {
                    const double cx = ob[4 * o], cy = ob[4 * o + 1], r2 = ob[4 * o + 3];
                    const int v = p.stretch ? link_circle_filter(px[1], py[1], ex, ey, U2, cx, cy, r2) : -1;
                    if (v >= 0) { hit = v != 0; continue; }
#pragma unroll
                    for (int k = 1; k < MAXL; k++)
                        if (k < p.n_links && !hit) hit = link_circle_hit(px[k], py[k], px[k + 1], py[k + 1], cx, cy, ob[4 * o + 2], r2);
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
    p.stretch = 1;
    for (int k = 1; k < n_links; k++)
        if (!(link_host[k] > 0.0)) p.stretch = 0;
    size_t smem = (size_t)S * O * 4 * sizeof(double) + (((size_t)S + 15) & ~(size_t)15);
    if (smem > 200 * 1024) return set_error(RRTK_ERR_INVALID, "S * O circles do not fit in shared memory (max 6400)");
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ARM_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        long long want = (long long)n_rows * ((M + ARM_THREADS - 1) / ARM_THREADS);   // row tiles
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
