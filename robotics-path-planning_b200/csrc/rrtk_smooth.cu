// rrtk_smooth.cu -- path_smoothing (rrt_04:1447-1479) for Q final courses at once: the step right after the planning
// loop, so the paths never leave the GPU between search_best_goal_node / generate_final_course
// (rrtk_extract_paths_dev) and the smoothed result.
//
// One warp per path.  The path and its segment lengths live in shared memory; the lengths (math.hypot of each pair,
// correctly rounded like CPython's) are recomputed in parallel only when the path changes -- get_path_length
// (rrt_04:1390-1398) and get_target_point (:1401-1420) re-evaluate the same hypot values every call in the reference,
// and only their left-to-right running sums (kept sequential here, the rounding order matters) differ per call.
// line_collision_check (:1423-1444, the infinite line through the two picks, not the segment) runs with the lanes
// split over the circles.  Upstream quirks kept: `ti = i - 1`, Python's negative index when ti = -1, the strict
// continue conditions, and the path that can grow by one point per shortcut.  Where the reference would raise
// ZeroDivisionError (a zero-length pair under the pick, or coincident picks) the query stops with RRTK_Q_DIV_ZERO.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int SM_WARPS = 4;
constexpr int SM_CAP = 512;   // points per path held in shared memory

struct SmoothWarp {
    double x[SM_CAP], y[SM_CAP], d[SM_CAP], tx[SM_CAP], ty[SM_CAP];
};

// get_target_point (rrt_04:1401-1420) on the cached segment lengths; uniform across the warp.  false = ZeroDivisionError
__device__ __forceinline__ bool target_point(const SmoothWarp &W, int len, double target, double &x, double &y, int &ti) {
    double le = 0.0, last = 0.0;
    ti = 0;
    for (int i = 0; i + 1 < len; i++) {
        const double d = W.d[i];
        le += d;
        if (le >= target) { ti = i - 1; last = d; break; }
    }
    if (last == 0.0) return false;
    const double ratio = (le - target) / last;
    const int a = ti < 0 ? len + ti : ti;   // Python's negative index
    x = W.x[a] + (W.x[ti + 1] - W.x[a]) * ratio;
    y = W.y[a] + (W.y[ti + 1] - W.y[a]) * ratio;
    return true;
}

__global__ void __launch_bounds__(SM_WARPS * 32)
smooth_paths_kernel(int n_queries, int path_cap, int max_iter, double2 *path_all, int32_t *len_all,
                    const double2 *__restrict__ draws, const double *__restrict__ obs3, int obs_stride,
                    const int32_t *__restrict__ n_obs_arr, int32_t *status_out, int32_t *iters_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    SmoothWarp &W = reinterpret_cast<SmoothWarp *>(smem_raw)[wib];
    const int q = blockIdx.x * SM_WARPS + wib;
    if (q >= n_queries) return;
    double2 *path = path_all + (size_t)q * path_cap;
    const double2 *dr = draws + (size_t)q * max_iter;
    const double *obs = obs3 + (size_t)q * obs_stride * 3;
    const int n_obs = n_obs_arr ? n_obs_arr[q] : 0;
    int len = len_all[q], status = RRTK_Q_OK, it = 0;
    if (len > path_cap || len > SM_CAP) { if (lane == 0) { status_out[q] = RRTK_Q_PATH_OVERFLOW; iters_out[q] = 0; } return; }
    for (int k = lane; k < len; k += 32) { const double2 a = path[k]; W.x[k] = a.x; W.y[k] = a.y; }
    __syncwarp();
    double le = 0.0;
    bool dirty = true;
    for (it = 0; it < max_iter && len >= 1; it++) {
        if (dirty) {   // segment lengths of the current path, then get_path_length's running sum (left to right)
            for (int k = lane; k + 1 < len; k += 32) W.d[k] = crm_hypot(W.x[k + 1] - W.x[k], W.y[k + 1] - W.y[k]);
            __syncwarp();
            le = 0.0;
            for (int k = 0; k + 1 < len; k++) le += W.d[k];
            dirty = false;
        }
        const double2 u = dr[it];
        double p0 = 0 + (le - 0) * u.x, p1 = 0 + (le - 0) * u.y;   // random.uniform(0, le) = a + (b - a) * random()
        if (p1 < p0) { const double t = p0; p0 = p1; p1 = t; }
        double fx, fy, sx, sy;
        int t1, t2;
        if (!target_point(W, len, p0, fx, fy, t1) || !target_point(W, len, p1, sx, sy, t2)) { status |= RRTK_Q_DIV_ZERO; break; }
        if (t1 <= 0 || t2 <= 0) continue;
        if (t2 + 1 > len) continue;
        if (t2 == t1) continue;
        // line_collision_check (rrt_04:1423-1444)
        const double a = sy - fy, b = -(sx - fx), c = sy * (sx - fx) - sx * (sy - fy);
        bool blocked = false;
        if (n_obs > 0) {
            const double h = crm_hypot(a, b);
            if (h == 0.0) { status |= RRTK_Q_DIV_ZERO; break; }
            for (int o = lane; o < n_obs && !blocked; o += 32)
                blocked = fabs(a * obs[3 * o] + b * obs[3 * o + 1] + c) / h <= obs[3 * o + 2];
            blocked = __any_sync(0xffffffffu, blocked);
        }
        if (blocked) continue;
        // newPath = path[:t1 + 1] + [first] + [second] + path[t2 + 1:]
        const int tail = len - t2 - 1, nl = t1 + 3 + tail;
        if (nl > path_cap || nl > SM_CAP) { status |= RRTK_Q_PATH_OVERFLOW; break; }
        for (int k = lane; k < tail; k += 32) { W.tx[k] = W.x[t2 + 1 + k]; W.ty[k] = W.y[t2 + 1 + k]; }
        __syncwarp();
        for (int k = lane; k < tail; k += 32) { W.x[t1 + 3 + k] = W.tx[k]; W.y[t1 + 3 + k] = W.ty[k]; }
        if (lane == 0) { W.x[t1 + 1] = fx; W.y[t1 + 1] = fy; W.x[t1 + 2] = sx; W.y[t1 + 2] = sy; }
        __syncwarp();
        len = nl;
        dirty = true;
    }
    __syncwarp();
    for (int k = lane; k < len; k += 32) path[k] = make_double2(W.x[k], W.y[k]);
    if (lane == 0) { len_all[q] = len; status_out[q] = status; iters_out[q] = it; }
}

int launch_smooth_paths(int n_queries, int path_cap, int max_iter, double *path, int32_t *path_len, const double *draws,
                        const double *obs3, int obs_stride, const int32_t *n_obs, int32_t *status, int32_t *iters_done,
                        cudaStream_t s) {
    const size_t smem = sizeof(SmoothWarp) * SM_WARPS;
    cudaError_t e = cudaFuncSetAttribute(smooth_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(smooth_paths_kernel)");
    const unsigned grid = (unsigned)((n_queries + SM_WARPS - 1) / SM_WARPS);
    smooth_paths_kernel<<<grid, SM_WARPS * 32, smem, s>>>(n_queries, path_cap, max_iter, reinterpret_cast<double2 *>(path),
                                                         path_len, reinterpret_cast<const double2 *>(draws), obs3,
                                                         obs_stride, n_obs, status, iters_done);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "smooth_paths_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
