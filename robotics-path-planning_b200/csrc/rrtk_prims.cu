// rrtk_prims.cu -- the reference's per-step methods as stand-alone device calls: what the methods of
// rrtk.RRT / rrtk.RRTStar run when user code calls them one at a time (the planning loop itself is the fused kernel).
//   steer's path_x / path_y            rrt_04:1086-1115   steer_points_kernel
//   check_collision of a point list     rrt_04:1216-1230   points_collide_kernel
//   get_nearest_node_index              rrt_04:1196-1202   nearest_f64_kernel   (FP64, first minimum)
//   find_near_nodes                     rrt_04:1314-1338   near_f64_kernel      (FP64, ascending, the `.index()` mapping)
// Same arithmetic as the planner kernels: the reference's operation order, correctly rounded hypot / atan2 / cos / sin,
// no multiply-add contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_rrtstar_common.cuh"

namespace rrtk {

// one lane per edge: the points steer appends to path_x / path_y, in order
__global__ void steer_points_kernel(long long n_req, const double2 *__restrict__ from_xy, const double2 *__restrict__ to_xy,
                                    const double *__restrict__ extend, double extend_all, double res, int pt_cap,
                                    double2 *points, int32_t *n_points) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_req) return;
    const double2 f = from_xy[r], t = to_xy[r];
    const Steer st = steer(f.x, f.y, t.x, t.y, extend ? extend[r] : extend_all, res);
    double2 *out = points + (size_t)r * pt_cap;
    int np = 0;
    double x = f.x, y = f.y;
    if (np < pt_cap) out[np] = make_double2(x, y);
    np++;
    for (int k = 0; k < st.n; k++) {
        x += st.stx;
        y += st.sty;
        if (np < pt_cap) out[np] = make_double2(x, y);
        np++;
    }
    if (st.snap) {
        if (np < pt_cap) out[np] = t;
        np++;
    }
    n_points[r] = np;        // len(path_x); > pt_cap = truncated
}

// one warp per point list: min over the points of the squared distance to each circle <= (size + robot_radius)**2
__global__ void points_collide_kernel(int n_req, const double2 *__restrict__ points, const int32_t *__restrict__ n_points,
                                      int pt_cap, const int32_t *__restrict__ obs_set, const double4 *__restrict__ obstacles,
                                      int obs_stride, const int32_t *__restrict__ n_obs_arr, uint8_t *free_flag) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n_req) return;
    const int set = obs_set ? obs_set[r] : 0;
    const double4 *obs = obstacles + (size_t)set * obs_stride;
    const int m = n_obs_arr ? n_obs_arr[set] : 0;
    const int np = n_points[r] < pt_cap ? n_points[r] : pt_cap;
    const double2 *pts = points + (size_t)r * pt_cap;
    bool hit = false;
    for (int j = lane; j < m && !hit; j += 32) {
        const double4 o = obs[j];
        for (int k = 0; k < np; k++) {
            const double dx = o.x - pts[k].x, dy = o.y - pts[k].y;
            if (dx * dx + dy * dy <= o.w) { hit = true; break; }
        }
    }
    const unsigned any = __ballot_sync(FULL, hit);
    if (lane == 0) free_flag[r] = any ? 0 : 1;
}

// one CTA per sample: dlist.index(min(dlist)) -- the first minimum of (x - sx)**2 + (y - sy)**2
__global__ void __launch_bounds__(256) nearest_f64_kernel(const double2 *__restrict__ xy, long long n,
                                                          const double2 *__restrict__ samples, int32_t *out_idx, double *out_d2) {
    __shared__ double s_d[8];
    __shared__ int s_i[8];
    const double2 smp = samples[blockIdx.x];
    double bd = CUDART_INF;
    int bi = 0x7fffffff;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 a = xy[i];
        const double dx = a.x - smp.x, dy = a.y - smp.y;
        const double d = dx * dx + dy * dy;
        if (d < bd) { bd = d; bi = (int)i; }
    }
    warp_argmin(bd, bi);
    if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = bd; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (s_d[w] < bd || (s_d[w] == bd && s_i[w] < bi)) { bd = s_d[w]; bi = s_i[w]; }
        out_idx[blockIdx.x] = bi;
        if (out_d2) out_d2[blockIdx.x] = bd;
    }
}

// one CTA: [dist_list.index(d) for d in dist_list if d <= r2] -- ascending node order, every hit replaced by the first
// node (of the whole list) with the same squared distance
__global__ void __launch_bounds__(256) near_f64_kernel(const double2 *__restrict__ xy, int n, double cx, double cy, double r2,
                                                       int32_t *out_idx, double *scratch_d2, int cap, int32_t *out_n) {
    __shared__ int s_cnt[9];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += blockDim.x) {
        const int i = b0 + threadIdx.x;
        bool hit = false;
        double d = 0.0;
        if (i < n) {
            const double2 a = xy[i];
            const double dx = a.x - cx, dy = a.y - cy;
            d = dx * dx + dy * dy;
            hit = d <= r2;
        }
        const unsigned m = __ballot_sync(FULL, hit);
        if (lane == 0) s_cnt[warp] = __popc(m);
        __syncthreads();
        int base = s_base;
        for (int w = 0; w < warp; w++) base += s_cnt[w];
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < cap) { out_idx[pos] = i; scratch_d2[pos] = d; }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < nw; w++) t += s_cnt[w]; s_base += t; }
        __syncthreads();
    }
    const int count = s_base;
    if (threadIdx.x == 0) *out_n = count;
    // `.index()`: the first LIST position with the same value.  A node before the first hit with an equal d2 is a hit
    // itself (d2 <= r2), so the first equal entry among the hits is the first equal entry of the whole list.
    const int m = count < cap ? count : cap;
    // (in place: an entry that changes is not a first occurrence, and only first occurrences are read)
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double dk = scratch_d2[k];
        for (int j = 0; j < k; j++)
            if (scratch_d2[j] == dk) { out_idx[k] = out_idx[j]; break; }
    }
}

int launch_steer_points(long long n_req, const double *from_xy, const double *to_xy, const double *extend, double extend_all,
                        double res, int pt_cap, double *points, int32_t *n_points, cudaStream_t s) {
    const int threads = 128;
    steer_points_kernel<<<(unsigned)((n_req + threads - 1) / threads), threads, 0, s>>>(
        n_req, reinterpret_cast<const double2 *>(from_xy), reinterpret_cast<const double2 *>(to_xy), extend, extend_all, res, pt_cap,
        reinterpret_cast<double2 *>(points), n_points);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "steer_points_kernel launch");
    return RRTK_OK;
}

int launch_points_collide(int n_req, const double *points, const int32_t *n_points, int pt_cap, const int32_t *obs_set,
                          const double *obstacles, int obs_stride, const int32_t *n_obs, uint8_t *free_flag, cudaStream_t s) {
    const int warps = 4;
    points_collide_kernel<<<(unsigned)((n_req + warps - 1) / warps), warps * 32, 0, s>>>(
        n_req, reinterpret_cast<const double2 *>(points), n_points, pt_cap, obs_set, reinterpret_cast<const double4 *>(obstacles),
        obs_stride, n_obs, free_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "points_collide_kernel launch");
    return RRTK_OK;
}

int launch_nearest_f64(const double *xy, long long n, const double *samples, int n_samples, int32_t *out_idx, double *out_d2,
                       cudaStream_t s) {
    nearest_f64_kernel<<<(unsigned)n_samples, 256, 0, s>>>(reinterpret_cast<const double2 *>(xy), n,
                                                           reinterpret_cast<const double2 *>(samples), out_idx, out_d2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "nearest_f64_kernel launch");
    return RRTK_OK;
}

int launch_near_f64(const double *xy, int n, double cx, double cy, double r2, int32_t *out_idx, double *scratch_d2, int cap,
                    int32_t *out_n, cudaStream_t s) {
    near_f64_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const double2 *>(xy), n, cx, cy, r2, out_idx, scratch_d2, cap, out_n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "near_f64_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
