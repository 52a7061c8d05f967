// rrtk_rs.cu -- batched Reeds-Shepp steering (reeds_shepp_path_planning rs00:496-515 == rrt_06:1426-1437; SURVEY.md 8f
// rank 3): the 12 path functions (rs00:166-363) under the 4 symmetries of generate_path (:366-428) = 48 candidate
// words, set_path's order-dependent de-duplication (:141-160), the first shortest path, its sampled course
// (generate_local_course / interpolate :431-470 with np.arange's  start + i * step  distances) rotated to the world
// frame (calc_paths :473-493), and the sampled collision test of RRT*-Reeds-Shepp's check_collision over the course.
//
// Eight requests per warp and round.  Their 48 words each are shared out over the lanes (rs_pick_coop, rrtk_rs.cuh): the
// atan2 / acos / asin / sin / cos are the correctly rounded crmath.h functions, so the lengths -- and with them every
// `>= 0`, `<= step_size` and minimum decision -- are platform independent; the insertion logic runs in the reference's
// order per request.  The warp then takes the courses one after the other, the points of a segment spread over the lanes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_dubins.cuh"
#include "rrtk_rs.cuh"

namespace rrtk {

constexpr int RS_WARPS = 4;

__global__ void __launch_bounds__(RS_WARPS * 32)
rs_steer_kernel(int n_req, double maxc, double step_size, const double *__restrict__ from3, const double *__restrict__ to3,
                const int32_t *__restrict__ obs_set, const double4 *__restrict__ obstacles, int obs_stride,
                const int32_t *__restrict__ n_obs_arr, int32_t *types_out, double *lengths_out, double *L_out,
                int32_t *n_paths_out, double *end_out, int32_t *n_pts_out, uint8_t *free_out, double *pts_out, int max_pts) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const double step = step_size * maxc;
    // eight requests per warp and round: their 48 words each are shared out over the lanes (rs_pick_coop: twelve uniform
    // family passes for the eight of them instead of twelve four-lane passes per request); then the warp takes the eight
    // courses one after the other, the points of a segment spread over the lanes
    for (int r0 = 8 * warp; r0 < n_req; r0 += 8 * n_warps) {
        const int nact = n_req - r0 < 8 ? n_req - r0 : 8;
        const int rl = r0 + (lane < nact ? lane : 0);
        const double l_sx = from3[3 * rl], l_sy = from3[3 * rl + 1], l_syaw = from3[3 * rl + 2];
        const double l_gx = to3[3 * rl], l_gy = to3[3 * rl + 1], l_gyaw = to3[3 * rl + 2];
        const RsPick pick = rs_pick_coop(nact, lane, l_sx, l_sy, l_syaw, l_gx, l_gy, l_gyaw, maxc, step);
#pragma unroll 1
      for (int e8 = 0; e8 < nact; e8++) {
        const int r = r0 + e8;
        const double sx = __shfl_sync(FULL, l_sx, e8), sy = __shfl_sync(FULL, l_sy, e8), syaw = __shfl_sync(FULL, l_syaw, e8);
        const int best = __shfl_sync(FULL, pick.best, e8), n_ins = __shfl_sync(FULL, pick.n_ins, e8);
        double bd[5];
#pragma unroll
        for (int i = 0; i < 5; i++) bd[i] = __shfl_sync(FULL, pick.d[i], e8);   // (time flip applied)
        if (best < 0) {
            if (lane == 0) {
                n_paths_out[r] = 0; n_pts_out[r] = 0; free_out[r] = 0; L_out[r] = 0.0;
                for (int i = 0; i < 5; i++) { types_out[5 * r + i] = -1; lengths_out[5 * r + i] = 0.0; }
            }
            continue;
        }
        const int f = best >> 2, k = best & 3, n = RS_N[f];
        double tot = 0.0;
        for (int i = 0; i < n; i++) tot += fabs(bd[i]);
        const double best_L = fabs(tot / maxc);   // abs(p.L): the time flip only negates
        const int set = obs_set ? obs_set[r] : 0;
        const double4 *obs = obstacles + (size_t)set * obs_stride;
        const int n_obs = n_obs_arr ? n_obs_arr[set] : 0;
        double *pts = pts_out ? pts_out + (size_t)r * max_pts * 4 : nullptr;
        double sm0, cm0;
        sincos_cr(-syaw, &sm0, &cm0);
        bool hit = false;
        int np = 0;
        double ox = 0.0, oy = 0.0, oyaw = 0.0, lastx = sx, lasty = sy, lastyaw = syaw;
        for (int i = 0; i < n; i++) {
            const double length = bd[i];
            const int t0 = RS_T[f][i];
            const int type = (k >= 2 && t0 != 1) ? 2 - t0 : t0;  // reflect
            const double dd = length >= 0.0 ? step : -step;
            long long na = length != 0.0 ? (long long)ceil((length - 0.0) / dd) : 0;   // len(np.arange(0.0, length, d_dist))
            if (na < 0) na = 0;
            double so, co, sm, cm;
            sincos_cr(oyaw, &so, &co);
            sincos_cr(-oyaw, &sm, &cm);
            const double dirn = length > 0.0 ? 1.0 : -1.0;
            for (long long j = lane; j <= na; j += 32) {
                const double dist = j < na ? 0.0 + (double)j * dd : length;
                double lx, ly, lyaw;
                rs_interp(dist, type, maxc, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);
                const double wx = cm0 * lx + sm0 * ly + sx, wy = -sm0 * lx + cm0 * ly + sy;   // rs00:481-485
                const long long idx = np + j;
                if (pts && idx < max_pts) {
                    pts[4 * idx] = wx; pts[4 * idx + 1] = wy; pts[4 * idx + 2] = angle_mod_pi(lyaw + syaw); pts[4 * idx + 3] = dirn;
                }
                for (int o = 0; o < n_obs && !hit; o++) {
                    const double4 ob = obs[o];
                    const double ex = ob.x - wx, ey = ob.y - wy;
                    if (ex * ex + ey * ey <= ob.w) hit = true;
                }
            }
            double lx, ly, lyaw;   // the segment's last point is the next origin (uniform)
            rs_interp(length, type, maxc, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);
            ox = lx; oy = ly; oyaw = lyaw;
            lastx = cm0 * lx + sm0 * ly + sx; lasty = -sm0 * lx + cm0 * ly + sy; lastyaw = angle_mod_pi(lyaw + syaw);
            np += (int)(na + 1);
        }
        const bool any_hit = __ballot_sync(FULL, hit) != 0u;
        if (lane == 0) {
            n_paths_out[r] = n_ins; n_pts_out[r] = np; free_out[r] = any_hit ? 0 : 1; L_out[r] = best_L;
            for (int i = 0; i < 5; i++) {
                const double length = i < n ? bd[i] : 0.0;
                const int t0 = i < n ? RS_T[f][i] : -1;
                types_out[5 * r + i] = i < n ? ((k >= 2 && t0 != 1) ? 2 - t0 : t0) : -1;
                lengths_out[5 * r + i] = i < n ? length / maxc : 0.0;
            }
            end_out[3 * r] = lastx; end_out[3 * r + 1] = lasty; end_out[3 * r + 2] = lastyaw;
        }
      }
    }
}

int launch_rs_steer(int n_req, double maxc, double step_size, const double *from3, const double *to3, const int32_t *obs_set,
                    const double *obstacles, int obs_stride, const int32_t *n_obs, int32_t *types, double *lengths, double *L,
                    int32_t *n_paths, double *end, int32_t *n_pts, uint8_t *free_flag, double *pts, int max_pts, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = ((long long)n_req + 8 * RS_WARPS - 1) / (8 * RS_WARPS);
    long long grid = (long long)sms * 8;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    rs_steer_kernel<<<(unsigned)grid, RS_WARPS * 32, 0, s>>>(n_req, maxc, step_size, from3, to3, obs_set,
                                                            reinterpret_cast<const double4 *>(obstacles), obs_stride, n_obs,
                                                            types, lengths, L, n_paths, end, n_pts, free_flag, pts, max_pts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rs_steer_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
