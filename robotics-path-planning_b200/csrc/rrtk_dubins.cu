// rrtk_dubins.cu -- batched Dubins steering (plan_dubins_path rrt_05:1021-1109 == dub00): shortest of the six
// words LSL,RSR,LSR,RSL,RLR,LRL (_dubins_path_planning_from_origin :1201-1229, word solvers :1125-1198), the
// sampled course (_generate_local_course / _interpolate :1232-1278, step 0.1 in normalised units) rotated to
// the world frame, and the reference's sampled collision test (check_collision rrt_05:1625-1638) over the
// course points -- what RRT*-Dubins' `steer` + `check_collision` do for one edge (rrt_05:1458-1479).
//
// One warp per steering request.  Lanes 0..5 solve one word each (the expensive atan2/acos run in parallel),
// a shuffle first-min picks the word in the reference's order; the course points of each segment are spread
// over the 32 lanes (one correctly rounded sin/cos pair per curve point), transformed and tested against the
// request's obstacle set.  FP64, reference operation order, crmath.h leaf functions, numpy's fma matmul.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_dubins.cuh"

namespace rrtk {

extern "C" __global__ void __launch_bounds__(128)
dubins_steer_kernel(int n_req, double kappa, double step, const double *__restrict__ from3,
                    const double *__restrict__ to3, const int32_t *__restrict__ obs_set,
                    const double4 *__restrict__ obstacles, int obs_stride, const int32_t *__restrict__ n_obs_arr,
                    int32_t *mode_out, double *lengths_out, double *end_out, int32_t *n_pts_out,
                    uint8_t *free_out, double *pts_out, int max_pts) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < n_req; r += n_warps) {
        const double s_x = from3[3 * r], s_y = from3[3 * r + 1], s_yaw = from3[3 * r + 2];
        const double g_x = to3[3 * r], g_y = to3[3 * r + 1], g_yaw = to3[3 * r + 2];
        double c, s;
        rot2d(s_yaw, &c, &s);
        const double vx = g_x - s_x, vy = g_y - s_y;
        const double lgx = fma(vy, s, vx * c), lgy = fma(vy, c, vx * -s);  // numpy (2,)@(2,2), reference platform
        const double lgyaw = g_yaw - s_yaw;
        const double d = crm_hypot(lgx, lgy) * kappa;
        const double theta = mod2pi(crm_atan2(lgy, lgx));
        const double alpha = mod2pi(-theta), beta = mod2pi(lgyaw - theta);
        // six words, one per lane; first minimum of |d1|+|d2|+|d3| in _PATH_TYPE_MAP order (:1214-1221)
        double w[3] = {0.0, 0.0, 0.0};
        double cost = CUDART_INF;
        int bi = 0x7fffffff;
        const DubTrig trig = dubins_trig(alpha, beta);
        if (lane < 6 && dubins_word<true>(lane, alpha, beta, d, trig, w)) {
            cost = fabs(w[0]) + fabs(w[1]) + fabs(w[2]);
            bi = lane;
        }
        warp_argmin(cost, bi);
        if (bi == 0x7fffffff) {  // no word is feasible
            if (lane == 0) { mode_out[r] = -1; n_pts_out[r] = 0; free_out[r] = 0; }
            continue;
        }
        double len[3];
#pragma unroll
        for (int k = 0; k < 3; k++) len[k] = __shfl_sync(FULL, w[k], bi);
        double c2, s2;
        rot2d(-s_yaw, &c2, &s2);
        const int set = obs_set ? obs_set[r] : 0;
        const double4 *obs = obstacles + (size_t)set * obs_stride;
        const int n_obs = n_obs_arr ? n_obs_arr[set] : 0;
        double *pts = pts_out ? pts_out + (size_t)r * max_pts * 3 : nullptr;
        bool hit = false;
        int np = 0;
        double lx = 0.0, ly = 0.0, lyaw = 0.0;
        // emit a local point: world transform (numpy (N,2)@(2,2) + offset), optional store, collision test
        auto emit = [&](double px, double py, double pyaw, int index) {
            const double wx = fma(py, s2, px * c2) + s_x;
            const double wy = fma(py, c2, px * -s2) + s_y;
            if (pts && index < max_pts) {
                pts[3 * index] = wx; pts[3 * index + 1] = wy; pts[3 * index + 2] = angle_mod_pi(pyaw + s_yaw);
            }
            for (int o = 0; o < n_obs && !hit; o++) {
                const double4 ob = obs[o];
                const double dx = ob.x - wx, dy = ob.y - wy;
                if (dx * dx + dy * dy <= ob.w) hit = true;
            }
        };
        if (lane == 0) emit(lx, ly, lyaw, 0);
        np = 1;
        for (int k = 0; k < 3; k++) {
            const double length = len[k];
            if (length == 0.0) continue;
            const int type = seg_type(bi, k);
            const double ox = lx, oy = ly, oyaw = lyaw;
            double so, co, sm, cm;
            sincos_cr(oyaw, &so, &co);
            sincos_cr(-oyaw, &sm, &cm);
            // interior points: cur = step, step + step, ... while |cur + step| <= |length|
            int cnt = 0;
            {
                double cur = step;
                while (fabs(cur + step) <= fabs(length)) { cnt++; cur += step; }
            }
            double cur = step;
            for (int t = 0; t < lane; t++) cur += step;  // lane's first point: `lane` sequential additions
            for (int j = lane; j < cnt; j += 32) {
                double x, y, yaw;
                interp(cur, type, kappa, ox, oy, oyaw, so, co, sm, cm, &x, &y, &yaw);
                emit(x, y, yaw, np + j);
#pragma unroll 1
                for (int t = 0; t < 32; t++) cur += step;
            }
            interp(length, type, kappa, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);  // segment end (uniform)
            if (lane == 0) emit(lx, ly, lyaw, np + cnt);
            np += cnt + 1;
        }
        const bool any_hit = __ballot_sync(FULL, hit) != 0u;
        if (lane == 0) {
            mode_out[r] = bi;
            n_pts_out[r] = np;
            free_out[r] = any_hit ? 0 : 1;
            for (int k = 0; k < 3; k++) lengths_out[3 * r + k] = len[k] / kappa;
            end_out[3 * r] = fma(ly, s2, lx * c2) + s_x;
            end_out[3 * r + 1] = fma(ly, c2, lx * -s2) + s_y;
            end_out[3 * r + 2] = angle_mod_pi(lyaw + s_yaw);
        }
    }
}

int launch_dubins_steer(int n_req, double kappa, double step, const double *from3, const double *to3,
                        const int32_t *obs_set, const double *obstacles, int obs_stride, const int32_t *n_obs,
                        int32_t *mode, double *lengths, double *end, int32_t *n_pts, uint8_t *free_flag,
                        double *pts, int max_pts, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = ((long long)n_req + 3) / 4;
    long long grid = (long long)sms * 8;  // persistent-style, a multiple of the SM count
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    dubins_steer_kernel<<<(unsigned)grid, 128, 0, s>>>(n_req, kappa, step, from3, to3, obs_set,
                                                       reinterpret_cast<const double4 *>(obstacles), obs_stride,
                                                       n_obs, mode, lengths, end, n_pts, free_flag, pts, max_pts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "dubins_steer_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
