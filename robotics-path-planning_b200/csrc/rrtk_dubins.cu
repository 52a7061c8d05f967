// rrtk_dubins.cu -- batched Dubins steering (plan_dubins_path rrt_05:1021-1109 == dub00): shortest of the six
// words LSL,RSR,LSR,RSL,RLR,LRL (_dubins_path_planning_from_origin :1201-1229, word solvers :1125-1198), the
// sampled course (_generate_local_course / _interpolate :1232-1278, step 0.1 in normalised units) rotated to
// the world frame, and the reference's sampled collision test (check_collision rrt_05:1625-1638) over the
// course points -- what RRT*-Dubins' `steer` + `check_collision` do for one edge (rrt_05:1458-1479).
//
// Eight steering requests per warp and round.  Lanes 0..7 take one request's frame, angles and sines / cosines each,
// the six words of the eight are shared out over the lanes (four CSC words x 8 requests = one pass, two CCC words x 8 =
// another; a shuffle first-min picks the word in the reference's order), then the warp emits the eight courses one
// after the other: the points of each segment spread over the 32 lanes (one correctly rounded sin/cos pair per curve
// point, the running sums that place them looked up in a per-CTA table), transformed and tested against the request's
// obstacle set.  FP64, reference operation order, crmath.h leaf functions, numpy's fma matmul.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_dubins.cuh"

namespace rrtk {

constexpr int DUB_CUR_TAB = 1024;   // running sums of `step` kept per CTA (8 KB): courses up to 1024 points per segment

extern "C" __global__ void __launch_bounds__(128)
dubins_steer_kernel(int n_req, double kappa, double step, const double *__restrict__ from3,
                    const double *__restrict__ to3, const int32_t *__restrict__ obs_set,
                    const double4 *__restrict__ obstacles, int obs_stride, const int32_t *__restrict__ n_obs_arr,
                    int32_t *mode_out, double *lengths_out, double *end_out, int32_t *n_pts_out,
                    uint8_t *free_out, double *pts_out, int max_pts) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // the interior points of a segment sit at cur = step, step + step, ... (rrt_05:1107-1123: a running sum, not j * step):
    // the sums are the same for every edge of the launch, so they are formed once per CTA -- s_cur[j] = cur after j
    // additions -- instead of being counted through (~260 additions per edge) and walked up to (32 per lane and round)
    __shared__ double s_cur[DUB_CUR_TAB];
    if (threadIdx.x == 0) {
        double cur = step;
        for (int j = 0; j < DUB_CUR_TAB; j++) { s_cur[j] = cur; cur += step; }
    }
    __syncthreads();
    // eight edges per warp and round: lanes 0..7 take one edge's frame, angles and sines / cosines each, the 6 words of the
    // eight are shared out over the lanes (dubins_words_coop), then the warp emits the eight courses one after the other
    for (int r0 = 8 * warp; r0 < n_req; r0 += 8 * n_warps) {
        const int nact = n_req - r0 < 8 ? n_req - r0 : 8;
        const int rl = r0 + (lane < nact ? lane : 0);
        const double l_sx = from3[3 * rl], l_sy = from3[3 * rl + 1], l_syaw = from3[3 * rl + 2];
        double l_c = 1.0, l_s = 0.0, l_alpha = 0.0, l_beta = 0.0, l_d = 0.0;
        DubTrig l_trig;
        l_trig.sa = l_trig.ca = l_trig.sb = l_trig.cb = l_trig.cab = 0.0;
        if (lane < nact)
            dubins_front(l_sx, l_sy, l_syaw, to3[3 * rl], to3[3 * rl + 1], to3[3 * rl + 2], kappa, l_c, l_s, l_alpha, l_beta, l_d, l_trig);
        // first minimum of |d1|+|d2|+|d3| in _PATH_TYPE_MAP order (:1214-1221)
        const DubWord l_w = dubins_words_coop(nact, lane, l_alpha, l_beta, l_d, l_trig);
#pragma unroll 1
      for (int e8 = 0; e8 < nact; e8++) {
        const int r = r0 + e8;
        const int bi = __shfl_sync(FULL, l_w.bi, e8);
        if (bi == 0x7fffffff) {  // no word is feasible
            if (lane == 0) { mode_out[r] = -1; n_pts_out[r] = 0; free_out[r] = 0; }
            continue;
        }
        const double s_x = __shfl_sync(FULL, l_sx, e8), s_y = __shfl_sync(FULL, l_sy, e8), s_yaw = __shfl_sync(FULL, l_syaw, e8);
        const double c = __shfl_sync(FULL, l_c, e8), s = __shfl_sync(FULL, l_s, e8);
        double len[3];
        len[0] = __shfl_sync(FULL, l_w.l0, e8); len[1] = __shfl_sync(FULL, l_w.l1, e8); len[2] = __shfl_sync(FULL, l_w.l2, e8);
        // rot_mat_2d(-s_yaw): the correctly rounded sin / cos are odd / even bit for bit, so c2 = c and s2 = -s exactly
        const double c2 = c, s2 = -s;
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
            double so, co;
            sincos_cr(oyaw, &so, &co);
            const double sm = -so, cm = co;               // sin / cos(-oyaw), exactly
            // interior points: cur = step, step + step, ... while |cur + step| <= |length|, i.e. as many as there are sums
            // s_cur[1..] not above |length| (the sums grow): a binary search; past the table the reference's walk
            const double alen = fabs(length);
            int cnt;
            if (s_cur[DUB_CUR_TAB - 1] <= alen) {
                cnt = DUB_CUR_TAB - 1;
                double cur = s_cur[DUB_CUR_TAB - 1];
                while (fabs(cur + step) <= alen) { cnt++; cur += step; }
            } else {
                int lo = 0, hi = DUB_CUR_TAB - 1;         // s_cur[hi] > alen; the answer is the last j >= 0 with s_cur[j] <= alen, or 0
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_cur[mid] <= alen) lo = mid; else hi = mid;
                }
                cnt = lo;                                  // (s_cur[0] = step itself does not count: cnt = #{j >= 1: s_cur[j] <= alen})
            }
            for (int j = lane; j < cnt; j += 32) {
                double cur;
                if (j < DUB_CUR_TAB) cur = s_cur[j];
                else { cur = s_cur[DUB_CUR_TAB - 1]; for (int t = DUB_CUR_TAB - 1; t < j; t++) cur += step; }
                double x, y, yaw;
                interp(cur, type, kappa, ox, oy, oyaw, so, co, sm, cm, &x, &y, &yaw);
                emit(x, y, yaw, np + j);
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
}

int launch_dubins_steer(int n_req, double kappa, double step, const double *from3, const double *to3,
                        const int32_t *obs_set, const double *obstacles, int obs_stride, const int32_t *n_obs,
                        int32_t *mode, double *lengths, double *end, int32_t *n_pts, uint8_t *free_flag,
                        double *pts, int max_pts, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = ((long long)n_req + 31) / 32;   // 4 warps x 8 edges per CTA and round
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
