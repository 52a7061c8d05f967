// rrtk_closed_loop.cu -- the feasibility filter of Closed-loop RRT* (rrt_10:1215-1582) for P candidate courses at once:
// extend_path (:1432-1447), set_stop_point (:1375-1419), closed_loop_prediction (:1307-1372: unicycle `update` :1224-1232,
// pure_pursuit_control :1255-1283, calc_target_index :1285-1304, PIDControl :1243-1252) and the checks of
// check_tracking_path_is_feasible (:1521-1559).  The module constants are rrt_10:1592-1607.
//
// One warp per course.  A simulation step is a serial chain of correctly rounded leaf functions (crmath.h), evaluated
// uniformly by the warp; what is parallel inside a step is calc_target_index's scan over the course points:
//   pass 1  min of the squared distances (plain FP64, lanes strided over the points, shuffle reduction);
//   pass 2  only points within 1e-12 (relative) of that minimum take the exact hypot -- np.hypot's argmin is decided
//           among those, lowest index first, exactly as np.argmin over the rounded distances does.
// The set-up (course copy, per-segment direction flags, segment lengths) and the final collision test run one lane per
// point.  FP64, reference operation order, -fmad=false.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_dubins.cuh"

namespace rrtk {

constexpr int CL_WARPS_PER_CTA = 4;
constexpr int CL_EXTEND = 6;  // int(Lf / 0.1) + 1 points appended by extend_path

// calc_target_index (:1285-1304): (index after the look-ahead walk, min distance)
static __device__ __forceinline__ int cl_target_index(double sx, double sy, const double *cx, const double *cy, int n, int lane,
                                                      double *mindis) {
    double m2 = CUDART_INF;
#pragma unroll 2
    for (int i = lane; i < n; i += 32) {
        const double dx = sx - cx[i], dy = sy - cy[i];
        const double d2 = dx * dx + dy * dy;
        m2 = d2 < m2 ? d2 : m2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(FULL, m2, o);
        m2 = t < m2 ? t : m2;
    }
    const double lim = m2 + m2 * 1e-12 + 1e-300;
    double bh = CUDART_INF;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const double dx = sx - cx[i], dy = sy - cy[i];
        if (dx * dx + dy * dy <= lim) {
            const double h = crm_hypot(dx, dy);
            if (h < bh) { bh = h; bi = i; }
        }
    }
    warp_argmin(bh, bi);
    *mindis = bh;
    int ind = bi;
    double le = 0.0;
    while (0.5 > le && ind + 1 < n) {
        le += crm_hypot(cx[ind + 1] - cx[ind], cy[ind + 1] - cy[ind]);
        ind++;
    }
    return ind;
}

__global__ void __launch_bounds__(CL_WARPS_PER_CTA * 32)
closed_loop_kernel(rrtk_closed_loop_params p, const double *__restrict__ course_all, const int32_t *__restrict__ n_course_arr,
                   const double4 *__restrict__ obstacles, const int32_t *__restrict__ obs_offset,
                   const int32_t *__restrict__ n_obs_arr, double *work_all, double *traj_all, int32_t *n_traj_out,
                   int32_t *bits_out) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * CL_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (k >= p.n_courses) return;
    const int nc = n_course_arr[k], n = nc + CL_EXTEND;
    const int wcap = p.course_cap + CL_EXTEND;
    const double *course = course_all + (size_t)k * p.course_cap * 3;
    double *cx = work_all + (size_t)k * wcap * 4, *cy = cx + wcap, *cyaw = cy + wcap, *sp = cyaw + wcap;
    double *traj = traj_all + (size_t)k * p.traj_cap * 7;
    const double4 *obs = obstacles + obs_offset[k];
    const int n_obs = n_obs_arr[k];
    const double DT = 0.05, WB = 0.9, ACC = 5.0, KP = 2.0, LF = 0.5, TMAX = 100.0, GOAL_DIS = 0.5, STOP = 0.5;
    const double steer_max = 40.0 * (D_PI / 180.0);  // np.deg2rad(40.0)

    for (int i = lane; i < nc; i += 32) { cx[i] = course[3 * i]; cy[i] = course[3 * i + 1]; cyaw[i] = course[3 * i + 2]; }
    __syncwarp();
    const double gx = cx[nc - 1], gy = cy[nc - 1], gyaw = cyaw[nc - 1];
    if (lane == 0) {  // extend_path
        const double md = crm_atan2(cy[nc - 1] - cy[nc - 3], cx[nc - 1] - cx[nc - 3]);
        const double idl = fabs(md - gyaw) >= D_PI / 2.0 ? -0.1 : 0.1;
        double s, c;
        sincos_cr(gyaw, &s, &c);
        for (int i = nc; i < n; i++) { cx[i] = cx[i - 1] + idl * c; cy[i] = cy[i - 1] + idl * s; cyaw[i] = gyaw; }
    }
    __syncwarp();
    // set_stop_point: direction flag of every segment in parallel (1 = backward, +2 = zero-length), then the serial
    // forward / backward state machine; the segment lengths (for origin_travel) replace cyaw afterwards
    for (int i = lane; i < n - 1; i += 32) {
        const double dx = cx[i + 1] - cx[i], dy = cy[i + 1] - cy[i];
        const int back = fabs(crm_atan2(dy, dx) - cyaw[i]) >= D_PI / 2.0 ? 1 : 0;
        sp[i] = (double)(back | ((dx == 0.0 && dy == 0.0) ? 2 : 0));
    }
    __syncwarp();
    for (int i = lane; i < n - 1; i += 32) cyaw[i] = crm_hypot(cx[i + 1] - cx[i], cy[i + 1] - cy[i]);  // np.hypot(np.diff ..)
    __syncwarp();
    double origin = 0.0;
    if (lane == 0) {
        bool forward = true;
        int back = 0;
        for (int i = 0; i < n - 1; i++) {
            const int f = (int)sp[i];
            back = f & 1;
            double v = p.target_speed;
            if (!(f & 2)) {
                v = back ? -p.target_speed : p.target_speed;
                if (back && forward) { v = 0.0; forward = false; }
                else if (!back && !forward) { v = 0.0; forward = true; }
            }
            sp[i] = v;
            origin = origin + cyaw[i];
        }
        sp[0] = 0.0;
        sp[n - 1] = back ? -STOP : STOP;
    }
    origin = __shfl_sync(FULL, origin, 0);
    __syncwarp();

    // closed_loop_prediction
    double sx = -0.0, sy = -0.0, syaw = 0.0, sv = 0.0, time = 0.0, dis = 0.0, travel = 0.0, last_yaw = 0.0;
    int cnt = 1, bits = RRTK_CL_NOT_REACHED;
    bool overflow = false;
    const double maxdis = 0.5, dcap = maxdis - 0.1;
    if (lane == 0) { traj[0] = sx; traj[1] = sy; traj[2] = angle_mod_pi(syaw); traj[3] = sv; traj[4] = 0.0; traj[5] = 0.0; traj[6] = 0.0; }
    last_yaw = angle_mod_pi(syaw);
    int target_ind = cl_target_index(sx, sy, cx, cy, n, lane, &dis);
#pragma unroll 1
    while (TMAX >= time) {
        int ind = cl_target_index(sx, sy, cx, cy, n, lane, &dis);  // pure_pursuit_control
        if (target_ind >= ind) ind = target_ind;
        double tx, ty;
        if (ind < n) { tx = cx[ind]; ty = cy[ind]; }
        else { tx = cx[n - 1]; ty = cy[n - 1]; ind = n - 1; }
        double alpha = crm_atan2(ty - sy, tx - sx) - syaw;
        if (sv <= 0.0) alpha = D_PI - alpha;
        double di = crm_atan2(2.0 * WB * crm_sin(alpha) / LF, 1.0);
        if (di > steer_max) di = steer_max;
        else if (di < -steer_max) di = -steer_max;
        target_ind = ind;
        double ts = sp[target_ind];
        ts = ts * (maxdis - (dcap < dis ? dcap : dis)) / maxdis;
        double ai = KP * (ts - sv);  // PIDControl
        if (ai > ACC) ai = ACC;
        else if (ai < -ACC) ai = -ACC;
        double s, c;
        sincos_cr(syaw, &s, &c);  // update
        const double nx = sx + sv * c * DT, ny = sy + sv * s * DT;
        const double nyaw = angle_mod_pi(syaw + sv / WB * crm_tan(di) * DT);
        sv = sv + ai * DT;
        sx = nx; sy = ny; syaw = nyaw;
        if (fabs(sv) <= STOP && target_ind <= n - 2) target_ind++;
        time = time + DT;
        if (crm_hypot(sx - gx, sy - gy) <= GOAL_DIS) { bits = 0; break; }
        if (cnt >= p.traj_cap) { overflow = true; break; }
        last_yaw = angle_mod_pi(syaw);  // the extra angle_mod of :1533
        if (lane == 0) {
            double *r = traj + 7 * (size_t)cnt;
            r[0] = sx; r[1] = sy; r[2] = last_yaw; r[3] = sv; r[4] = time; r[5] = ai; r[6] = di;
        }
        travel = travel + fabs(sv);
        cnt++;
    }
    // check_tracking_path_is_feasible
    if (fabs(last_yaw - gyaw) >= p.yaw_th * 10.0) bits |= RRTK_CL_BAD_ANGLE;
    if ((DT * travel) / origin >= p.invalid_travel_ratio) bits |= RRTK_CL_TOO_LONG;
    __syncwarp();
    bool hit = false;
    for (int i = lane; i < cnt; i += 32) {
        const double px = traj[7 * (size_t)i], py = traj[7 * (size_t)i + 1];
        for (int o = 0; o < n_obs && !hit; o++) {
            const double4 ob = obs[o];
            const double ex = ob.x - px, ey = ob.y - py;
            if (ex * ex + ey * ey <= ob.w) hit = true;
        }
    }
    if (__any_sync(FULL, hit)) bits |= RRTK_CL_COLLISION;
    if (overflow) bits |= RRTK_CL_TRAJ_OVERFLOW;
    if (lane == 0) { n_traj_out[k] = cnt; bits_out[k] = bits; }
}

int launch_closed_loop(const rrtk_closed_loop_params &p, const double *course, const int32_t *n_course, const double *obstacles,
                       const int32_t *obs_offset, const int32_t *n_obs, double *work, double *traj, int32_t *n_traj,
                       int32_t *bits, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.n_courses + CL_WARPS_PER_CTA - 1) / CL_WARPS_PER_CTA);
    closed_loop_kernel<<<grid, CL_WARPS_PER_CTA * 32, 0, s>>>(p, course, n_course, reinterpret_cast<const double4 *>(obstacles),
                                                              obs_offset, n_obs, work, traj, n_traj, bits);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "closed_loop_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
