// rrtk_rrtstar_common.cuh -- device functions shared by the two executions of the RRT / RRT* loop (rrt_04:1036-1084):
// the warp-per-query kernel (rrtk_rrtstar.cu) and the CTA-per-query kernel (rrtk_rrtstar_cta.cu).  steer, the sampled
// collision test (exact and error-banded forms), the play-area test and the samplers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"

namespace rrtk {

struct Steer {
    double ex, ey;    // end point
    double stx, sty;  // step vector = res * (cos, sin)
    double d;         // hypot(to - from): what calc_distance_and_angle returns (rrt_04:1232-1238)
    int n;            // number of accumulated steps
    bool snap;        // final point snapped to the target
};

// steer (rrt_04:1086-1115)
static __device__ __noinline__ Steer steer(double fx, double fy, double tx, double ty, double extend,
                                    double res) {
    Steer st;
    double dx = tx - fx, dy = ty - fy;
    double d = crm_hypot(dx, dy);
    st.d = d;
    double s, c;
    (void)crm_atan2_sincos(dy, dx, &s, &c);
    if (extend > d) extend = d;
    double q = floor(extend / res);
    int n = q < 2.0e9 ? (int)q : 2000000000;
    st.stx = res * c;
    st.sty = res * s;
    double x = fx, y = fy;
    for (int k = 0; k < n; k++) {
        x += st.stx;
        y += st.sty;
    }
    double d2 = crm_hypot(tx - x, ty - y);
    st.snap = d2 <= res;
    if (st.snap) {
        x = tx;
        y = ty;
    }
    st.ex = x;
    st.ey = y;
    st.n = n;
    return st;
}

// check_collision (rrt_04:1216-1230) of one edge by ONE lane: any path point within any circle.
static __device__ __noinline__ bool edge_free_lane(double fx, double fy, const Steer &st, double tx,
                                               double ty, const ObsList &L) {
    double x = fx, y = fy;
    for (int k = 0;; k++) {
        for (int j = 0; j < L.m; j++) {
            double dx = L.ox[j * L.stride] - x, dy = L.oy[j * L.stride] - y;
            if (dx * dx + dy * dy <= L.r2[j * L.stride]) return false;
        }
        if (k == st.n) break;
        x += st.stx;
        y += st.sty;
    }
    if (st.snap) {
        for (int j = 0; j < L.m; j++) {
            double dx = L.ox[j * L.stride] - tx, dy = L.oy[j * L.stride] - ty;
            if (dx * dx + dy * dy <= L.r2[j * L.stride]) return false;
        }
    }
    return true;
}

// Cheap verdict of the edge f -> t steered with extend_length = inf (choose_parent / rewire candidates), WITHOUT the
// correctly rounded atan2/cos/sin: the reference's path points are f, f + k * res * (cos, sin)(theta) (k = 1..n,
// accumulated) and -- when the last one lies within `res` of t -- t itself (rrt_04:1099-1113).  n = floor(d / res)
// is computed exactly as steer does; the first and last points are f and t exactly; the intermediate points are
// reproduced to within eps_pos with the direction (t - f) / d, so a point-circle test whose margin exceeds the
// error band has the reference's verdict.  Returns 1 = free and snapped (end point == t), 0 = blocked,
// -1 = too close to call (a test inside the band, or the snap decision within 1e-9 of its threshold): the caller
// runs the exact steer + edge_free_lane.  `extend` = steer's extend_length (inf for choose_parent / rewire,
// expand_dis for the first edge); obstacles j0, j0 + jstep, ... are tested (lanes can split them).
static __device__ __noinline__ int edge_verdict_fast(double fx, double fy, double tx, double ty, double d, double extend,
                                              double res, const ObsList &L, int j0, int jstep) {
    if (!(d > 0.0)) return -1;
    const double q = floor((extend > d ? d : extend) / res);   // steer's n_expand (rrt_04:1096-1099)
    if (!(q < 1.0e6)) return -1;
    const int n = (int)q;
    const double rem = d - q * res;                    // distance left after n steps
    if (!(rem <= res * (1.0 - 1e-9))) return -1;       // snap (d2 <= res, rrt_04:1107) must be certain
    const double inv = res / d;
    const double ux = (tx - fx) * inv, uy = (ty - fy) * inv;
    const double e4 = 4.0 * ((double)(n + 8) * 2.3e-16 * (fabs(fx) + fabs(fy) + fabs(tx) + fabs(ty) + 1.0)) + 4e-15;
    bool unsure = false;
    const double wx = tx - fx, wy = ty - fy, invl2 = 1.0 / (d * d);
    for (int j = j0; j < L.m; j += jstep) {
        const double ox = L.ox[j * L.stride], oy = L.oy[j * L.stride], r2 = L.r2[j * L.stride];
        double dx = ox - fx, dy = oy - fy;             // first point: f itself, exact test
        if (dx * dx + dy * dy <= r2) return 0;
        const double bj = (2.02 + 2.02 * r2) * e4;     // >= (2 + dd + r2) * e4 wherever |dd - r2| is that small
        // every path point lies on the segment f-t (to within eps_pos): a circle farther than its radius (plus
        // the band) from the segment cannot contain one
        double sp = (dx * wx + dy * wy) * invl2;
        sp = sp < 0.0 ? 0.0 : (sp > 1.0 ? 1.0 : sp);
        const double px = dx - sp * wx, py = dy - sp * wy;
        if (px * px + py * py - r2 > bj + 1e-9 * (1.0 + r2)) continue;
        dx = ox - tx; dy = oy - ty;                    // last point: t itself (snapped), exact test
        if (dx * dx + dy * dy <= r2) return 0;
        double x = fx, y = fy;
        for (int k = 1; k <= n; k++) {
            x += ux; y += uy;
            dx = ox - x; dy = oy - y;
            const double t = dx * dx + dy * dy - r2;
            if (t <= bj) {
                if (t <= -bj) return 0;                // certainly inside: blocked whatever the others say
                unsure = true;
            }
        }
    }
    return unsure ? -1 : 1;
}

// the same verdict computed by the whole warp (lanes split the obstacles); uniform result
static __device__ __noinline__ bool edge_free_warp(double fx, double fy, const Steer &st, double tx,
                                               double ty, const ObsList &L, int lane) {
    bool hit = false;
    for (int j = lane; j < L.m && !hit; j += 32) {
        double ox = L.ox[j * L.stride], oy = L.oy[j * L.stride], r2 = L.r2[j * L.stride];
        double x = fx, y = fy;
        for (int k = 0;; k++) {
            double dx = ox - x, dy = oy - y;
            if (dx * dx + dy * dy <= r2) { hit = true; break; }
            if (k == st.n) break;
            x += st.stx;
            y += st.sty;
        }
        if (!hit && st.snap) {
            double dx = ox - tx, dy = oy - ty;
            if (dx * dx + dy * dy <= r2) hit = true;
        }
    }
    return __ballot_sync(FULL, hit) == 0u;
}

static __device__ __forceinline__ bool inside_play(const rrtk_rrtstar_params &p, double x, double y) {
    if (!p.has_play_area) return true;  // rrt_04:1207-1208
    return !(x < p.play_area[0] || x > p.play_area[1] || y < p.play_area[2] || y > p.play_area[3]);
}

struct Sample { double x, y; };

// Sobol state of one query: the current point (30-bit integers) and its index; advanced with the
// Antonov-Saleev update point(n+1) = point(n) ^ V[lowest zero bit of n] (what i4_sobol does, rrt_04:448-452)
struct SobolState {
    int64_t n;
    uint32_t q0, q1;
};

// get_random_node / get_random_node_sobol (rrt_04:1132-1153) with a counter-based coin
static __device__ __forceinline__ Sample draw_sample(const rrtk_rrtstar_params &p, int q, int it, int it_key, double gx,
                                              double gy, const double2 *stream, SobolState &sob) {
    Sample s;
    if (p.sampler == RRTK_SAMPLER_STREAM) {
        double2 v = stream[it];
        s.x = v.x; s.y = v.y;
        return s;
    }
    uint64_t k0 = rng_key(p.seed, (uint64_t)q, (uint64_t)it_key);   // it_key = iteration counter of the whole run
    int coin = (int)(splitmix64(k0) % 101ull);  // random.randint(0, 100)
    if (coin > p.goal_sample_rate) {
        double w = p.max_rand - p.min_rand;
        if (p.sampler == RRTK_SAMPLER_SOBOL) {
            const double recipd = 1.0 / 1073741824.0;
            s.x = p.min_rand + ((double)sob.q0 * recipd) * w;
            s.y = p.min_rand + ((double)sob.q1 * recipd) * w;
            int c = __ffsll(~sob.n) - 1;  // lowest zero bit of the index just used
            if (c < SOBOL_BITS) { sob.q0 ^= c_sobol.v[0][c]; sob.q1 ^= c_sobol.v[1][c]; }
            sob.n++;
        } else {
            s.x = p.min_rand + w * u01(splitmix64(k0 + 1));
            s.y = p.min_rand + w * u01(splitmix64(k0 + 2));
        }
    } else {
        s.x = gx; s.y = gy;
    }
    return s;
}

// search_best_goal_node (rrt_04:1284-1312).  Returns the goal node index or -1.  Uniform result.
static __device__ __noinline__ int best_goal(const rrtk_rrtstar_params &p, int n, const double2 *xy,
                                      const double *cost, double gx, double gy, const ObsList &G,
                                      int *near_idx, double *nd, int near_cap, int lane, bool &overflow) {
    // candidates: dist <= expand_dis, each mapped to the first index with the same dist
    int count = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool hit = false;
        double d = 0.0;
        if (i < n) {
            double2 a = xy[i];
            d = crm_hypot(a.x - gx, a.y - gy);
            hit = d <= p.expand_dis;
        }
        unsigned mask = __ballot_sync(FULL, hit);
        int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < near_cap) { near_idx[pos] = i; nd[pos] = d; }
        count += __popc(mask);
    }
    __syncwarp();
    if (count > near_cap) { overflow = true; count = near_cap; }
    double best_c = CUDART_INF;
    int best_k = 0x7fffffff;
    for (int k = lane; k < count; k += 32) {
        double dk = nd[k];
        int f = k;
        for (int j = 0; j < k; j++)
            if (nd[j] == dk) { f = j; break; }
        int i = near_idx[f];
        double2 a = xy[i];
        Steer st = steer(a.x, a.y, gx, gy, CUDART_INF, p.path_resolution);
        bool ok = edge_free_lane(a.x, a.y, st, gx, gy, G) && inside_play(p, st.ex, st.ey);
        if (ok) {
            double c = cost[i] + crm_hypot(a.x - gx, a.y - gy);
            // first minimum over the candidate list; equal costs keep the earlier list entry
            if (c < best_c) { best_c = c; best_k = k; }
        }
    }
    // reduce over lanes: min cost, ties -> smaller list position
    warp_argmin(best_c, best_k);
    if (best_k == 0x7fffffff) return -1;
    // map the list position back to the node index (first index with the same distance)
    double dk = nd[best_k];
    int f = best_k;
    for (int j = 0; j < best_k; j++)
        if (nd[j] == dk) { f = j; break; }
    return near_idx[f];
}

}  // namespace rrtk
