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
// (small structs go BY VALUE into out-of-line functions: a reference makes the caller spill the struct to local memory and
// the callee load it back -- L2 round trips in the CTA-per-query kernel, whose shared-memory carve-out leaves almost no L1)
static __device__ __noinline__ bool edge_free_lane(double fx, double fy, const Steer st, double tx,
                                               double ty, const ObsList L) {
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
// as steer computes it; the first and last points are f and t exactly; the intermediate points are reproduced to
// within eps_pos with the direction (t - f) / d, so a point-circle test whose margin exceeds the error band has the
// reference's verdict.  v: 1 = free and snapped (end point == t), 0 = blocked, -1 = too close to call (a test inside
// the band, or the snap decision within 1e-9 of its threshold): the caller runs the exact steer + edge_free_lane.
//   d        hypot(t - f): the correctly rounded value (d_exact) or any value within a few ulp of it (the decisions
//            below keep a 1e-9 margin; near a multiple of the resolution only the exact value can decide)
//   q_ext    floor(extend / res) for steer's extend_length (huge for inf), used when d >= extend (the first edge)
//   inv_res  1 / res: n = floor(d * inv_res) wherever d / res is not within 1e-9 of an integer -- one division
//            (1 / d) per edge instead of four
//   j0, jstep  circles j0, j0 + jstep, ... are tested (lanes can split them)
//   only     bit b set = test circle j0 + b * jstep (b < 64; circles beyond 64 are always tested)
// near: the circles (same bit numbering) that are NOT farther than their radius plus the band from the segment f-t.
// Every path point of the edge t -> f lies on the same segment, so the reverse edge (rewire, rrt_04:1340-1373) needs
// only those: `only = near` of a FREE forward edge, and nothing at all when it is empty.  ~0 = unknown (blocked or
// undecided before every circle was seen, or more than 64 circles).
// steer's snap onto its target is certain for an edge of length d (the head of edge_verdict_fast, for callers that
// want to know before they gather the circles)
static __device__ __forceinline__ bool snap_certain(double d, bool d_exact, double extend, double q_ext, double res,
                                                    double inv_res) {
    if (!(d > 0.0)) return false;
    double q = q_ext;
    if (extend > d) {
        q = floor(d * inv_res);
        const double rem0 = d - q * res;
        if (!(rem0 >= res * 1e-9 && rem0 <= res * (1.0 - 1e-9))) {
            if (!d_exact) return false;
            q = floor(d / res);
        }
    }
    return q < 1.0e6 && d - q * res <= res * (1.0 - 1e-9);
}

struct EdgeVerdict {
    int v;
    unsigned long long near;
};
// MASK = false: `only` is ignored and `near` is not produced (the warp-per-query kernel: lanes of a warp run the reverse
// edges in lockstep, so skipping circles in some lanes saves nothing there and the bookkeeping costs 5 % -- measured).
template <bool MASK>
static __device__ __noinline__ EdgeVerdict edge_verdict_fast(double fx, double fy, double tx, double ty, double d, bool d_exact,
                                                             double extend, double q_ext, double res, double inv_res,
                                                             const ObsList L, int j0, int jstep, unsigned long long only) {
    EdgeVerdict r;
    r.v = -1; r.near = ~0ull;
    if (!(d > 0.0)) return r;
    double q = q_ext;                                  // steer's n_expand (rrt_04:1096-1099)
    if (extend > d) {
        q = floor(d * inv_res);
        const double rem0 = d - q * res;
        if (!(rem0 >= res * 1e-9 && rem0 <= res * (1.0 - 1e-9))) {
            if (!d_exact) return r;
            q = floor(d / res);                        // the reference's own quotient decides next to a multiple of res
        }
    }
    if (!(q < 1.0e6)) return r;
    const int n = (int)q;
    const double rem = d - q * res;                    // distance left after n steps
    if (!(rem <= res * (1.0 - 1e-9))) return r;        // snap (d2 <= res, rrt_04:1107) must be certain
    const double inv_d = 1.0 / d;
    const double inv = res * inv_d;
    const double ux = (tx - fx) * inv, uy = (ty - fy) * inv;
    // position error of the reproduced points: n accumulated roundings + the direction's (res / d as two rounded
    // operations on a d that may be a few ulp off: 4 of the 12 units)
    const double e4 = 4.0 * ((double)(n + 12) * 2.3e-16 * (fabs(fx) + fabs(fy) + fabs(tx) + fabs(ty) + 1.0)) + 4e-15;
    bool unsure = false;
    unsigned long long near = 0ull;
    const double wx = tx - fx, wy = ty - fy, invl2 = inv_d * inv_d;
    int b = 0;
    for (int j = j0; j < L.m; j += jstep, b++) {
        if (MASK && b < 64 && !((only >> b) & 1ull)) continue;
        const double ox = L.ox[j * L.stride], oy = L.oy[j * L.stride], r2 = L.r2[j * L.stride];
        double dx = ox - fx, dy = oy - fy;             // first point: f itself, exact test
        if (dx * dx + dy * dy <= r2) { r.v = 0; return r; }
        const double bj = (2.02 + 2.02 * r2) * e4;     // >= (2 + dd + r2) * e4 wherever |dd - r2| is that small
        // every path point lies on the segment f-t (to within eps_pos): a circle farther than its radius (plus
        // the band) from the segment cannot contain one
        double sp = (dx * wx + dy * wy) * invl2;
        sp = sp < 0.0 ? 0.0 : (sp > 1.0 ? 1.0 : sp);
        const double px = dx - sp * wx, py = dy - sp * wy;
        if (px * px + py * py - r2 > bj + 1e-9 * (1.0 + r2)) continue;
        if (MASK) near |= b < 64 ? (1ull << b) : 0ull;
        dx = ox - tx; dy = oy - ty;                    // last point: t itself (snapped), exact test
        if (dx * dx + dy * dy <= r2) { r.v = 0; return r; }
        double x = fx, y = fy;
        for (int k = 1; k <= n; k++) {
            x += ux; y += uy;
            dx = ox - x; dy = oy - y;
            const double t = dx * dx + dy * dy - r2;
            if (t <= bj) {
                if (t <= -bj) { r.v = 0; return r; }   // certainly inside: blocked whatever the others say
                unsure = true;
            }
        }
    }
    r.v = unsure ? -1 : 1;
    r.near = (!MASK || b > 64) ? ~0ull : near;
    return r;
}

// the same verdict computed by the whole warp (lanes split the obstacles); uniform result
static __device__ __noinline__ bool edge_free_warp(double fx, double fy, const Steer st, double tx,
                                               double ty, const ObsList L, int lane) {
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

// What the out-of-line helpers need of the parameter block, BY VALUE: taking the address of the kernel parameter `p` (a
// reference argument of a __noinline__ function) makes the compiler keep a copy of all 168 bytes in local memory and read
// p.play_area etc. from THERE instead of the constant bank -- in the hot loop.
struct PlanConsts {
    double res, expand_dis, play[4];
    int has_play;
};
static __device__ __forceinline__ PlanConsts plan_consts(const rrtk_rrtstar_params &p) {
    PlanConsts c;
    c.res = p.path_resolution; c.expand_dis = p.expand_dis; c.has_play = p.has_play_area;
    c.play[0] = p.play_area[0]; c.play[1] = p.play_area[1]; c.play[2] = p.play_area[2]; c.play[3] = p.play_area[3];
    return c;
}
static __device__ __forceinline__ bool inside_play(const PlanConsts &c, double x, double y) {
    if (!c.has_play) return true;
    return !(x < c.play[0] || x > c.play[1] || y < c.play[2] || y > c.play[3]);
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
#ifndef RRTK_DRAW_SAMPLE_INLINE
#define RRTK_DRAW_SAMPLE_INLINE __forceinline__
#endif
static __device__ RRTK_DRAW_SAMPLE_INLINE Sample draw_sample(const rrtk_rrtstar_params &p, int q, int it, int it_key, double gx,
                                              double gy, const double2 *stream, SobolState &sob) {
    Sample s;
    if (p.sampler == RRTK_SAMPLER_STREAM) {
        double2 v = stream[it];
        s.x = v.x; s.y = v.y;
        return s;
    }
    uint64_t k0 = rng_key(p.seed, (uint64_t)(q + p.query_base), (uint64_t)it_key);   // it_key = iteration counter of the whole run
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
static __device__ __noinline__ int best_goal(const PlanConsts p, int n, const double2 *xy,
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
        Steer st = steer(a.x, a.y, gx, gy, CUDART_INF, p.res);
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
