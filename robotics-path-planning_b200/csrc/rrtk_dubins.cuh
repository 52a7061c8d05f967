// rrtk_dubins.cuh -- device functions of the Dubins local planner (rrt_05:935-1278 == dub00), shared by the
// batched steering kernel and the RRT*-Dubins planner kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr double D_TWO_PI = 6.283185307179586;  // 2 * math.pi
constexpr double D_PI = 3.141592653589793;

// Python float `%` / numpy mod
static __device__ __forceinline__ double py_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if ((b < 0) != (r < 0)) r += b; }
    else r = copysign(0.0, b);
    return r;
}
// t % (2 * math.pi).  For |t| < 4 pi -- every call of the planners -- fmod(t, 2 pi) is t or t -+ 2 pi, and that difference
// is exact (Sterbenz), so the result is the one fmod gives: a few compares inline; the fmod routine (long) stays out of
// line for anything larger (the steering code takes ~25 of these per edge)
static __device__ __noinline__ double mod2pi_slow(double t) {
    if (t >= 0.0) {
        if (t < 2.0 * D_TWO_PI) return t - D_TWO_PI;      // [2 pi, 4 pi): exact
    } else if (t > -2.0 * D_TWO_PI) {
        const double r = t + D_TWO_PI;                    // (-4 pi, -2 pi]: exact
        return r != 0.0 ? r + D_TWO_PI : 0.0;
    }
    return py_mod(t, D_TWO_PI);
}
static __device__ __forceinline__ double mod2pi(double t) {
    if (t >= 0.0 && t < D_TWO_PI) return t + 0.0;         // fmod = t; a zero result is +0.0 (-0.0 >= 0.0 lands here too)
    if (t < 0.0 && t > -D_TWO_PI) return t + D_TWO_PI;    // fmod = t (negative, non-zero): r += b, one rounding
    return mod2pi_slow(t);
}
static __device__ __forceinline__ double angle_mod_pi(double x) { return mod2pi(x + D_PI) - D_PI; }
// a / b, correctly rounded as always -- one out-of-line copy of the division sequence (~400 B of SASS per inline site,
// ~15 sites per edge evaluator) instead of one per use
static __device__ __noinline__ double div_rn(double a, double b) { return a / b; }

static __device__ __forceinline__ void sincos_cr(double x, double *s, double *c) {
    crm_sincos(x, s, c);                 // (first-phase evaluation + double-double fallback, crmath.h)
    if (x == 0.0) *s = x;
}

// rot_mat_2d(angle) = [[c, -s], [s, c]] via SciPy's quaternion: c = w*w - z*z, s = 2*(z*w)
static __device__ __forceinline__ void rot2d(double angle, double *c, double *s) {
    double z, w;
    sincos_cr(angle / 2, &z, &w);
    *c = w * w - z * z;
    *s = 2 * (z * w);
}

// word k in _PATH_TYPE_MAP order; false = infeasible
// sin / cos of alpha and beta and cos(alpha - beta): every word function of the reference recomputes them (rrt_05:1125-1198);
// they depend on the edge only, so they are evaluated once per edge
struct DubTrig { double sa, ca, sb, cb, cab; };
static __device__ __forceinline__ DubTrig dubins_trig(double alpha, double beta) {
    DubTrig t;
    sincos_cr(alpha, &t.sa, &t.ca);
    sincos_cr(beta, &t.sb, &t.cb);
    t.cab = crm_cos(alpha - beta);
    return t;
}
// CR = true: the reference's word (correctly rounded atan2 / acos).  CR = false: the same formulas with libdevice's
// atan2 / acos (<= 2 ulp), used only to rule words out (dubins_best_word); *wrap is set when a mod2pi argument lands
// within 1e-9 of a multiple of 2 pi, where the approximate value says nothing about the exact one.  Feasibility (p2 < 0,
// |tmp| > 1) is decided on values both variants compute identically.
template <bool CR>
static __device__ __noinline__ bool dubins_word(int k, double alpha, double beta, double d, const DubTrig &t, double *w,
                                                bool *wrap = nullptr) {
    // ONE body for the six word functions (rrt_05:1125-1198): they differ in signs only, so the expensive calls (atan2,
    // acos, mod2pi) have one call site each instead of 8 / 2 / 16 -- this function was 11 KB of the kernel's 36 KB hot
    // code, and the kernel is bound by instruction fetch.  Every sub-expression is SELECTED between the forms the reference
    // writes (never multiplied by +-1: the sign of a zero decides atan2(+-0, x < 0) = +-pi), in the reference's order.
    const double sa = t.sa, ca = t.ca, sb = t.sb, cb = t.cb, cab = t.cab;
    const double d2 = d * d;
    auto at2 = [](double y, double x) { return CR ? crm_atan2(y, x) : atan2(y, x); };
    auto m2p = [&](double x) {
        const double v = mod2pi(x);
        if (!CR && (v < 1e-9 || v > D_TWO_PI - 1e-9)) *wrap = true;
        return v;
    };
    // the atan2 every word takes: y = +-ca +-cb, x = (d +- sa) +- sb
    //   LSL (cb - ca, d + sa - sb)   RSR (ca - cb, d - sa + sb)   LSR (-ca - cb, d + sa + sb)   RSL (ca + cb, d - sa - sb)
    //   RLR (ca - cb, d - sa + sb)   LRL (ca - cb, d + sa - sb)
    const double y = k == 0 ? cb - ca : (k == 2 ? -ca - cb : (k == 3 ? ca + cb : ca - cb));
    const double x1 = (k == 0 || k == 2 || k == 5) ? d + sa : d - sa;
    const double x = (k == 0 || k == 3 || k == 5) ? x1 - sb : x1 + sb;
    if (k < 4) {   // CSC
        const double s = k < 2 ? (k == 0 ? sa - sb : sb - sa) : sa + sb;
        const double base = k < 2 ? 2 + d2 - (2 * cab) : d2 - 2 + (2 * cab);     // (-2 + d2 of LSR is the same double)
        const double p2 = k == 3 ? base - (2 * d * s) : base + (2 * d * s);
        if (p2 < 0) return false;
        const double d1 = sqrt(p2);
        double tmp = at2(y, x);
        if (k >= 2) tmp = tmp - at2(k == 2 ? -2.0 : 2.0, d1);
        const bool lfirst = (k & 1) == 0;   // LSL, LSR start with a left turn
        w[0] = m2p(lfirst ? -alpha + tmp : alpha - tmp);
        w[1] = d1;
        w[2] = m2p(k == 0 || k == 3 ? beta - tmp : (k == 1 ? -beta + tmp : -mod2pi(beta) + tmp));
        return true;
    }
    // CCC
    const double tmp = (6.0 - d2 + 2.0 * cab + 2.0 * d * (k == 4 ? sa - sb : -sa + sb)) / 8.0;
    if (fabs(tmp) > 1.0) return false;
    w[1] = m2p(2 * D_PI - (CR ? crm_acos(tmp) : acos(tmp)));
    const double T = at2(y, x);
    w[0] = m2p((k == 4 ? alpha - T : -alpha - T) + w[1] / 2.0);
    w[2] = m2p(k == 4 ? alpha - beta - w[0] + w[1] : mod2pi(beta) - alpha - w[0] + mod2pi(w[1]));
    return true;
}

// (Ruling words out with approximate lengths first -- dubins_word<false> -- was tried: the lanes of a warp then disagree on
// which words to evaluate exactly and the warp runs all six anyway, plus the approximations: 63 -> 87 ms at config 4.)
// segment type of word `mode` at position k: 0 = L, 1 = S, 2 = R
static __device__ __forceinline__ int seg_type(int mode, int k) {
    // LSL RSR LSR RSL RLR LRL
    const int t[6][3] = {{0, 1, 0}, {2, 1, 2}, {0, 1, 2}, {2, 1, 0}, {2, 0, 2}, {0, 2, 0}};
    return t[mode][k];
}

// _interpolate (rrt_05:1232-1255); so/co = sin/cos(origin_yaw), sm/cm = sin/cos(-origin_yaw)
static __device__ __forceinline__ void interp(double length, int type, double kappa, double ox, double oy, double oyaw,
                                       double so, double co, double sm, double cm, double *x, double *y,
                                       double *yaw) {
    if (type == 1) {
        const double lk = div_rn(length, kappa);
        *x = ox + lk * co;
        *y = oy + lk * so;
        *yaw = oyaw;
    } else {
        double sl, cl;
        sincos_cr(length, &sl, &cl);
        const double ldx = div_rn(sl, kappa);
        const double q = div_rn(1.0 - cl, kappa);
        const double ldy = type == 0 ? q : -q;          // x / -k == -(x / k) exactly
        const double gdx = cm * ldx + sm * ldy;
        const double gdy = -sm * ldx + cm * ldy;
        *x = ox + gdx;
        *y = oy + gdy;
        *yaw = type == 0 ? oyaw + length : oyaw - length;
    }
}


// ---- collision pre-test of the interior course points (the planners' edge evaluators only) ----
// A blocked / free verdict needs a course point only to within the margin by which it clears the circles.  The arc point
// is first taken with libdevice's sincos (<= 2 ulp; the same formula otherwise), which places it within ~1e-13 of the
// correctly rounded one for coordinates below 1e4; the verdict is accepted when every circle is cleared (or one is
// entered) by more than a band of 1e-9 * (1 + R^2) in squared distance, and the point is re-evaluated with the correctly
// rounded functions otherwise.  Segment end points (the next origin, the end pose) are always exact.
static __device__ __forceinline__ void arc_fast(double length, int type, double rho, double ox, double oy, double sm,
                                                double cm, double *x, double *y) {
    double sl, cl;
    sincos(length, &sl, &cl);
    const double ldx = sl * rho;                      // rho = 1 / kappa rounded: one more ulp, far inside the band
    const double q = (1.0 - cl) * rho;
    const double ldy = type == 0 ? q : -q;
    *x = ox + (cm * ldx + sm * ldy);
    *y = oy + (-sm * ldx + cm * ldy);
}
// 1: inside some circle beyond the band, 0: outside every circle beyond the band, -1: too close to call
static __device__ __forceinline__ int circle_verdict(double wx, double wy, const double4 *obs, int n_obs) {
    int v = 0;
    for (int o = 0; o < n_obs; o++) {
        const double4 ob = obs[o];
        const double dx = ob.x - wx, dy = ob.y - wy;
        const double d2 = dx * dx + dy * dy, band = 1e-9 * (1.0 + ob.w);
        if (d2 <= ob.w - band) return 1;
        if (d2 <= ob.w + band) v = -1;
    }
    return v;
}
static __device__ __forceinline__ bool prefilter_ok(double s_x, double s_y, double g_x, double g_y, double kappa) {
    return fabs(s_x) < 1e4 && fabs(s_y) < 1e4 && fabs(g_x) < 1e4 && fabs(g_y) < 1e4 && kappa > 1e-3;
}

// ---- segment-level cull (the planners' edge evaluators only) ----
// Every sampled point of a course segment lies on the segment's curve: a straight piece from its origin to its end point,
// or an arc of radius 1 / kappa around c = origin + (-+ sin, +- cos)(origin yaw) / kappa.  A circle (centre o, radius R)
// whose distance to that CURVE exceeds R + 1e-7 cannot contain any of the points, so the interior points of a segment
// that no circle comes near are never evaluated (their count is not needed: the planners only ask for the verdict and
// the end pose, which is always computed exactly).  All in the course's local frame (start pose = origin, yaw 0); the
// circle centres are brought there with the same rotation the goal is (rounding ~1e-12 for coordinates < 1e4, far
// inside the 1e-7 margin).  Conservative by construction: any doubt (an angular position within tolerance of the arc's
// ends) counts as "near".
//   type 0 = L, 1 = S, 2 = R;  (ox, oy) origin, (so, co) = sin / cos(origin yaw), (ex, ey) the segment's end point,
//   length = the word length (arc angle in radians for L / R; negative = driven backwards, Reeds-Shepp: the same circle
//   run through the other way);  (px, py) = the circle centre in the local frame
//   aux = segment_aux(...): 1 / kappa for an arc, 1 / |end - origin|^2 for a straight piece (per segment, not per circle)
static __device__ __forceinline__ double segment_aux(int type, double rho, double ox, double oy, double ex, double ey) {
    if (type != 1) return rho;
    const double wx = ex - ox, wy = ey - oy, l2 = wx * wx + wy * wy;
    return l2 > 0.0 ? 1.0 / l2 : 0.0;
}
static __device__ __forceinline__ bool circle_near_segment(int type, double aux, double ox, double oy, double so, double co,
                                                           double ex, double ey, double length, double px, double py,
                                                           double R) {
    const double m = R + 1e-7, m2 = m * m;
    if (type == 1) {
        const double wx = ex - ox, wy = ey - oy;
        const double ax = px - ox, ay = py - oy;
        double t = (ax * wx + ay * wy) * aux;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double qx = ax - t * wx, qy = ay - t * wy;
        return qx * qx + qy * qy <= m2;
    }
    const double rho = aux, sgn = type == 0 ? 1.0 : -1.0;
    const double cx = ox - sgn * rho * so, cy = oy + sgn * rho * co;
    const double vx = px - cx, vy = py - cy;
    const double dc = sqrt(vx * vx + vy * vy);
    if (fabs(dc - rho) > m) return false;                      // clear of the whole circle the arc lies on
    const double ux = ox - cx, uy = oy - cy, wx = ex - cx, wy = ey - cy;
    const double dir = length < 0.0 ? -sgn : sgn;
    const double c1 = dir * (ux * vy - uy * vx), c2 = dir * (vx * wy - vy * wx);   // > 0: o is past the start / before the end
    const double tol = 1e-9 * (rho * dc + 1.0);
    const bool in_span = fabs(length) < D_PI ? (c1 >= -tol && c2 >= -tol) : !(c1 < -tol && c2 < -tol);
    if (in_span) return true;
    const double ax = px - ox, ay = py - oy, bx = px - ex, by = py - ey;           // nearest arc point = one of its ends
    return ax * ax + ay * ay <= m2 || bx * bx + by * by <= m2;
}

struct DubEdge {
    double ex, ey, eyaw;  // last course point (the node pose steer returns, rrt_05:1469-1471); free edges only
    int npts;             // len(px): steer returns None when <= 1 (for a blocked edge only "<= 1 or not" is kept)
    bool free_;           // check_collision over the course points (rrt_05:1625-1638)
};

// ---- pieces of plan_dubins_path for kernels that hold SEVERAL edges per warp (the batched steering kernel) ----
// dubins_front: the goal in the start pose's frame, alpha / beta / d and their sines and cosines (one lane per edge);
// dubins_words_coop: the six words of up to 32 edges shared out over the warp.  (The planners' candidate rounds evaluate
// their words per lane: there the words are a fifth of an edge and the extra code cost more than it saved)
static __device__ __forceinline__ void dubins_front(double s_x, double s_y, double s_yaw, double g_x, double g_y, double g_yaw,
                                                    double kappa, double &c, double &s, double &alpha, double &beta, double &d,
                                                    DubTrig &trig) {
    rot2d(s_yaw, &c, &s);
    const double vx = g_x - s_x, vy = g_y - s_y;
    const double lgx = fma(vy, s, vx * c), lgy = fma(vy, c, vx * -s);
    const double lgyaw = g_yaw - s_yaw;
    d = crm_hypot(lgx, lgy) * kappa;
    const double theta = mod2pi(crm_atan2(lgy, lgx));
    alpha = mod2pi(-theta);
    beta = mod2pi(lgyaw - theta);
    trig = dubins_trig(alpha, beta);
}

// the winning word of an edge: index in _PATH_TYPE_MAP order (0x7fffffff = no word) and its three lengths
struct DubWord { double l0, l1, l2; int bi; };

// Warp-collective: lanes 0..nact-1 hold one edge each (its alpha, beta, d, trig); every such lane gets its edge's first
// minimum of the summed lengths in _PATH_TYPE_MAP order (`best > cost` over k = 0..5, rrt_05:1088-1094).  Two passes so
// that a round runs one kind of word: the four CSC words of 8 edges per round, then the two CCC words of 16 edges per
// round; the first minimum inside an edge's lane group by shuffles, then the edge's own lane takes it.
static __device__ __noinline__ DubWord dubins_words_coop(int nact, int lane, double alpha, double beta, double d, DubTrig trig) {
    DubWord r;
    r.l0 = r.l1 = r.l2 = 0.0;
    r.bi = 0x7fffffff;
    double best = CUDART_INF;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int gl = pass == 0 ? 4 : 2, per = 32 / gl;   // lanes per edge, edges per round
#pragma unroll 1
        for (int c0 = 0; c0 < nact; c0 += per) {
            const int sub = lane & (gl - 1);
            const int cnd = c0 + (pass == 0 ? lane >> 2 : lane >> 1);   // the edge this lane evaluates a word of
            const int k = pass == 0 ? sub : 4 + sub;
            const int src = cnd < 32 ? cnd : 31;
            const double a_ = __shfl_sync(FULL, alpha, src), b_ = __shfl_sync(FULL, beta, src), d_ = __shfl_sync(FULL, d, src);
            DubTrig t_;
            t_.sa = __shfl_sync(FULL, trig.sa, src); t_.ca = __shfl_sync(FULL, trig.ca, src);
            t_.sb = __shfl_sync(FULL, trig.sb, src); t_.cb = __shfl_sync(FULL, trig.cb, src);
            t_.cab = __shfl_sync(FULL, trig.cab, src);
            double w[3] = {0.0, 0.0, 0.0}, cost = CUDART_INF;
            int kk = 0x7fffffff;
            if (cnd < nact && dubins_word<true>(k, a_, b_, d_, t_, w)) { cost = fabs(w[0]) + fabs(w[1]) + fabs(w[2]); kk = k; }
            for (int off = 1; off < gl; off <<= 1) {   // first minimum inside the group
                const double oc = __shfl_xor_sync(FULL, cost, off);
                const int ok = __shfl_xor_sync(FULL, kk, off);
                if (oc < cost || (oc == cost && ok < kk)) { cost = oc; kk = ok; }
            }
            const int g = lane - c0;                   // this lane's own edge is group g of this round when 0 <= g < per
            const bool mine = g >= 0 && g < per && lane < nact;
            const int gsrc = mine ? g * gl : 0;
            const double gc = __shfl_sync(FULL, cost, gsrc);
            const int gk = __shfl_sync(FULL, kk, gsrc);
            const int wsrc = gsrc + (gk != 0x7fffffff ? (pass == 0 ? gk : gk - 4) : 0);
            const double w0 = __shfl_sync(FULL, w[0], wsrc), w1 = __shfl_sync(FULL, w[1], wsrc), w2 = __shfl_sync(FULL, w[2], wsrc);
            if (mine && best > gc) { best = gc; r.bi = gk; r.l0 = w0; r.l1 = w1; r.l2 = w2; }
        }
    }
    return r;
}

// One Dubins edge: plan_dubins_path + sampled collision test.  ONE out-of-line body serves both ways the planners evaluate
// an edge (the kernels are bound by instruction fetch at 7 warps per SM: two 20 KB copies of this code were half of the hot
// footprint):
//   warp = false  the calling lane evaluates the edge alone (choose_parent / rewire: one candidate per lane);
//   warp = true   the whole warp evaluates ONE edge (uniform arguments, uniform result): the six words on six lanes, the
//                 circles of the segment cull and the interior points of a segment spread over the lanes (each lane reaches
//                 its points by the reference's repeated `cur += step`).  Used where the planner has one edge to evaluate
//                 (first steer, re-planned rewire edges, try_goal_path).
// obs rows: x, y, size + robot_radius, (size + robot_radius)**2
// lengths_out (optional): the three course lengths plan_dubins_path returns (word lengths / curvature, rrt_05:1095).
static __device__ __noinline__ DubEdge dubins_edge(double s_x, double s_y, double s_yaw, double g_x, double g_y, double g_yaw,
                                                   double kappa, double step, const double4 *obs, int n_obs, bool warp, int lane,
                                                   double *lengths_out) {
    DubEdge e;
    e.ex = e.ey = e.eyaw = 0.0;
    e.npts = 0;
    e.free_ = false;
    double c, s;
    rot2d(s_yaw, &c, &s);
    const double vx = g_x - s_x, vy = g_y - s_y;
    const double lgx = fma(vy, s, vx * c), lgy = fma(vy, c, vx * -s);
    const double lgyaw = g_yaw - s_yaw;
    const double d = crm_hypot(lgx, lgy) * kappa;
    const double theta = mod2pi(crm_atan2(lgy, lgx));
    const double alpha = mod2pi(-theta), beta = mod2pi(lgyaw - theta);
    const DubTrig trig = dubins_trig(alpha, beta);
    // first minimum of the summed lengths in _PATH_TYPE_MAP order (`best > cost` over k = 0..5, rrt_05:1088-1094)
    double len[3] = {0.0, 0.0, 0.0}, best = CUDART_INF;
    int bi = 0x7fffffff;
    const int k0 = warp ? lane : 0, k1 = warp ? (lane < 6 ? lane + 1 : lane) : 6;
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
        double w[3];
        if (!dubins_word<true>(k, alpha, beta, d, trig, w)) continue;
        const double cost = fabs(w[0]) + fabs(w[1]) + fabs(w[2]);
        if (best > cost) { best = cost; bi = k; len[0] = w[0]; len[1] = w[1]; len[2] = w[2]; }
    }
    if (warp) {
        warp_argmin(best, bi);
        if (bi != 0x7fffffff) {
#pragma unroll
            for (int k = 0; k < 3; k++) len[k] = __shfl_sync(FULL, len[k], bi);
        }
    }
    if (bi == 0x7fffffff) return e;
    if (lengths_out) {
#pragma unroll
        for (int k = 0; k < 3; k++) lengths_out[k] = div_rn(len[k], kappa);
    }
    // rot_mat_2d(-s_yaw): the correctly rounded sin / cos are odd / even bit for bit, so c2 = c and s2 = -s exactly
    const double c2 = c, s2 = -s;
    const bool filt = prefilter_ok(s_x, s_y, g_x, g_y, kappa);
    const double rho = 1.0 / kappa;                   // the approximate tests' radius (the exact points divide by kappa)
    const bool tester = !warp || lane == 0;          // who tests the segment end points
    const int jstep = warp ? 32 : 1;
    bool hit = false;
    double lx = 0.0, ly = 0.0, lyaw = 0.0;
    int np = 1;
    auto test = [&](double px, double py) {
        const double wx = fma(py, s2, px * c2) + s_x;
        const double wy = fma(py, c2, px * -s2) + s_y;
        for (int o = 0; o < n_obs && !hit; o++) {
            const double4 ob = obs[o];
            const double dx = ob.x - wx, dy = ob.y - wy;
            if (dx * dx + dy * dy <= ob.w) hit = true;
        }
    };
    if (tester) test(lx, ly);
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const double length = len[k];
        if (length == 0.0) continue;
        const int type = seg_type(bi, k);
        const double ox = lx, oy = ly, oyaw = lyaw;
        double so = 0.0, co = 1.0;                    // sin / cos(0) of the first segment's origin
        if (oyaw != 0.0) sincos_cr(oyaw, &so, &co);
        const double sm = -so, cm = co;               // sin / cos(-oyaw), exactly (see c2, s2)
        interp(length, type, kappa, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);   // the segment's end point: always exact
        if (tester) test(lx, ly);
        // segment-level cull (circle_near_segment); in warp mode the lanes split the circles
        bool near = !filt;
        const double aux = segment_aux(type, rho, ox, oy, lx, ly);
        for (int o = warp ? lane : 0; o < n_obs && !near; o += jstep) {
            const double4 ob = obs[o];
            const double ux = ob.x - s_x, uy = ob.y - s_y;
            near = circle_near_segment(type, aux, ox, oy, so, co, lx, ly, length, fma(uy, s, ux * c), fma(uy, c, ux * -s), ob.z);
        }
        if (warp) near = __any_sync(FULL, near);
        if (near && !hit) {
            double cur = step;
            for (int t = 0; t < (warp ? lane : 0); t++) cur += step;
#pragma unroll 1
            while (fabs(cur + step) <= fabs(length)) {
                int v = -1;
                if (filt && type != 1) {
                    double x, y;
                    arc_fast(cur, type, rho, ox, oy, sm, cm, &x, &y);
                    v = circle_verdict(fma(y, s2, x * c2) + s_x, fma(y, c2, x * -s2) + s_y, obs, n_obs);
                }
                if (v < 0) {
                    double x, y, yaw;
                    interp(cur, type, kappa, ox, oy, oyaw, so, co, sm, cm, &x, &y, &yaw);
                    test(x, y);
                } else if (v == 1) {
                    hit = true;
                }
                if (hit) break;
#pragma unroll 1
                for (int t = 0; t < jstep; t++) cur += step;
            }
        }
        np++;                                         // (np only has to tell "more than one point")
        if (warp) hit = __any_sync(FULL, hit);
        if (hit) break;
    }
    if (warp) hit = __any_sync(FULL, hit);            // (only the origin can be pending here)
    if (hit) {
        // blocked: the callers only ask whether steer returned a node at all (len(px) > 1, i.e. some segment is non-zero);
        // the rest of the course, its point count and its end pose are never read for a blocked edge
        e.npts = (len[0] != 0.0 || len[1] != 0.0 || len[2] != 0.0) ? 2 : 1;
        return e;
    }
    e.npts = np;
    e.free_ = true;
    e.ex = fma(ly, s2, lx * c2) + s_x;
    e.ey = fma(ly, c2, lx * -s2) + s_y;
    e.eyaw = angle_mod_pi(lyaw + s_yaw);
    return e;
}
static __device__ __forceinline__ DubEdge dubins_edge_lane(double s_x, double s_y, double s_yaw, double g_x, double g_y, double g_yaw,
                                                           double kappa, double step, const double4 *obs, int n_obs) {
    return dubins_edge(s_x, s_y, s_yaw, g_x, g_y, g_yaw, kappa, step, obs, n_obs, false, 0, nullptr);
}
static __device__ __forceinline__ DubEdge dubins_edge_warp(double s_x, double s_y, double s_yaw, double g_x, double g_y, double g_yaw,
                                                           double kappa, double step, const double4 *obs, int n_obs, int lane,
                                                           double *lengths_out = nullptr) {
    return dubins_edge(s_x, s_y, s_yaw, g_x, g_y, g_yaw, kappa, step, obs, n_obs, true, lane, lengths_out);
}

}  // namespace rrtk
