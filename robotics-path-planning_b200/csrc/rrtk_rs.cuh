// rrtk_rs.cuh -- device functions of the Reeds-Shepp local planner (rs00:73-515 == rrt_06:1021-1437), shared by the batched
// steering kernel (rrtk_rs.cu) and the RRT*-Reeds-Shepp planner kernel (rrtk_rrtstar_dubins.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_dubins.cuh"

namespace rrtk {

// rs00:130-139: C fmod semantics, then wrap into [-pi, pi].  Out of line, and without the fmod routine for |x| < 4 pi (every
// call of the planners): there fmod(x, +-2 pi) is x or x -+ 2 pi, an exact difference (Sterbenz).
static __device__ __noinline__ double rs_mod2pi(double x) {
    const double ax = fabs(x);
    double v;
    if (ax < D_TWO_PI) v = x;
    else if (ax < 2.0 * D_TWO_PI) v = x - copysign(D_TWO_PI, x);
    else v = fmod(x, copysign(D_TWO_PI, x));
    if (v == 0.0) v = copysign(0.0, x);
    if (v < -D_PI) v += 2.0 * D_PI;
    else if (v > D_PI) v -= 2.0 * D_PI;
    return v;
}

static __device__ const int8_t RS_T[12][5] = {{0, 1, 0, -1, -1}, {0, 1, 2, -1, -1}, {0, 2, 0, -1, -1}, {0, 2, 0, -1, -1}, {0, 2, 0, -1, -1},
                                       {0, 2, 0, 2, -1}, {0, 2, 0, 2, -1},   {0, 2, 1, 0, -1},  {0, 2, 1, 2, -1},  {0, 1, 2, 0, -1},
                                       {0, 1, 0, 2, -1}, {0, 2, 1, 0, 2}};
static __device__ const int8_t RS_N[12] = {3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5};
// set_path (rs00:141-160) drops a candidate when an already inserted path of the SAME type string is not longer by more than
// step_size: `(existing - new) <= step` for ANY of them, i.e. for the SHORTEST of them (the rounded difference is monotone in
// the existing length).  The 48 candidates (path function f = cand >> 2, symmetry k = cand & 3) have 18 distinct type strings:
// RS_SLOT[cand] numbers them, so the test is one look-up of the shortest inserted length of that string instead of a scan
// of everything inserted so far.
constexpr int RS_SLOTS = 18;
static __device__ const int8_t RS_SLOT[48] = {0, 0, 5, 5, 1, 1, 4, 4, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3, 3, 8, 8, 13, 13, 8, 8, 13, 13,
                                              9, 9, 12, 12, 10, 10, 11, 11, 7, 7, 14, 14, 6, 6, 15, 15, 16, 16, 17, 17};

// the path functions of rs00:166-363 in path_functions order; true + travel distances, or false.
// Every one of them starts from polar(x -/+ sin(phi), y - 1 +/- cos(phi)) (rs00:163-166): (um, thm) is the "minus" pair
// (x - sin, y - 1 + cos), (up, thp) the "plus" pair (x + sin, y - 1 - cos).  The twelve functions of one symmetry share
// these two polars, so an evaluator that walks all 48 candidates computes them 8 times instead of 48 (rs_edge_lane).
static __device__ __forceinline__ bool rs_word_core(int f, double phi, double um, double thm, double up, double thp, double *d) {
    const double pi = D_PI;
    double u, t, v, u1, theta, A;
    if (f == 0) {
        u = um;
        t = thm;
        if (0.0 <= t && t <= pi) {
            v = rs_mod2pi(phi - t);
            if (0.0 <= v && v <= pi) { d[0] = t; d[1] = u; d[2] = v; return true; }
        }
        return false;
    }
    if (f == 1) {
        const double t1 = thp;
        u1 = up * up;
        if (u1 >= 4.0) {
            u = sqrt(u1 - 4.0);
            theta = crm_atan2(2.0, u);
            t = rs_mod2pi(t1 + theta);
            v = rs_mod2pi(t - phi);
            if (t >= 0.0 && v >= 0.0) { d[0] = t; d[1] = u; d[2] = v; return true; }
        }
        return false;
    }
    const bool minus = f == 2 || f == 3 || f == 4 || f == 7 || f == 9;
    u1 = minus ? um : up;
    theta = minus ? thm : thp;
    switch (f) {
        case 2:
            if (u1 <= 4.0) {
                A = crm_acos(0.25 * u1);
                t = rs_mod2pi(A + theta + pi / 2); u = rs_mod2pi(pi - 2 * A); v = rs_mod2pi(phi - t - u);
                d[0] = t; d[1] = -u; d[2] = v; return true;
            }
            return false;
        case 3:
            if (u1 <= 4.0) {
                A = crm_acos(0.25 * u1);
                t = rs_mod2pi(A + theta + pi / 2); u = rs_mod2pi(pi - 2 * A); v = rs_mod2pi(-phi + t + u);
                d[0] = t; d[1] = -u; d[2] = -v; return true;
            }
            return false;
        case 4:
            if (u1 <= 4.0) {
                u = crm_acos(1 - u1 * u1 * 0.125);
                A = crm_asin(2 * crm_sin(u) / u1);
                t = rs_mod2pi(-A + theta + pi / 2); v = rs_mod2pi(t - u - phi);
                d[0] = t; d[1] = u; d[2] = -v; return true;
            }
            return false;
        case 5:
            if (u1 <= 2) {
                A = crm_acos((u1 + 2) * 0.25);
                t = rs_mod2pi(theta + A + pi / 2); u = rs_mod2pi(A); v = rs_mod2pi(phi - t + 2 * u);
                if (t >= 0 && u >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = -u; d[3] = -v; return true; }
            }
            return false;
        case 6: {
            const double u2 = (20 - u1 * u1) / 16;
            if (0 <= u2 && u2 <= 1) {
                u = crm_acos(u2);
                A = crm_asin(2 * crm_sin(u) / u1);
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -u; d[2] = -u; d[3] = v; return true; }
            }
            return false;
        }
        case 7:
            if (u1 >= 2.0) {
                u = sqrt(u1 * u1 - 4) - 2;
                A = crm_atan2(2, sqrt(u1 * u1 - 4));
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi + pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -v; return true; }
            }
            return false;
        case 8:
            if (u1 >= 2.0) {
                t = rs_mod2pi(theta + pi / 2); u = u1 - 2; v = rs_mod2pi(phi - t - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -v; return true; }
            }
            return false;
        case 9:
            if (u1 >= 2.0) {
                u = sqrt(u1 * u1 - 4) - 2;
                A = crm_atan2(sqrt(u1 * u1 - 4), 2);
                t = rs_mod2pi(theta - A + pi / 2); v = rs_mod2pi(t - phi - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = pi / 2; d[3] = -v; return true; }
            }
            return false;
        case 10:
            if (u1 >= 2.0) {
                t = rs_mod2pi(theta); u = u1 - 2; v = rs_mod2pi(phi - t - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = pi / 2; d[3] = -v; return true; }
            }
            return false;
        default:
            if (u1 >= 4.0) {
                u = sqrt(u1 * u1 - 4) - 4;
                A = crm_atan2(2, sqrt(u1 * u1 - 4));
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -pi / 2; d[4] = v; return true; }
            }
            return false;
    }
}


static __device__ __forceinline__ void rs_polars(double x, double y, double phi, bool want_minus, bool want_plus, double *um,
                                                 double *thm, double *up, double *thp) {
    double sp, cp;
    sincos_cr(phi, &sp, &cp);
    if (want_minus) { *um = crm_hypot(x - sp, y - 1.0 + cp); *thm = crm_atan2(y - 1.0 + cp, x - sp); }
    if (want_plus) { *up = crm_hypot(x + sp, y - 1.0 - cp); *thp = crm_atan2(y - 1.0 - cp, x + sp); }
}
static __device__ __noinline__ bool rs_word(int f, double x, double y, double phi, double *d) {
    const bool minus = f == 0 || f == 2 || f == 3 || f == 4 || f == 7 || f == 9;
    double um = 0.0, thm = 0.0, up = 0.0, thp = 0.0;
    rs_polars(x, y, phi, minus, !minus, &um, &thm, &up, &thp);
    return rs_word_core(f, phi, um, thm, up, thp, d);
}

// interpolate (rs00:449-470); sm / cm = sin / cos(-origin_yaw), so / co = sin / cos(origin_yaw)
static __device__ __forceinline__ void rs_interp(double dist, int type, double maxc, double ox, double oy, double oyaw,
                                                 double so, double co, double sm, double cm, double *x, double *y, double *yaw) {
    if (type == 1) {
        const double lk = div_rn(dist, maxc);
        *x = ox + lk * co;
        *y = oy + lk * so;
        *yaw = oyaw;
    } else {
        double sl, cl;
        sincos_cr(dist, &sl, &cl);
        const double ldx = div_rn(sl, maxc);
        const double q = div_rn(1.0 - cl, maxc);
        const double ldy = type == 0 ? q : -q;          // x / -k == -(x / k) exactly
        *yaw = type == 0 ? oyaw + dist : oyaw - dist;
        const double gdx = cm * ldx + sm * ldy;
        const double gdy = -sm * ldx + cm * ldy;
        *x = ox + gdx;
        *y = oy + gdy;
    }
}


// One Reeds-Shepp edge evaluated by ONE lane: reeds_shepp_path_planning + the sampled collision test -- steer
// (rrt_06:1584-1604) and check_collision (:1749-1762).  npts = len(px) (0: steer returns None), lsum = sum(|lengths|).
struct RsEdge {
    double ex, ey, eyaw, lsum;
    int npts;
    bool free_;
};

// The word of an edge: generate_path + set_path in the reference's order, the first shortest inserted word.
//   best = candidate index (family * 4 + symmetry), -1: no course at all; d = its signed lengths (time flip applied);
//   n_ins = len(paths): the words set_path inserted (0 with best = -1)
struct RsPick { double d[5]; int best, n_ins; };
static __device__ __noinline__ RsPick rs_pick_lane(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                   double maxc, double step) {
    RsPick P;
    P.best = -1;
    P.n_ins = 0;
    int n_ins = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) P.d[i] = 0.0;
    const double dx = gx - sx, dy = gy - sy, dth = gyaw - syaw;
    double s0, c0;
    sincos_cr(syaw, &s0, &c0);
    const double x = (c0 * dx + s0 * dy) * maxc, y = (-s0 * dx + c0 * dy) * maxc;
    double min_L[RS_SLOTS], best_L = CUDART_INF;
    int best = -1;
#pragma unroll
    for (int j = 0; j < RS_SLOTS; j++) min_L[j] = CUDART_INF;
    double pol[4][4];   // per symmetry: um, thm, up, thp
#pragma unroll 1
    for (int k = 0; k < 4; k++)
        rs_polars((k & 1) ? -x : x, (k & 2) ? -y : y, (k == 1 || k == 2) ? -dth : dth, true, true, &pol[k][0], &pol[k][1],
                  &pol[k][2], &pol[k][3]);
#pragma unroll 1
    for (int cand = 0; cand < 48; cand++) {
        const int f = cand >> 2, k = cand & 3, n = RS_N[f];
        double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (!rs_word_core(f, (k == 1 || k == 2) ? -dth : dth, pol[k][0], pol[k][1], pol[k][2], pol[k][3], d)) continue;
        double tot = 0.0;
        for (int i = 0; i < n; i++) tot += fabs(d[i]);
        for (int i = 0; i < n; i++) {
            const double a = fabs(d[i]);
            if (0.1 * tot < a && a < step) return P;   // "Step size too large for Reeds-Shepp paths." -> no path at all
        }
        const int slot = RS_SLOT[cand];
        const double shortest = min_L[slot];
        if ((shortest - tot) <= step || tot <= step) continue;   // (inf - tot = inf: nothing of this type string inserted yet)
        if (tot < shortest) min_L[slot] = tot;
        n_ins++;
        const double tl = fabs(div_rn(tot, maxc));
        if (tl < best_L) {
            best_L = tl; best = cand;
            for (int i = 0; i < 5; i++) P.d[i] = (k == 1 || k == 3) ? -d[i] : d[i];   // timeflip
        }
    }
    P.best = best;
    P.n_ins = n_ins;
    return P;
}

// The same for up to 32 edges at once -- warp-collective, lanes 0..nact-1 hold one edge each.  A lane that evaluates its
// 48 candidate words one after the other spends ~90 % of an edge there while (choose_parent / rewire hold ~10 candidates)
// two thirds of the warp idle.  Here lane L works on edge L / 4 of a group of eight and symmetry L % 4 (its polar
// coordinates stay in its registers), the twelve word families run one after the other for the whole warp (uniform code),
// and set_path's bookkeeping runs replicated in the four lanes of an edge, fed by shuffles in candidate order:
// 12 word evaluations per group of eight edges instead of 48 per edge.
static __device__ __noinline__ RsPick rs_pick_coop(int nact, int lane, double sx, double sy, double syaw, double gx, double gy,
                                                   double gyaw, double maxc, double step) {
    RsPick R;
    R.best = -1;
    R.n_ins = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) R.d[i] = 0.0;
#pragma unroll 1
    for (int g0 = 0; g0 < nact; g0 += 8) {
        const int e = g0 + (lane >> 2), k = lane & 3, base = lane & ~3;
        const bool valid = e < nact;
        const int src = valid ? e : 0;
        const double esx = __shfl_sync(FULL, sx, src), esy = __shfl_sync(FULL, sy, src), esyaw = __shfl_sync(FULL, syaw, src);
        const double egx = __shfl_sync(FULL, gx, src), egy = __shfl_sync(FULL, gy, src), egyaw = __shfl_sync(FULL, gyaw, src);
        const double dx = egx - esx, dy = egy - esy, dth = egyaw - esyaw;
        double s0, c0;
        sincos_cr(esyaw, &s0, &c0);
        const double x = (c0 * dx + s0 * dy) * maxc, y = (-s0 * dx + c0 * dy) * maxc;
        const double dthk = (k == 1 || k == 2) ? -dth : dth;
        double um, thm, up, thp;
        rs_polars((k & 1) ? -x : x, (k & 2) ? -y : y, dthk, true, true, &um, &thm, &up, &thp);
        double min_L[RS_SLOTS], best_L = CUDART_INF, bd[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        int best = -1, n_ins = 0;
        bool dead = false;   // "Step size too large": no path at all
#pragma unroll
        for (int j = 0; j < RS_SLOTS; j++) min_L[j] = CUDART_INF;
#pragma unroll 1
        for (int f = 0; f < 12; f++) {
            const int n = RS_N[f];
            double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            const bool ok = valid && rs_word_core(f, dthk, um, thm, up, thp, d);
            double tot = 0.0;
            for (int i = 0; i < n; i++) tot += fabs(d[i]);
            bool bad = false;
            for (int i = 0; i < n; i++) {
                const double a = fabs(d[i]);
                if (0.1 * tot < a && a < step) bad = true;
            }
#pragma unroll 1
            for (int kk = 0; kk < 4; kk++) {   // the four symmetries of this family in candidate order
                const int sl = base + kk;
                const bool okk = __shfl_sync(FULL, (int)ok, sl) != 0, badk = __shfl_sync(FULL, (int)bad, sl) != 0;
                const double totk = __shfl_sync(FULL, tot, sl);
                bool newbest = false;
                if (!dead && okk) {
                    if (badk) dead = true;
                    else {
                        const int cand = 4 * f + kk, slot = RS_SLOT[cand];
                        const double shortest = min_L[slot];
                        if (!((shortest - totk) <= step || totk <= step)) {
                            if (totk < shortest) min_L[slot] = totk;
                            n_ins++;
                            const double tl = fabs(div_rn(totk, maxc));
                            if (tl < best_L) { best_L = tl; best = cand; newbest = true; }
                        }
                    }
                }
                if (__any_sync(FULL, newbest)) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const double di = __shfl_sync(FULL, d[i], sl);
                        if (newbest) bd[i] = (kk == 1 || kk == 3) ? -di : di;   // timeflip
                    }
                }
            }
        }
        if (dead) { best = -1; n_ins = 0; }
        // the edge's own lane takes its group's result
        const int j = lane - g0;
        const bool mine = j >= 0 && j < 8 && lane < nact;
        const int from = mine ? 4 * j : 0;
        const int b = __shfl_sync(FULL, best, from), ni = __shfl_sync(FULL, n_ins, from);
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const double di = __shfl_sync(FULL, bd[i], from);
            if (mine) R.d[i] = di;
        }
        if (mine) { R.best = b; R.n_ins = ni; }
    }
    return R;
}

// The course of the picked word: sampled collision test, end pose, length (steer rrt_06:1584-1604, check_collision
// :1749-1762).  npts = len(px) (0: steer returns None), lsum = sum(|lengths|).
// lengths_only: stop after the lengths (npts = 1 marks "a course exists", lsum = its length) -- what rrt_10's
// calc_new_cost (:1153-1161) needs from reeds_shepp_path_planning.
static __device__ __noinline__ RsEdge rs_course_lane(double sx, double sy, double syaw, double gx, double gy, double maxc,
                                                     double step_size, const double4 *obs, int n_obs, RsPick pick,
                                                     bool lengths_only) {
    RsEdge e;
    e.ex = e.ey = e.eyaw = e.lsum = 0.0;
    e.npts = 0;
    e.free_ = false;
    const int best = pick.best;
    if (best < 0) return e;
    const double step = step_size * maxc;
    double s0, c0;
    sincos_cr(syaw, &s0, &c0);
    double best_d[5];
#pragma unroll
    for (int i = 0; i < 5; i++) best_d[i] = pick.d[i];
    const int f = best >> 2, k = best & 3, n = RS_N[f];
    if (lengths_only) {
        double ls = 0.0;
        for (int i = 0; i < n; i++) ls = ls + fabs(div_rn(best_d[i], maxc));
        e.lsum = ls;
        e.npts = 1;
        return e;
    }
    const double sm0 = -s0, cm0 = c0;   // sin / cos(-syaw): the correctly rounded functions are odd / even bit for bit
    const bool filt = prefilter_ok(sx, sy, gx, gy, maxc);
    const double rho = 1.0 / maxc;                    // the segment cull's radius (approximate test)
    bool hit = false;
    int np = 1;
    double ox = 0.0, oy = 0.0, oyaw = 0.0, lsum = 0.0;
    auto test = [&](double lx, double ly) {
        const double wx = cm0 * lx + sm0 * ly + sx, wy = -sm0 * lx + cm0 * ly + sy;
        for (int o = 0; o < n_obs && !hit; o++) {
            const double4 ob = obs[o];
            const double ex = ob.x - wx, ey = ob.y - wy;
            if (ex * ex + ey * ey <= ob.w) hit = true;
        }
    };
    test(0.0, 0.0);                     // the first point of the course (np.arange starts at 0 in every segment)
#pragma unroll 1
    for (int i = 0; i < n && !hit; i++) {
        const double length = best_d[i];
        lsum = lsum + fabs(div_rn(length, maxc));
        const int t0 = RS_T[f][i];
        const int type = (k >= 2 && t0 != 1) ? 2 - t0 : t0;
        const double dd = length >= 0.0 ? step : -step;
        long long na = length != 0.0 ? (long long)ceil((length - 0.0) / dd) : 0;
        if (na < 0) na = 0;
        double so = 0.0, co = 1.0;
        if (oyaw != 0.0) sincos_cr(oyaw, &so, &co);
        const double sm = -so, cm = co;
        double lx = 0.0, ly = 0.0, lyaw = 0.0;
        rs_interp(length, type, maxc, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);   // the segment's last point: always exact
        test(lx, ly);
        // segment-level cull (circle_near_segment, rrtk_dubins.cuh): the interior points lie on the arc / straight piece
        // between the origin and that last point; they are evaluated only when a circle comes near it
        bool near = !filt;
        const double aux = segment_aux(type, rho, ox, oy, lx, ly);
        for (int o = 0; o < n_obs && !near && !hit; o++) {
            const double4 ob = obs[o];
            const double ux = ob.x - sx, uy = ob.y - sy;
            near = circle_near_segment(type, aux, ox, oy, so, co, lx, ly, length, c0 * ux + s0 * uy, -s0 * ux + c0 * uy, ob.z);
        }
        if (near && !hit) {
#pragma unroll 1
            for (long long j = 1; j < na && !hit; j++) {
                double px, py, pyaw;
                rs_interp(0.0 + (double)j * dd, type, maxc, ox, oy, oyaw, so, co, sm, cm, &px, &py, &pyaw);
                test(px, py);
            }
        }
        ox = lx; oy = ly; oyaw = lyaw;
        np++;                           // (the callers only ask whether a course exists)
    }
    if (hit) {   // blocked: the callers read neither the end pose nor the length of a blocked edge, only that a course exists
        e.npts = 1;
        return e;
    }
    e.npts = np;
    e.free_ = true;
    e.lsum = lsum;
    e.ex = cm0 * ox + sm0 * oy + sx;
    e.ey = -sm0 * ox + cm0 * oy + sy;
    e.eyaw = angle_mod_pi(oyaw + syaw);
    return e;
}
// One Reeds-Shepp edge evaluated by ONE lane: reeds_shepp_path_planning + the sampled collision test
static __device__ __forceinline__ RsEdge rs_edge_lane(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                      double maxc, double step_size, const double4 *obs, int n_obs,
                                                      bool lengths_only = false) {
    const RsPick pick = rs_pick_lane(sx, sy, syaw, gx, gy, gyaw, maxc, step_size * maxc);
    return rs_course_lane(sx, sy, syaw, gx, gy, maxc, step_size, obs, n_obs, pick, lengths_only);
}
// Warp-collective: lanes 0..nact-1 evaluate one edge each, the candidate words shared out (rs_pick_coop)
static __device__ __forceinline__ RsEdge rs_edges_coop(int nact, int lane, double sx, double sy, double syaw, double gx, double gy,
                                                       double gyaw, double maxc, double step_size, const double4 *obs, int n_obs) {
    const RsPick pick = rs_pick_coop(nact, lane, sx, sy, syaw, gx, gy, gyaw, maxc, step_size * maxc);
    RsEdge e;
    e.ex = e.ey = e.eyaw = e.lsum = 0.0;
    e.npts = 0;
    e.free_ = false;
    if (lane < nact) e = rs_course_lane(sx, sy, syaw, gx, gy, maxc, step_size, obs, n_obs, pick, false);
    return e;
}

// Shared-memory scratch of a warp-cooperative Reeds-Shepp evaluation: the 48 candidate words.
struct RsWarp {
    double d[48][5];
    int ok[48];
};

// The same edge evaluated by the WHOLE warp (uniform arguments, uniform result), as rs_steer_kernel does: the 48 words one
// per lane in two rounds, set_path's insertion logic uniformly in the reference's order, the np.arange points of each
// segment spread over the lanes.
static __device__ __noinline__ RsEdge rs_edge_warp(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                   double maxc, double step_size, const double4 *obs, int n_obs, int lane,
                                                   RsWarp &W) {
    RsEdge e;
    e.ex = e.ey = e.eyaw = e.lsum = 0.0;
    e.npts = 0;
    e.free_ = false;
    const double step = step_size * maxc;
    const double dx = gx - sx, dy = gy - sy, dth = gyaw - syaw;
    double s0, c0;
    sincos_cr(syaw, &s0, &c0);
    const double x = (c0 * dx + s0 * dy) * maxc, y = (-s0 * dx + c0 * dy) * maxc;
    __syncwarp();
#pragma unroll 1
    for (int cand = lane; cand < 48; cand += 32) {
        const int f = cand >> 2, k = cand & 3;
        double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        const bool ok = rs_word(f, (k & 1) ? -x : x, (k & 2) ? -y : y, (k == 1 || k == 2) ? -dth : dth, d);
#pragma unroll
        for (int i = 0; i < 5; i++) W.d[cand][i] = d[i];
        W.ok[cand] = ok ? 1 : 0;
    }
    __syncwarp();
    double min_L[RS_SLOTS], best_L = CUDART_INF;
    int best = -1;
#pragma unroll
    for (int j = 0; j < RS_SLOTS; j++) min_L[j] = CUDART_INF;
#pragma unroll 1
    for (int cand = 0; cand < 48; cand++) {
        if (!W.ok[cand]) continue;
        const int f = cand >> 2, n = RS_N[f];
        double tot = 0.0;
        for (int i = 0; i < n; i++) tot += fabs(W.d[cand][i]);
        for (int i = 0; i < n; i++) {
            const double a = fabs(W.d[cand][i]);
            if (0.1 * tot < a && a < step) return e;   // "Step size too large for Reeds-Shepp paths." -> no path at all
        }
        const int slot = RS_SLOT[cand];
        const double shortest = min_L[slot];
        if ((shortest - tot) <= step || tot <= step) continue;
        if (tot < shortest) min_L[slot] = tot;
        const double tl = fabs(div_rn(tot, maxc));
        if (tl < best_L) { best_L = tl; best = cand; }
    }
    if (best < 0) return e;
    const int f = best >> 2, k = best & 3, n = RS_N[f];
    const double sm0 = -s0, cm0 = c0;   // sin / cos(-syaw), exactly (see rs_edge_lane)
    const bool filt = prefilter_ok(sx, sy, gx, gy, maxc);
    const double rho = 1.0 / maxc;                    // the segment cull's radius (approximate test)
    bool hit = false;
    int np = 1;
    double ox = 0.0, oy = 0.0, oyaw = 0.0, lsum = 0.0;
    auto test = [&](double lx, double ly) {
        const double wx = cm0 * lx + sm0 * ly + sx, wy = -sm0 * lx + cm0 * ly + sy;
        for (int o = 0; o < n_obs && !hit; o++) {
            const double4 ob = obs[o];
            const double ex = ob.x - wx, ey = ob.y - wy;
            if (ex * ex + ey * ey <= ob.w) hit = true;
        }
    };
    if (lane == 0) test(0.0, 0.0);
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        double length = W.d[best][i];
        if (k == 1 || k == 3) length = -length;   // timeflip
        lsum = lsum + fabs(div_rn(length, maxc));
        const int t0 = RS_T[f][i];
        const int type = (k >= 2 && t0 != 1) ? 2 - t0 : t0;
        const double dd = length >= 0.0 ? step : -step;
        long long na = length != 0.0 ? (long long)ceil((length - 0.0) / dd) : 0;
        if (na < 0) na = 0;
        double so = 0.0, co = 1.0;
        if (oyaw != 0.0) sincos_cr(oyaw, &so, &co);
        const double sm = -so, cm = co;
        double lx, ly, lyaw;   // the segment's last point is the next origin (uniform)
        rs_interp(length, type, maxc, ox, oy, oyaw, so, co, sm, cm, &lx, &ly, &lyaw);
        if (lane == 0) test(lx, ly);
        bool near = !filt;     // segment-level cull, the lanes split the circles
        const double aux = segment_aux(type, rho, ox, oy, lx, ly);
        for (int o = lane; o < n_obs && !near; o += 32) {
            const double4 ob = obs[o];
            const double ux = ob.x - sx, uy = ob.y - sy;
            near = circle_near_segment(type, aux, ox, oy, so, co, lx, ly, length, c0 * ux + s0 * uy, -s0 * ux + c0 * uy, ob.z);
        }
        if (__any_sync(FULL, near)) {
#pragma unroll 1
            for (long long j = 1 + lane; j < na; j += 32) {
                double px, py, pyaw;
                rs_interp(0.0 + (double)j * dd, type, maxc, ox, oy, oyaw, so, co, sm, cm, &px, &py, &pyaw);
                test(px, py);
            }
        }
        if (__any_sync(FULL, hit)) {   // blocked: see rs_edge_lane
            e.npts = 1;
            return e;
        }
        ox = lx; oy = ly; oyaw = lyaw;
        np++;
    }
    e.npts = np;
    e.free_ = true;
    e.lsum = lsum;
    e.ex = cm0 * ox + sm0 * oy + sx;
    e.ey = -sm0 * ox + cm0 * oy + sy;
    e.eyaw = angle_mod_pi(oyaw + syaw);
    return e;
}

}  // namespace rrtk
