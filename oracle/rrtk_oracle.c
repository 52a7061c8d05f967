/* TEST INFRASTRUCTURE ONLY -- plain-C (FP64) restatement of the reference hot path.
 *
 * CPU oracle for the RRT-family inner loop of gouldberg/robotics-path-planning.  It performs
 * the same IEEE-754 double operations, in the same order, as the reference's Python code:
 *   - `math.hypot`  -> orc_hypot(): CPython's correctly-rounded vector_norm algorithm
 *                      (glibc hypot() differs from CPython's in ~0.6 % of inputs),
 *   - `x ** 2`      -> pow(x, 2.0)  (CPython float_pow calls libm pow; it is NOT always x*x),
 *   - `math.cos/sin/atan2/log/sqrt/floor` -> the same glibc functions CPython calls.
 * It is pinned bit-for-bit against tests/golden/ (outputs of the unmodified reference, made
 * by oracle/make_golden.py); see tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product never does.
 *
 * Citations are alias:line into /root/reference/src_path_planning.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math, no -march flags).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* "cr" math mode: the correctly-rounded leaf functions the GPU path uses (same header, so the
 * GPU and the oracle's cr mode are bit-identical by construction; the header itself is checked
 * against mpmath in tests/test_crmath.py).  The "libm" mode uses none of it. */
#include "../robotics-path-planning_b200/csrc/crmath.h"
#define ORC_MATH_LIBM 0
#define ORC_MATH_CR 1

/* ------------------------------------------------------------------------------------ */
/* math.hypot of CPython >= 3.10 (Modules/mathmodule.c vector_norm, n = 2)               */
/* ------------------------------------------------------------------------------------ */
typedef struct { double hi, lo; } dl_t;

static inline dl_t dl_fast_sum(double a, double b) {
    double x = a + b;
    double y = (a - x) + b;
    dl_t r = {x, y};
    return r;
}

static inline dl_t dl_mul(double x, double y) {
    double z = x * y;
    double zz = fma(x, y, -z); /* exact low word; software fma when the ISA has none */
    dl_t r = {z, zz};
    return r;
}

ORC_EXPORT double orc_hypot(double a, double b) {
    double v[2];
    v[0] = fabs(a);
    v[1] = fabs(b);
    double max = v[0] > v[1] ? v[0] : v[1];
    if (isinf(v[0]) || isinf(v[1])) return INFINITY;
    if (isnan(a) || isnan(b)) return NAN;
    if (max == 0.0) return max;
    int max_e;
    frexp(max, &max_e);
    if (max_e < -1023) return DBL_MIN * orc_hypot(v[0] / DBL_MIN, v[1] / DBL_MIN);
    double scale = ldexp(1.0, -max_e);
    double csum = 1.0, frac1 = 0.0, frac2 = 0.0, x, h;
    dl_t pr, sm;
    for (int i = 0; i < 2; i++) {
        x = v[i] * scale;
        pr = dl_mul(x, x);
        sm = dl_fast_sum(csum, pr.hi);
        csum = sm.hi;
        frac1 += pr.lo;
        frac2 += sm.lo;
    }
    h = sqrt(csum - 1.0 + (frac1 + frac2));
    pr = dl_mul(-h, h);
    sm = dl_fast_sum(csum, pr.hi);
    csum = sm.hi;
    frac1 += pr.lo;
    frac2 += sm.lo;
    x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);
    return h / scale;
}

static inline double sq_libm(double x) { return pow(x, 2.0); } /* Python `x ** 2` */
/* per-node squared distances: libm mode = Python `**` (libm pow); cr mode = exact IEEE square */
static inline double sq(int mode, double x) { return mode == ORC_MATH_LIBM ? pow(x, 2.0) : x * x; }

ORC_EXPORT double orc_cr_sin(double x) { return crm_sin(x); }
ORC_EXPORT double orc_cr_cos(double x) { return crm_cos(x); }
ORC_EXPORT double orc_cr_atan2(double y, double x) { return crm_atan2(y, x); }
ORC_EXPORT double orc_cr_hypot(double a, double b) { return crm_hypot(a, b); }
ORC_EXPORT double orc_cr_acos(double x) { return crm_acos(x); }
ORC_EXPORT double orc_cr_asin(double x) { return crm_asin(x); }
ORC_EXPORT double orc_cr_tan(double x) { return crm_tan(x); }
ORC_EXPORT double orc_cr_atan2_sincos(double y, double x, double *s, double *c) {
    return crm_atan2_sincos(y, x, s, c);
}

/* ------------------------------------------------------------------------------------ */
/* Sobol (rrt_04:230-503): closed form, Gray-code order                                  */
/* ------------------------------------------------------------------------------------ */
#define SOBOL_BITS 30
#define SOBOL_DIM_MAX 40
static const int k_poly[SOBOL_DIM_MAX] = {
    1, 3, 7, 11, 13, 19, 25, 37, 59, 47, 61, 55, 41, 67, 97, 91, 109, 103, 115, 131,
    193, 137, 145, 143, 241, 157, 185, 167, 229, 171, 213, 191, 253, 203, 211, 239,
    247, 285, 369, 299};
/* initial m_j per column: (first dimension index, count, values) -- the data at rrt_04:320-356 */
static const int k_c1[] = {1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3};
static const int k_c2[] = {7, 5, 1, 3, 3, 7, 5, 5, 7, 7, 1, 3, 3, 7, 5, 1, 1, 5, 3, 3, 1, 7, 5, 1, 3, 3, 7, 5, 1, 1, 5, 7, 7, 5, 1, 3, 3};
static const int k_c3[] = {1, 7, 9, 13, 11, 1, 3, 7, 9, 5, 13, 13, 11, 3, 15, 5, 3, 15, 7, 9, 13, 9, 1, 11, 7, 5, 15, 1, 15, 11, 5, 3, 1, 7, 9};
static const int k_c4[] = {9, 3, 27, 15, 29, 21, 23, 19, 11, 25, 7, 13, 17, 1, 25, 29, 3, 31, 11, 5, 23, 27, 19, 21, 5, 1, 17, 13, 7, 15, 9, 31, 9};
static const int k_c5[] = {37, 33, 7, 5, 11, 39, 63, 27, 17, 15, 23, 29, 3, 21, 13, 31, 25, 9, 49, 33, 19, 29, 11, 19, 27, 15, 25};
static const int k_c6[] = {13, 33, 115, 41, 79, 17, 29, 119, 75, 73, 105, 7, 59, 65, 21, 3, 113, 61, 89, 45, 107};
static const int k_c7[] = {7, 23, 39};

ORC_EXPORT int orc_sobol_table(int dim_num, uint32_t *v /* [dim_num][30] */) {
    if (dim_num < 1 || dim_num > SOBOL_DIM_MAX) return -1;
    static const int *cols[8] = {0, k_c1, k_c2, k_c3, k_c4, k_c5, k_c6, k_c7};
    static const int first[8] = {0, 2, 3, 5, 7, 13, 19, 37};
    uint32_t m[SOBOL_DIM_MAX][SOBOL_BITS];
    memset(m, 0, sizeof m);
    for (int d = 0; d < dim_num; d++) m[d][0] = 1;
    for (int c = 1; c < 8; c++)
        for (int d = first[c]; d < dim_num; d++) m[d][c] = (uint32_t)cols[c][d - first[c]];
    for (int j = 0; j < SOBOL_BITS; j++) m[0][j] = 1;
    for (int d = 1; d < dim_num; d++) {
        int poly = k_poly[d], deg = 0;
        for (int p = poly >> 1; p; p >>= 1) deg++;
        for (int j = deg; j < SOBOL_BITS; j++) {
            uint32_t newv = m[d][j - deg];
            for (int k = 0; k < deg; k++)
                if ((poly >> (deg - 1 - k)) & 1) newv ^= (2u << k) * m[d][j - k - 1];
            m[d][j] = newv;
        }
    }
    for (int d = 0; d < dim_num; d++)
        for (int j = 0; j < SOBOL_BITS; j++) v[d * SOBOL_BITS + j] = m[d][j] << (SOBOL_BITS - 1 - j);
    return 0;
}

/* out[i*dim + d] = d-th coordinate of point (first_index + i)  (i4_sobol, rrt_04:494-503) */
ORC_EXPORT int orc_sobol_fill(int dim_num, int64_t first_index, int64_t count, double *out) {
    uint32_t v[SOBOL_DIM_MAX * SOBOL_BITS];
    if (orc_sobol_table(dim_num, v)) return -1;
    const double recipd = 1.0 / 1073741824.0;
    for (int64_t i = 0; i < count; i++) {
        int64_t n = first_index + i;
        if (n < 0) n = 0;
        uint64_t g = (uint64_t)n ^ ((uint64_t)n >> 1);
        for (int d = 0; d < dim_num; d++) {
            uint32_t q = 0;
            uint64_t gg = g;
            for (int j = 0; gg && j < SOBOL_BITS; j++, gg >>= 1)
                if (gg & 1) q ^= v[d * SOBOL_BITS + j];
            out[i * dim_num + d] = (double)q * recipd;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Scenario + tree                                                                       */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double sx, sy, gx, gy;
    double expand_dis, res, robot_radius, connect_circle_dist;
    double play[4]; /* xmin xmax ymin ymax */
    int32_t has_play;
    int32_t max_iter;
    int32_t search_until_max_iter;
    int32_t n_obs;
    int32_t math_mode; /* ORC_MATH_LIBM (the reference on this platform) or ORC_MATH_CR */
    int32_t pad_;
} orc_params_t;

#define MAXPTS 4096

typedef struct {
    const orc_params_t *p;
    const double *obs; /* n_obs * 3 */
    double *x, *y, *cost;
    int32_t *parent;
    int n;
    uint8_t *verdicts;
    int64_t n_verdicts, verdict_cap;
    /* scratch for one edge */
    double px[MAXPTS], py[MAXPTS];
    int npts;
    /* tie log: decisions whose relative margin is below ORC_TIE_EPS (north_star: "ties within 1e-6
     * are logged").  rows of (iteration, kind, margin); kinds: 0 snap `d <= res`, 1 floor(extend/res),
     * 2 near `d2 <= r2`, 3 collision `min d2 <= R2`, 4 rewire `cost > edge cost`, 5 nearest argmin */
    double *ties;
    int64_t n_ties, tie_cap;
    int cur_it;
    /* work counters: point-circle tests and node-distance evaluations the reference performs */
    int64_t work_pairs, work_scan;
} orc_tree_t;

#define ORC_TIE_EPS 1e-6
static void log_tie(orc_tree_t *t, int kind, double margin) {
    if (fabs(margin) >= ORC_TIE_EPS) return;
    if (t->ties && t->n_ties < t->tie_cap) {
        t->ties[3 * t->n_ties] = (double)t->cur_it;
        t->ties[3 * t->n_ties + 1] = (double)kind;
        t->ties[3 * t->n_ties + 2] = margin;
    }
    t->n_ties++;
}

/* steer (rrt_04:1086-1115): fills t->px/py, returns the end point */
static void steer(orc_tree_t *t, double fx, double fy, double tx, double ty, double extend,
                  double *ex, double *ey) {
    const double res = t->p->res;
    double x = fx, y = fy;
    double dx = tx - x, dy = ty - y;
    double d = orc_hypot(dx, dy);
    double ct, st;
    if (t->p->math_mode == ORC_MATH_LIBM) {
        double theta = atan2(dy, dx);
        ct = cos(theta);
        st = sin(theta);
    } else {
        (void)crm_atan2_sincos(dy, dx, &st, &ct);
    }
    int np = 0;
    t->px[np] = x; t->py[np] = y; np++;
    if (extend > d) extend = d;
    double q = floor(extend / res);
    long n_expand = (long)q;
    if (t->ties) { double fr = extend / res - q; log_tie(t, 1, fr < 0.5 ? fr : fr - 1.0); }
    for (long k = 0; k < n_expand; k++) {
        x += res * ct;
        y += res * st;
        if (np < MAXPTS) { t->px[np] = x; t->py[np] = y; np++; }
    }
    double d2 = orc_hypot(tx - x, ty - y);
    if (t->ties) log_tie(t, 0, (d2 - res) / res);
    if (d2 <= res) {
        if (np < MAXPTS) { t->px[np] = tx; t->py[np] = ty; np++; }
        x = tx; y = ty;
    }
    t->npts = np;
    *ex = x; *ey = y;
}

/* check_collision (rrt_04:1216-1230) on the edge in scratch; logs the verdict. 1 = safe */
static int collision_free(orc_tree_t *t) {
    const orc_params_t *p = t->p;
    int ok = 1;
    for (int o = 0; o < p->n_obs && ok; o++) {
        double ox = t->obs[3 * o], oy = t->obs[3 * o + 1], size = t->obs[3 * o + 2];
        double mn = INFINITY;
        t->work_pairs += t->npts;
        for (int k = 0; k < t->npts; k++) {
            double dx = ox - t->px[k], dy = oy - t->py[k];
            double dd = dx * dx + dy * dy;
            if (dd < mn) mn = dd;
        }
        if (t->ties) log_tie(t, 3, (mn - sq_libm(size + p->robot_radius)) / sq_libm(size + p->robot_radius));
        if (mn <= sq_libm(size + p->robot_radius)) ok = 0; /* per-obstacle constant: host libm */
    }
    if (t->verdicts && t->n_verdicts < t->verdict_cap) t->verdicts[t->n_verdicts] = (uint8_t)ok;
    t->n_verdicts++;
    return ok;
}

static int inside_play(const orc_params_t *p, double x, double y) { /* rrt_04:1204-1214 */
    if (!p->has_play) return 1;
    if (x < p->play[0] || x > p->play[1] || y < p->play[2] || y > p->play[3]) return 0;
    return 1;
}

static int edge_ok(orc_tree_t *t, double ex, double ey) { /* rrt_04:1267-1269 order */
    return collision_free(t) && inside_play(t->p, ex, ey);
}

static void propagate(orc_tree_t *t, int p) { /* rrt_04:1379-1384 */
    for (int c = 0; c < t->n; c++)
        if (t->parent[c] == p) {
            t->cost[c] = t->cost[p] + orc_hypot(t->x[c] - t->x[p], t->y[c] - t->y[p]);
            propagate(t, c);
        }
}

/* search_best_goal_node (rrt_04:1284-1312); -1 = None.  `dist`, `first` are scratch [n]. */
static int best_goal(orc_tree_t *t, double *dist, int *cand) {
    const orc_params_t *p = t->p;
    int n = t->n, nc = 0;
    for (int i = 0; i < n; i++) dist[i] = orc_hypot(t->x[i] - p->gx, t->y[i] - p->gy);
    for (int i = 0; i < n; i++)
        if (dist[i] <= p->expand_dis) {
            int f = 0;
            while (dist[f] != dist[i]) f++; /* list.index(): first equal value */
            cand[nc++] = f;
        }
    int best = -1;
    double best_cost = INFINITY;
    for (int k = 0; k < nc; k++) {
        int gi = cand[k];
        double ex, ey;
        steer(t, t->x[gi], t->y[gi], p->gx, p->gy, INFINITY, &ex, &ey);
        if (edge_ok(t, ex, ey)) {
            double c = t->cost[gi] + orc_hypot(t->x[gi] - p->gx, t->y[gi] - p->gy);
            if (best < 0 || c < best_cost) { best = gi; best_cost = c; } /* first minimum */
        }
    }
    return best;
}

/* RRT* planning loop (rrt_04:1036-1084).  Arrays x,y,cost,parent have capacity max_iter+1.
 * trace (optional): int32[max_iter][8] = nearest,status,n_near,parent,cp_ok,rw_ok,rw_applied,n_after
 * Returns 0.  *goal_index = best goal node (-1 = no path). */
ORC_EXPORT int orc_rrtstar_run(const orc_params_t *p, const double *obs, const double *stream,
                               double *x, double *y, double *cost, int32_t *parent,
                               int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
                               int32_t *trace, uint8_t *verdicts, int64_t verdict_cap,
                               int64_t *n_verdicts, double *ties, int64_t tie_cap, int64_t *n_ties,
                               int64_t *work /* [2] or NULL */) {
    orc_tree_t *t = (orc_tree_t *)calloc(1, sizeof(orc_tree_t));
    int cap = p->max_iter + 1;
    double *dist = (double *)malloc(sizeof(double) * cap);
    int *near = (int *)malloc(sizeof(int) * cap);
    double *costs = (double *)malloc(sizeof(double) * cap);
    t->p = p; t->obs = obs; t->x = x; t->y = y; t->cost = cost; t->parent = parent;
    t->verdicts = verdicts; t->verdict_cap = verdict_cap;
    t->ties = ties; t->tie_cap = tie_cap;
    x[0] = p->sx; y[0] = p->sy; cost[0] = 0.0; parent[0] = -1; t->n = 1;
    int gi = -1, it = 0, done = 0;
    for (it = 0; it < p->max_iter; it++) {
        double rx = stream[2 * it], ry = stream[2 * it + 1];
        int n = t->n;
        t->cur_it = it;
        /* nearest (rrt_04:1196-1202): first minimum */
        int ni = 0;
        double dmin = INFINITY, dsecond = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(p->math_mode, x[i] - rx) + sq(p->math_mode, y[i] - ry);
            if (d < dmin) { dsecond = dmin; dmin = d; ni = i; }
            else if (d < dsecond) dsecond = d;
        }
        if (t->ties && dsecond < INFINITY && dsecond != dmin && dmin > 0.0) log_tie(t, 5, (dsecond - dmin) / dmin);
        t->work_scan += n;
        double nx, ny;
        steer(t, x[ni], y[ni], rx, ry, p->expand_dis, &nx, &ny);
        double ncost = cost[ni] + orc_hypot(nx - x[ni], ny - y[ni]);
        int status = 0, n_near = 0, par = -1, cp_ok = 0, rw_ok = 0, rw_applied = 0;
        if (inside_play(p, nx, ny)) {
            status = 1;
            if (collision_free(t)) {
                /* find_near_nodes (rrt_04:1314-1338) */
                double nnode = (double)(n + 1);
                double r = p->connect_circle_dist * sqrt(log(nnode) / nnode);
                if (p->expand_dis < r) r = p->expand_dis;
                double r2 = sq_libm(r); /* per-iteration constant: host libm in the product too */
                t->work_scan += n;
                for (int i = 0; i < n; i++) dist[i] = sq(p->math_mode, x[i] - nx) + sq(p->math_mode, y[i] - ny);
                for (int i = 0; i < n; i++) {
                    if (t->ties && r2 > 0.0) log_tie(t, 2, (dist[i] - r2) / r2);
                    if (dist[i] <= r2) {
                        int f = 0;
                        while (dist[f] != dist[i]) f++;
                        near[n_near++] = f;
                    }
                }
                /* choose_parent (rrt_04:1242-1282) */
                int best = -1;
                double min_cost = INFINITY;
                for (int k = 0; k < n_near; k++) {
                    int i = near[k];
                    double ex, ey;
                    steer(t, x[i], y[i], nx, ny, INFINITY, &ex, &ey);
                    if (edge_ok(t, ex, ey)) {
                        costs[k] = cost[i] + orc_hypot(nx - x[i], ny - y[i]);
                        cp_ok++;
                    } else {
                        costs[k] = INFINITY;
                    }
                    if (costs[k] < min_cost) { min_cost = costs[k]; best = i; }
                }
                if (best >= 0) {
                    double cx, cy;
                    steer(t, x[best], y[best], nx, ny, INFINITY, &cx, &cy);
                    double ccost = min_cost;
                    /* rewire (rrt_04:1340-1373), before the append (rrt_04:1064-1065) */
                    for (int k = 0; k < n_near; k++) {
                        int i = near[k];
                        double ex, ey;
                        steer(t, cx, cy, x[i], y[i], INFINITY, &ex, &ey);
                        double ecost = ccost + orc_hypot(x[i] - cx, y[i] - cy);
                        int ok = edge_ok(t, ex, ey);
                        rw_ok += ok;
                        if (t->ties && ok && cost[i] != ecost) log_tie(t, 4, (cost[i] - ecost) / ecost);
                        if (ok && cost[i] > ecost) {
                            x[i] = ex; y[i] = ey; cost[i] = ecost; parent[i] = n;
                            rw_applied++;
                            propagate(t, i);
                        }
                    }
                    x[n] = cx; y[n] = cy; cost[n] = ccost; parent[n] = best;
                    status = 3; par = best;
                } else {
                    x[n] = nx; y[n] = ny; cost[n] = ncost; parent[n] = ni;
                    status = 2; par = ni;
                }
                t->n = n + 1;
            }
        }
        if (trace) {
            int32_t *tr = trace + 8 * it;
            tr[0] = ni; tr[1] = status; tr[2] = n_near; tr[3] = par; tr[4] = cp_ok;
            tr[5] = rw_ok; tr[6] = rw_applied; tr[7] = t->n;
        }
        if (!p->search_until_max_iter) {
            gi = best_goal(t, dist, near);
            if (gi >= 0) { it++; done = 1; break; }
        }
    }
    if (!done) gi = best_goal(t, dist, near);
    *n_nodes = t->n; *iters_done = it; *goal_index = gi;
    if (n_verdicts) *n_verdicts = t->n_verdicts;
    if (n_ties) *n_ties = t->n_ties;
    if (work) { work[0] = t->work_pairs; work[1] = t->work_scan; }
    free(dist); free(near); free(costs); free(t);
    return 0;
}

/* basic RRT loop (rrt_01:71-101); *goal_index = index of the node that reached the goal */
ORC_EXPORT int orc_rrt_run(const orc_params_t *p, const double *obs, const double *stream,
                           double *x, double *y, int32_t *parent, int32_t *n_nodes,
                           int32_t *iters_done, int32_t *goal_index) {
    orc_tree_t *t = (orc_tree_t *)calloc(1, sizeof(orc_tree_t));
    t->p = p; t->obs = obs; t->x = x; t->y = y; t->parent = parent;
    x[0] = p->sx; y[0] = p->sy; parent[0] = -1; t->n = 1;
    int gi = -1, it;
    for (it = 0; it < p->max_iter; it++) {
        double rx = stream[2 * it], ry = stream[2 * it + 1];
        int n = t->n, ni = 0;
        double dmin = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(p->math_mode, x[i] - rx) + sq(p->math_mode, y[i] - ry);
            if (d < dmin) { dmin = d; ni = i; }
        }
        double nx, ny;
        steer(t, x[ni], y[ni], rx, ry, p->expand_dis, &nx, &ny);
        if (inside_play(p, nx, ny) && collision_free(t)) {
            x[n] = nx; y[n] = ny; parent[n] = ni; t->n = n + 1;
        }
        double lx = x[t->n - 1], ly = y[t->n - 1];
        if (orc_hypot(lx - p->gx, ly - p->gy) <= p->expand_dis) {
            double ex, ey;
            steer(t, lx, ly, p->gx, p->gy, p->expand_dis, &ex, &ey);
            if (collision_free(t)) { gi = t->n - 1; it++; break; }
        }
    }
    *n_nodes = t->n; *iters_done = it; *goal_index = gi;
    free(t);
    return 0;
}

/* single-edge helper for unit tests: steer + collision + play area */
ORC_EXPORT int orc_steer_collide(const orc_params_t *p, const double *obs, double fx, double fy,
                                 double tx, double ty, double extend, double *ex, double *ey,
                                 int32_t *npts, int32_t *free_flag, int32_t *inside_flag) {
    orc_tree_t *t = (orc_tree_t *)calloc(1, sizeof(orc_tree_t));
    t->p = p; t->obs = obs;
    steer(t, fx, fy, tx, ty, extend, ex, ey);
    *npts = t->npts;
    *free_flag = collision_free(t);
    *inside_flag = inside_play(p, *ex, *ey);
    free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Arm C-space occupancy grid (arm02:46-110, :257-262)                                   */
/* ------------------------------------------------------------------------------------ */
/* numpy's dot of two 2-vectors on this platform (OpenBLAS ddot, FMA kernel) evaluates
 * fma(a1, b1, a0 * b0) -- verified against numpy for 20 000 random pairs */
static inline double dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

/* detect_collision (arm02:46-76): 1 = the link [a, b] touches the circle */
static int arm_detect_collision(double ax, double ay, double bx, double by, double cx, double cy, double r) {
    double l0 = bx - ax, l1 = by - ay;
    double mag = sqrt(dot2(l0, l1, l0, l1)); /* np.linalg.norm */
    double v0 = cx - ax, v1 = cy - ay;
    double proj = dot2(v0, v1, l0 / mag, l1 / mag);
    double p0, p1;
    if (proj <= 0) { p0 = ax; p1 = ay; }
    else if (proj >= mag) { p0 = bx; p1 = by; }
    else { p0 = ax + l0 * proj / mag; p1 = ay + l1 * proj / mag; }
    double w0 = p0 - cx, w1 = p1 - cy;
    double dist = sqrt(dot2(w0, w1, w0, w1));
    if (dist > r) return 0;
    return 1;
}

/* get_occupancy_grid (arm02:79-110) for rows [row0, row0 + n_rows): grid[(i - row0) * M + j].
 * theta[M] is the reference's theta_list (host-evaluated); joint k uses theta1 (k = 1) or
 * theta1 + theta2 (k >= 2), the 2-angle prefix-sum rule of arm02:259-260. */
ORC_EXPORT int orc_arm_grid(int32_t M, const double *theta, int32_t row0, int32_t n_rows, int32_t n_links,
                            const double *link, const double *obs /* [O][3] */, int32_t O,
                            int32_t math_mode, uint8_t *grid) {
    if (n_links < 1 || n_links > 16) return -1;
    for (int i = row0; i < row0 + n_rows; i++)
        for (int j = 0; j < M; j++) {
            double px[17], py[17];
            px[0] = 0.0; py[0] = 0.0;
            double a1 = theta[i], a2 = theta[i] + theta[j];
            for (int k = 1; k <= n_links; k++) {
                double ang = k == 1 ? a1 : a2;
                double c, s;
                if (math_mode == ORC_MATH_LIBM) { c = cos(ang); s = sin(ang); }
                else { c = crm_cos(ang); s = crm_sin(ang); }
                px[k] = px[k - 1] + link[k - 1] * c;
                py[k] = py[k - 1] + link[k - 1] * s;
            }
            int hit = 0;
            for (int k = 0; k < n_links && !hit; k++)
                for (int o = 0; o < O && !hit; o++)
                    hit = arm_detect_collision(px[k], py[k], px[k + 1], py[k + 1], obs[3 * o], obs[3 * o + 1],
                                               obs[3 * o + 2]);
            grid[(size_t)(i - row0) * M + j] = (uint8_t)hit;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Informed RRT* (rrt_07:1027-1285)                                                      */
/* ------------------------------------------------------------------------------------ */
/* check_segment_collision + distance_squared_point_to_segment (rrt_07:1249-1269); obs rows are
 * (x, y, size, size**2) with the square taken by the host as Python does.  1 = free */
static int segment_free(double x1, double y1, double x2, double y2, const double *obs4, int n_obs) {
    for (int o = 0; o < n_obs; o++) {
        double ox = obs4[4 * o], oy = obs4[4 * o + 1], r2 = obs4[4 * o + 3];
        double dd;
        if (x1 == x2 && y1 == y2) {
            dd = dot2(ox - x1, oy - y1, ox - x1, oy - y1);
        } else {
            double wx = x2 - x1, wy = y2 - y1;
            double l2 = dot2(wx, wy, wx, wy);
            double t = dot2(ox - x1, oy - y1, wx, wy) / l2;
            t = t < 1.0 ? t : 1.0;   /* min(1, t); a NaN t stays NaN like Python's min/max chain... */
            t = t > 0.0 ? t : 0.0;   /* ...then max(0, nan) = 0 in Python because `nan > 0` is False */
            double px = x1 + t * wx, py = y1 + t * wy;
            dd = dot2(ox - px, oy - py, ox - px, oy - py);
        }
        if (dd <= r2) return 0;
    }
    return 1;
}

typedef struct {
    double sx, sy, gx, gy;
    double expand_dis;
    double rot[4];      /* 2 x 2 block of C (rrt_07:1063-1068), host-evaluated */
    int32_t max_iter, n_obs, math_mode, path_cap;
} orc_informed_params_t;

static void dir_trig(int mode, double dy, double dx, double *theta, double *s, double *c) {
    if (mode == ORC_MATH_LIBM) { *theta = atan2(dy, dx); *c = cos(*theta); *s = sin(*theta); }
    else *theta = crm_atan2_sincos(dy, dx, s, c);
}

/* check_collision (rrt_07:1271-1276) given cos/sin of theta */
static int inf_check(double x, double y, double c, double s, double d, const double *obs4, int n_obs) {
    return segment_free(x, y, x + c * d, y + s * d, obs4, n_obs);
}

ORC_EXPORT int orc_informed_run(const orc_informed_params_t *p, const double *obs4, const double *free_s,
                                const double *ball, double *x, double *y, double *cost, int32_t *parent,
                                int32_t *n_nodes, double *path /* [path_cap][2] */, int32_t *path_len,
                                double *c_best_out) {
    const int mode = p->math_mode;
    int cap = p->max_iter + 1, n = 1;
    double *dl = (double *)malloc(sizeof(double) * cap);
    int *near = (int *)malloc(sizeof(int) * cap);
    int hb = 10;
    while ((1 << hb) < 4 * cap && hb < 28) hb++;
    uint32_t *hgen = (uint32_t *)calloc((size_t)1 << hb, sizeof(uint32_t));
    uint64_t *hkey = (uint64_t *)malloc(sizeof(uint64_t) << hb);
    int *hidx = (int *)malloc(sizeof(int) << hb);
    x[0] = p->sx; y[0] = p->sy; cost[0] = 0.0; parent[0] = -1;
    double c_best = INFINITY;
    const double c_min = orc_hypot(p->sx - p->gx, p->sy - p->gy);
    const double xc = (p->sx + p->gx) / 2.0, yc = (p->sy + p->gy) / 2.0, ed = p->expand_dis;
    *path_len = 0;
    for (int it = 0; it < p->max_iter; it++) {
        double rx, ry;
        if (c_best < INFINITY) { /* informed_sample rrt_07:1145-1159 */
            double r0 = c_best / 2.0;
            double r1 = sqrt(sq(mode, c_best) - sq(mode, c_min)) / 2.0;
            double a = ball[2 * it], b = ball[2 * it + 1];
            if (b < a) { double t = a; a = b; b = t; }
            double ang = 2 * 3.141592653589793 * a / b, bx, by; /* Python: 2 * math.pi * a / b */
            if (mode == ORC_MATH_LIBM) { bx = b * cos(ang); by = b * sin(ang); }
            else { bx = b * crm_cos(ang); by = b * crm_sin(ang); }
            double m00 = p->rot[0] * r0, m01 = p->rot[1] * r1, m10 = p->rot[2] * r0, m11 = p->rot[3] * r1;
            rx = fma(m00, bx, m01 * by) + xc; /* numpy (3x3)@(3x1) on this platform, see DESIGN.md */
            ry = fma(m10, bx, m11 * by) + yc;
        } else {
            rx = free_s[2 * it]; ry = free_s[2 * it + 1];
        }
        int ni = 0;
        double dmin = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(mode, x[i] - rx) + sq(mode, y[i] - ry);
            if (d < dmin) { dmin = d; ni = i; }
        }
        double theta, st, ct;
        dir_trig(mode, ry - y[ni], rx - x[ni], &theta, &st, &ct);
        double nx = x[ni] + ed * ct, ny = y[ni] + ed * st; /* get_new_node rrt_07:1216-1224 */
        double ncost = cost[ni] + ed;
        int npar = ni;
        double d = orc_hypot(x[ni] - nx, y[ni] - ny);
        if (!inf_check(x[ni], y[ni], ct, st, d, obs4, p->n_obs)) continue;
        /* find_near_nodes rrt_07:1137-1143 */
        double r = 50.0 * sqrt(log((double)n) / (double)n);
        double r2 = sq_libm(r); /* per-size constant: host libm in the product too */
        int n_near = 0;
        for (int i = 0; i < n; i++) dl[i] = sq(mode, x[i] - nx) + sq(mode, y[i] - ny);
        /* near_inds = [d_list.index(v) for v in d_list if v <= r ** 2]: `.index` returns the FIRST position
         * holding an equal value; that position is itself a hit, so it is found through a hash of the hits'
         * bit patterns (generation-stamped, no clearing) instead of a scan from 0 -- same result, O(1). */
        for (int i = 0; i < n; i++)
            if (dl[i] <= r2) {
                uint64_t key;
                memcpy(&key, &dl[i], 8);
                uint64_t h = key * 0x9E3779B97F4A7C15ull;
                uint32_t slot = (uint32_t)(h >> (64 - hb));
                int f = i;
                for (;;) {
                    if (hgen[slot] != (uint32_t)(it + 1)) { hgen[slot] = (uint32_t)(it + 1); hkey[slot] = key; hidx[slot] = i; break; }
                    if (hkey[slot] == key) { f = hidx[slot]; break; }
                    slot = (slot + 1) & ((1u << hb) - 1u);
                }
                near[n_near++] = f;
            }
        /* choose_parent rrt_07:1110-1135 */
        double mc = INFINITY;
        int best = -1;
        for (int k = 0; k < n_near; k++) {
            int i = near[k];
            double dx = nx - x[i], dy = ny - y[i], th, s2, c2;
            double dd = orc_hypot(dx, dy);
            dir_trig(mode, dy, dx, &th, &s2, &c2);
            double c = inf_check(x[i], y[i], c2, s2, dd, obs4, p->n_obs) ? cost[i] + dd : INFINITY;
            if (c < mc) { mc = c; best = i; }
        }
        if (best >= 0) { ncost = mc; npar = best; }
        x[n] = nx; y[n] = ny; cost[n] = ncost; parent[n] = npar;
        int newi = n;
        n++;
        /* rewire rrt_07:1232-1246 */
        for (int k = 0; k < n_near; k++) {
            int i = near[k];
            double dd = orc_hypot(x[i] - nx, y[i] - ny);
            double sc = ncost + dd;
            if (cost[i] > sc) {
                double th, s2, c2;
                dir_trig(mode, ny - y[i], nx - x[i], &th, &s2, &c2);
                if (inf_check(x[i], y[i], c2, s2, dd, obs4, p->n_obs)) { parent[i] = newi; cost[i] = sc; }
            }
        }
        /* goal bookkeeping rrt_07:1094-1103 */
        if (orc_hypot(nx - p->gx, ny - p->gy) < ed && segment_free(nx, ny, p->gx, p->gy, obs4, p->n_obs)) {
            double plen = 0.0, qx = p->gx, qy = p->gy;
            int k = newi, len = 1;
            while (parent[k] >= 0) {
                plen += orc_hypot(x[k] - qx, y[k] - qy);
                qx = x[k]; qy = y[k];
                k = parent[k]; len++;
            }
            plen += orc_hypot(p->sx - qx, p->sy - qy);
            len++;
            if (plen < c_best) {
                c_best = plen;
                int w = 0;
                if (w < p->path_cap) { path[0] = p->gx; path[1] = p->gy; }
                w++;
                for (k = newi; parent[k] >= 0; k = parent[k], w++)
                    if (w < p->path_cap) { path[2 * w] = x[k]; path[2 * w + 1] = y[k]; }
                if (w < p->path_cap) { path[2 * w] = p->sx; path[2 * w + 1] = p->sy; }
                w++;
                *path_len = w;
            }
        }
    }
    *n_nodes = n;
    *c_best_out = c_best;
    free(dl); free(near); free(hgen); free(hkey); free(hidx);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* path smoothing (rrt_04:1390-1479): random shortcutting of the final course             */
/* ------------------------------------------------------------------------------------ */
static double sm_path_length(const double *path, int len) { /* rrt_04:1390-1398 */
    double le = 0.0;
    for (int i = 0; i + 1 < len; i++) le += orc_hypot(path[2 * (i + 1)] - path[2 * i], path[2 * (i + 1) + 1] - path[2 * i + 1]);
    return le;
}
/* get_target_point (rrt_04:1401-1420), ti = i - 1 quirk included; returns 0, or 1 for the reference's ZeroDivisionError */
static int sm_target_point(const double *path, int len, double target, double *x, double *y, int *ti_out) {
    double le = 0.0, last = 0.0;
    int ti = 0;
    for (int i = 0; i + 1 < len; i++) {
        double d = orc_hypot(path[2 * (i + 1)] - path[2 * i], path[2 * (i + 1) + 1] - path[2 * i + 1]);
        le += d;
        if (le >= target) { ti = i - 1; last = d; break; }
    }
    if (last == 0.0) return 1;
    double ratio = (le - target) / last;
    int a = ti < 0 ? len + ti : ti; /* Python negative index */
    *x = path[2 * a] + (path[2 * (ti + 1)] - path[2 * a]) * ratio;
    *y = path[2 * a + 1] + (path[2 * (ti + 1) + 1] - path[2 * a + 1]) * ratio;
    *ti_out = ti;
    return 0;
}
/* path [cap][2] in/out, *len in/out; draws [max_iter][2] in [0,1); obs [n_obs][3] = x, y, size.
 * returns 0, 1 = the reference would raise ZeroDivisionError (path as of that iteration), 2 = cap too small */
ORC_EXPORT int orc_path_smoothing(double *path, int32_t *len_io, int32_t cap, const double *draws, int32_t max_iter,
                                  const double *obs, int32_t n_obs, int32_t *iters_done) {
    int len = *len_io;
    double *tmp = (double *)malloc(sizeof(double) * 2 * (size_t)cap);
    double le = sm_path_length(path, len);
    int rc = 0, it = 0;
    for (it = 0; it < max_iter; it++) {
        double p0 = 0 + (le - 0) * draws[2 * it], p1 = 0 + (le - 0) * draws[2 * it + 1];
        if (p1 < p0) { double t = p0; p0 = p1; p1 = t; }
        double fx, fy, sx, sy;
        int t1, t2;
        if (sm_target_point(path, len, p0, &fx, &fy, &t1) || sm_target_point(path, len, p1, &sx, &sy, &t2)) { rc = 1; break; }
        if (t1 <= 0 || t2 <= 0) continue;
        if (t2 + 1 > len) continue;
        if (t2 == t1) continue;
        /* line_collision_check rrt_04:1423-1444 */
        double a = sy - fy, b = -(sx - fx), c = sy * (sx - fx) - sx * (sy - fy);
        int ok = 1;
        for (int o = 0; o < n_obs && ok; o++) {
            double h = orc_hypot(a, b);
            if (h == 0.0) { rc = 1; ok = -1; break; }
            double d = fabs(a * obs[3 * o] + b * obs[3 * o + 1] + c) / h;
            if (d <= obs[3 * o + 2]) ok = 0;
        }
        if (ok < 0) break;
        if (!ok) continue;
        int nl = (t1 + 1) + 2 + (len - t2 - 1);
        if (nl > cap) { rc = 2; break; }
        memcpy(tmp, path, sizeof(double) * 2 * (size_t)(t1 + 1));
        tmp[2 * (t1 + 1)] = fx; tmp[2 * (t1 + 1) + 1] = fy;
        tmp[2 * (t1 + 2)] = sx; tmp[2 * (t1 + 2) + 1] = sy;
        memcpy(tmp + 2 * (t1 + 3), path + 2 * (t2 + 1), sizeof(double) * 2 * (size_t)(len - t2 - 1));
        memcpy(path, tmp, sizeof(double) * 2 * (size_t)nl);
        len = nl;
        le = sm_path_length(path, len);
    }
    *len_io = len;
    *iters_done = it;
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* astar_torus (arm02:113-233): greedy best-first search on the joint-space torus         */
/* ------------------------------------------------------------------------------------ */
static int64_t at_orig(int M, int gi, int gj, int i, int j) { (void)M; return (int64_t)abs(j - gj) + (int64_t)abs(i - gi); }
/* calc_heuristic_map (arm02:221-233): the value cell (i, j) ends with, in-place update order included */
static int64_t at_new(int M, int gi, int gj, int i, int j, int depth) {
    int64_t v = at_orig(M, gi, gj, i, j), t;
    int64_t row0 = (i > 0 && depth == 0) ? at_new(M, gi, gj, 0, j, 1) : at_orig(M, gi, gj, 0, j);
    int64_t col0 = (j > 0) ? at_new(M, gi, gj, i, 0, 2) : at_orig(M, gi, gj, i, 0);
    if (depth == 2 && i > 0) row0 = at_new(M, gi, gj, 0, 0, 1); /* column-0 cell of a later row reads updated (0, 0) */
    if (depth == 1) row0 = at_orig(M, gi, gj, 0, j);            /* a row-0 cell reads its own (original) value */
    t = i + 1 + at_orig(M, gi, gj, M - 1, j); if (t < v) v = t;
    t = M - i + row0; if (t < v) v = t;
    t = j + 1 + at_orig(M, gi, gj, i, M - 1); if (t < v) v = t;
    t = M - j + col0; if (t < v) v = t;
    return v;
}
ORC_EXPORT void orc_astar_heuristic(int32_t M, int32_t gi, int32_t gj, int64_t *out) {
    /* plain sequential restatement (the recursive closed form above is what the GPU uses; tests compare the two) */
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) out[(size_t)i * M + j] = at_orig(M, gi, gj, i, j);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            int64_t v = out[(size_t)i * M + j], t;
            t = i + 1 + out[(size_t)(M - 1) * M + j]; if (t < v) v = t;
            t = M - i + out[j]; if (t < v) v = t;
            t = j + 1 + out[(size_t)i * M + M - 1]; if (t < v) v = t;
            t = M - j + out[(size_t)i * M]; if (t < v) v = t;
            out[(size_t)i * M + j] = v;
        }
}
ORC_EXPORT int64_t orc_astar_heuristic_closed(int32_t M, int32_t gi, int32_t gj, int32_t i, int32_t j) {
    return at_new(M, gi, gj, i, j, 0);
}
static void at_push(uint64_t *heap, int *n, uint64_t key) {
    int k = (*n)++;
    while (k > 0) { int p = (k - 1) / 2; if (heap[p] <= key) break; heap[k] = heap[p]; k = p; }
    heap[k] = key;
}
static void at_pop(uint64_t *heap, int *n) {
    uint64_t key = heap[--(*n)];
    int k = 0;
    for (;;) {
        int c = 2 * k + 1;
        if (c >= *n) break;
        if (c + 1 < *n && heap[c + 1] < heap[c]) c++;
        if (heap[c] >= key) break;
        heap[k] = heap[c]; k = c;
    }
    if (*n > 0) heap[k] = key;
}
/* grid [M][M] uint8 in/out (0 free, 1 occupied -> marks 2..6); route [cap][2] start -> goal; returns route length
 * (0 = no route), or -1 if the route does not fit */
ORC_EXPORT int orc_astar_torus(uint8_t *grid, int32_t M, int32_t si, int32_t sj, int32_t gi, int32_t gj, int32_t *route,
                               int32_t cap) {
    size_t cells = (size_t)M * M;
    int64_t *h = (int64_t *)malloc(sizeof(int64_t) * cells);
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * cells);
    uint64_t *heap = (uint64_t *)malloc(sizeof(uint64_t) * (cells + 8));
    int nheap = 0, found = 0, s = si * M + sj, g = gi * M + gj;
    orc_astar_heuristic(M, gi, gj, h);
    for (size_t k = 0; k < cells; k++) parent[k] = -1;
    grid[s] = 4; grid[g] = 5;
    at_push(heap, &nheap, ((uint64_t)h[s] << 32) | (uint32_t)s);
    for (;;) {
        grid[s] = 4; grid[g] = 5;
        if (nheap == 0) break;
        int cur = (int)(heap[0] & 0xffffffffu);
        if (cur == g) { found = 1; break; }
        at_pop(heap, &nheap);
        grid[cur] = 2;
        int i = cur / M, j = cur % M;
        int nb[4] = {(i - 1 >= 0 ? i - 1 : M - 1) * M + j, (i + 1 < M ? i + 1 : 0) * M + j,
                     i * M + (j - 1 >= 0 ? j - 1 : M - 1), i * M + (j + 1 < M ? j + 1 : 0)};
        for (int k = 0; k < 4; k++)
            if (grid[nb[k]] == 0 || grid[nb[k]] == 5) {
                at_push(heap, &nheap, ((uint64_t)h[nb[k]] << 32) | (uint32_t)nb[k]);
                parent[nb[k]] = cur;
                grid[nb[k]] = 3;
            }
    }
    int len = 0;
    if (found) {
        for (int k = g; k >= 0; k = parent[k]) len++;
        if (len > cap) len = -1;
        else {
            int w = len - 1;
            for (int k = g; k >= 0; k = parent[k], w--) { route[2 * w] = k / M; route[2 * w + 1] = k % M; if (w >= 1) grid[k] = 6; }
        }
    }
    free(h); free(parent); free(heap);
    return len;
}

/* ------------------------------------------------------------------------------------ */
/* Dubins local planner (rrt_05:935-1278 == dub00)                                       */
/* ------------------------------------------------------------------------------------ */
#define ORC_TWO_PI 6.283185307179586 /* 2 * math.pi */
#define ORC_PI 3.141592653589793

typedef struct { int mode; } orc_math_t;
static inline double m_sin(int m, double x) { return m == ORC_MATH_LIBM ? sin(x) : crm_sin(x); }
static inline double m_cos(int m, double x) { return m == ORC_MATH_LIBM ? cos(x) : crm_cos(x); }
static inline double m_atan2(int m, double y, double x) { return m == ORC_MATH_LIBM ? atan2(y, x) : crm_atan2(y, x); }
static inline double m_acos(int m, double x) { return m == ORC_MATH_LIBM ? acos(x) : crm_acos(x); }

/* Python float `%` (== numpy mod): fmod, then move into the sign of the divisor */
static inline double py_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if ((b < 0) != (r < 0)) r += b; }
    else r = copysign(0.0, b);
    return r;
}
static inline double mod2pi(double t) { return py_mod(t, ORC_TWO_PI); }              /* rrt_05:1112 */
static inline double angle_mod_pi(double x) { return py_mod(x + ORC_PI, ORC_TWO_PI) - ORC_PI; } /* :1005 */

/* rot_mat_2d(angle) = [[c, -s], [s, c]] through SciPy's quaternion (see oracle/pyport.py) */
static void rot2d(int m, double angle, double *c, double *s) {
    double z = m_sin(m, angle / 2), w = m_cos(m, angle / 2);
    *c = w * w - z * z;
    *s = 2 * (z * w);
}

/* word k of _PATH_TYPE_MAP order LSL,RSR,LSR,RSL,RLR,LRL (rrt_05:1125-1198); returns 0 if infeasible */
static int dubins_word(int m, int k, double alpha, double beta, double d, double *w) {
    double sa = m_sin(m, alpha), sb = m_sin(m, beta), ca = m_cos(m, alpha), cb = m_cos(m, beta);
    double cab = m_cos(m, alpha - beta);
    double d2 = sq(m, d), p2, tmp, d1;
    switch (k) {
        case 0:
            p2 = 2 + d2 - (2 * cab) + (2 * d * (sa - sb));
            if (p2 < 0) return 0;
            tmp = m_atan2(m, (cb - ca), d + sa - sb);
            w[0] = mod2pi(-alpha + tmp); w[1] = sqrt(p2); w[2] = mod2pi(beta - tmp);
            return 1;
        case 1:
            p2 = 2 + d2 - (2 * cab) + (2 * d * (sb - sa));
            if (p2 < 0) return 0;
            tmp = m_atan2(m, (ca - cb), d - sa + sb);
            w[0] = mod2pi(alpha - tmp); w[1] = sqrt(p2); w[2] = mod2pi(-beta + tmp);
            return 1;
        case 2:
            p2 = -2 + d2 + (2 * cab) + (2 * d * (sa + sb));
            if (p2 < 0) return 0;
            d1 = sqrt(p2);
            tmp = m_atan2(m, (-ca - cb), (d + sa + sb)) - m_atan2(m, -2.0, d1);
            w[0] = mod2pi(-alpha + tmp); w[1] = d1; w[2] = mod2pi(-mod2pi(beta) + tmp);
            return 1;
        case 3:
            p2 = d2 - 2 + (2 * cab) - (2 * d * (sa + sb));
            if (p2 < 0) return 0;
            d1 = sqrt(p2);
            tmp = m_atan2(m, (ca + cb), (d - sa - sb)) - m_atan2(m, 2.0, d1);
            w[0] = mod2pi(alpha - tmp); w[1] = d1; w[2] = mod2pi(beta - tmp);
            return 1;
        case 4:
            tmp = (6.0 - d2 + 2.0 * cab + 2.0 * d * (sa - sb)) / 8.0;
            if (fabs(tmp) > 1.0) return 0;
            w[1] = mod2pi(2 * ORC_PI - m_acos(m, tmp));
            w[0] = mod2pi(alpha - m_atan2(m, ca - cb, d - sa + sb) + w[1] / 2.0);
            w[2] = mod2pi(alpha - beta - w[0] + w[1]);
            return 1;
        default:
            tmp = (6.0 - d2 + 2.0 * cab + 2.0 * d * (-sa + sb)) / 8.0;
            if (fabs(tmp) > 1.0) return 0;
            w[1] = mod2pi(2 * ORC_PI - m_acos(m, tmp));
            w[0] = mod2pi(-alpha - m_atan2(m, ca - cb, d + sa - sb) + w[1] / 2.0);
            w[2] = mod2pi(mod2pi(beta) - alpha - w[0] + mod2pi(w[1]));
            return 1;
    }
}

static const char k_dubins_modes[6][4] = {"LSL", "RSR", "LSR", "RSL", "RLR", "LRL"};

/* _interpolate (rrt_05:1232-1255) */
static void dubins_interp(int m, double length, char mode, double kappa, double ox, double oy, double oyaw,
                          double *x, double *y, double *yaw) {
    if (mode == 'S') {
        *x = ox + length / kappa * m_cos(m, oyaw);
        *y = oy + length / kappa * m_sin(m, oyaw);
        *yaw = oyaw;
    } else {
        double ldx = m_sin(m, length) / kappa, ldy;
        if (mode == 'L') ldy = (1.0 - m_cos(m, length)) / kappa;
        else ldy = (1.0 - m_cos(m, length)) / -kappa;
        double gdx = m_cos(m, -oyaw) * ldx + m_sin(m, -oyaw) * ldy;
        double gdy = -m_sin(m, -oyaw) * ldx + m_cos(m, -oyaw) * ldy;
        *x = ox + gdx;
        *y = oy + gdy;
        *yaw = mode == 'L' ? oyaw + length : oyaw - length;
    }
}

/* plan_dubins_path (rrt_05:1021-1109).  out_xyyaw [max_pts][3] may be NULL.  Returns the number of points. */
ORC_EXPORT int orc_dubins_plan(double s_x, double s_y, double s_yaw, double g_x, double g_y, double g_yaw,
                               double kappa, double step, int math_mode, int32_t *mode_out, double *lengths,
                               double *out_xyyaw, int32_t max_pts) {
    const int m = math_mode;
    double c, s;
    rot2d(m, s_yaw, &c, &s);
    double vx = g_x - s_x, vy = g_y - s_y;
    double lgx = fma(vy, s, vx * c), lgy = fma(vy, c, vx * -s); /* numpy (2,)@(2,2) on this platform */
    double lgyaw = g_yaw - s_yaw;
    double d = orc_hypot(lgx, lgy) * kappa;
    double theta = mod2pi(m_atan2(m, lgy, lgx));
    double alpha = mod2pi(-theta), beta = mod2pi(lgyaw - theta);
    double best_cost = INFINITY, bw[3] = {0, 0, 0};
    int bi = -1;
    for (int k = 0; k < 6; k++) {
        double w[3];
        if (!dubins_word(m, k, alpha, beta, d, w)) continue;
        double cost = fabs(w[0]) + fabs(w[1]) + fabs(w[2]);
        if (best_cost > cost) { best_cost = cost; bi = k; bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2]; }
    }
    if (bi < 0) return -1;
    *mode_out = bi;
    for (int k = 0; k < 3; k++) lengths[k] = bw[k] / kappa;
    double c2, s2;
    rot2d(m, -s_yaw, &c2, &s2);
    /* _generate_local_course (rrt_05:1258-1278), each local point converted to the world frame */
    int np = 0;
    double lx = 0.0, ly = 0.0, lyaw = 0.0;
#define EMIT(px, py, pyaw)                                                                   \
    do {                                                                                     \
        if (out_xyyaw && np < max_pts) {                                                     \
            out_xyyaw[3 * np] = fma((py), s2, (px) * c2) + s_x;                              \
            out_xyyaw[3 * np + 1] = fma((py), c2, (px) * -s2) + s_y;                          \
            out_xyyaw[3 * np + 2] = angle_mod_pi((pyaw) + s_yaw);                            \
        }                                                                                    \
        np++;                                                                                \
    } while (0)
    EMIT(lx, ly, lyaw);
    for (int k = 0; k < 3; k++) {
        double length = bw[k];
        char md = k_dubins_modes[bi][k];
        if (length == 0.0) continue;
        double ox = lx, oy = ly, oyaw = lyaw, cur = step;
        while (fabs(cur + step) <= fabs(length)) {
            dubins_interp(m, cur, md, kappa, ox, oy, oyaw, &lx, &ly, &lyaw);
            EMIT(lx, ly, lyaw);
            cur += step;
        }
        dubins_interp(m, length, md, kappa, ox, oy, oyaw, &lx, &ly, &lyaw);
        EMIT(lx, ly, lyaw);
    }
#undef EMIT
    return np;
}

/* ------------------------------------------------------------------------------------ */
/* RRT*-Dubins planning loop (rrt_05:1416-1779)                                          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double sx, sy, syaw, gx, gy, gyaw;
    double expand_dis, robot_radius, connect_circle_dist, kappa, goal_yaw_th, goal_xy_th;
    int32_t max_iter, n_obs, search_until_max_iter, math_mode;
} orc_dubins_params_t;

#define DUB_MAXPTS 8192
typedef struct { double end[3]; int npts; int free_; } dub_edge_t;

/* steer (rrt_05:1458-1479) + check_collision (:1625-1638): npts <= 1 means steer returned None */
static dub_edge_t dubins_edge(const orc_dubins_params_t *p, const double *obs3, const double *f, const double *t,
                              double *buf) {
    dub_edge_t e;
    int32_t mode;
    double lengths[3];
    int n = orc_dubins_plan(f[0], f[1], f[2], t[0], t[1], t[2], p->kappa, 0.1, p->math_mode, &mode, lengths, buf,
                            DUB_MAXPTS);
    e.npts = n;
    e.free_ = 0;
    e.end[0] = e.end[1] = e.end[2] = 0.0;
    if (n <= 1) return e;
    if (n > DUB_MAXPTS) n = DUB_MAXPTS;
    e.end[0] = buf[3 * (n - 1)]; e.end[1] = buf[3 * (n - 1) + 1]; e.end[2] = buf[3 * (n - 1) + 2];
    int ok = 1;
    for (int o = 0; o < p->n_obs && ok; o++) {
        double ox = obs3[3 * o], oy = obs3[3 * o + 1], size = obs3[3 * o + 2], mn = INFINITY;
        for (int k = 0; k < n; k++) {
            double dx = ox - buf[3 * k], dy = oy - buf[3 * k + 1], dd = dx * dx + dy * dy;
            if (dd < mn) mn = dd;
        }
        if (mn <= sq_libm(size + p->robot_radius)) ok = 0;
    }
    e.free_ = ok;
    return e;
}

static void dub_propagate(int n, double *x, double *y, double *cost, const int32_t *parent, int p) {
    for (int c = 0; c < n; c++)
        if (parent[c] == p) {
            cost[c] = cost[p] + orc_hypot(x[c] - x[p], y[c] - y[p]);
            dub_propagate(n, x, y, cost, parent, c);
        }
}

static int dub_best_goal(const orc_dubins_params_t *p, int n, const double *x, const double *y, const double *yaw,
                         const double *cost) {
    int best = -1;
    double bc = INFINITY;
    for (int i = 0; i < n; i++)
        if (orc_hypot(x[i] - p->gx, y[i] - p->gy) <= p->goal_xy_th && fabs(yaw[i] - p->gyaw) <= p->goal_yaw_th)
            if (best < 0 || cost[i] < bc) { best = i; bc = cost[i]; }
    return best;
}

/* edge_from / edge_to [cap][3]: the pose pair whose Dubins course is the node's path_x/path_y/path_yaw
 * (plan_dubins_path(edge_from, edge_to) regenerates it).  *goal_index: -1 none (index 0 also counts as none) */
ORC_EXPORT int orc_rrtstar_dubins_run(const orc_dubins_params_t *p, const double *obs3, const double *stream3,
                                      double *x, double *y, double *yaw, double *cost, int32_t *parent,
                                      double *edge_from, double *edge_to, int32_t *n_nodes, int32_t *iters_done,
                                      int32_t *goal_index) {
    const int mode = p->math_mode;
    int cap = p->max_iter + 1, n = 1, it, gi = -1, done = 0;
    double *buf = (double *)malloc(sizeof(double) * 3 * DUB_MAXPTS);
    double *dl = (double *)malloc(sizeof(double) * cap);
    int *near = (int *)malloc(sizeof(int) * cap);
    x[0] = p->sx; y[0] = p->sy; yaw[0] = p->syaw; cost[0] = 0.0; parent[0] = -1;
    for (it = 0; it < p->max_iter; it++) {
        const double *rnd = stream3 + 3 * it;
        int ni = 0;
        double dmin = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(mode, x[i] - rnd[0]) + sq(mode, y[i] - rnd[1]);
            if (d < dmin) { dmin = d; ni = i; }
        }
        double from[3] = {x[ni], y[ni], yaw[ni]};
        dub_edge_t e0 = dubins_edge(p, obs3, from, rnd, buf);
        int truthy = e0.npts > 1;
        if (e0.npts > 1 && e0.free_) {
            double nw[3] = {e0.end[0], e0.end[1], e0.end[2]};
            double nnode = (double)(n + 1);
            double r = p->connect_circle_dist * sqrt(log(nnode) / nnode);
            if (p->expand_dis < r) r = p->expand_dis;
            double r2 = sq_libm(r);
            int n_near = 0;
            for (int i = 0; i < n; i++) dl[i] = sq(mode, x[i] - nw[0]) + sq(mode, y[i] - nw[1]);
            for (int i = 0; i < n; i++)
                if (dl[i] <= r2) {
                    int f = 0;
                    while (dl[f] != dl[i]) f++;
                    near[n_near++] = f;
                }
            truthy = 0;
            int best = -1;
            double mc = INFINITY;
            for (int k = 0; k < n_near; k++) {
                int i = near[k];
                double fi[3] = {x[i], y[i], yaw[i]};
                dub_edge_t e = dubins_edge(p, obs3, fi, nw, buf);
                double c = (e.npts > 1 && e.free_) ? cost[i] + orc_hypot(nw[0] - x[i], nw[1] - y[i]) : INFINITY;
                if (c < mc) { mc = c; best = i; }
            }
            if (best >= 0) {
                double fb[3] = {x[best], y[best], yaw[best]};
                dub_edge_t e = dubins_edge(p, obs3, fb, nw, buf);
                int newi = n;
                x[newi] = e.end[0]; y[newi] = e.end[1]; yaw[newi] = e.end[2]; cost[newi] = mc; parent[newi] = best;
                memcpy(edge_from + 3 * newi, fb, sizeof fb);
                memcpy(edge_to + 3 * newi, nw, sizeof nw);
                n++;
                truthy = 1;
                double cp[3] = {e.end[0], e.end[1], e.end[2]};
                for (int k = 0; k < n_near; k++) { /* rewire (:1741-1775), after the append */
                    int i = near[k];
                    double ti[3] = {x[i], y[i], yaw[i]};
                    dub_edge_t ed = dubins_edge(p, obs3, cp, ti, buf);
                    if (ed.npts <= 1) continue;
                    double ecost = mc + orc_hypot(x[i] - cp[0], y[i] - cp[1]);
                    if (ed.free_ && cost[i] > ecost) {
                        x[i] = ed.end[0]; y[i] = ed.end[1]; yaw[i] = ed.end[2]; cost[i] = ecost; parent[i] = newi;
                        memcpy(edge_from + 3 * i, cp, sizeof cp);
                        memcpy(edge_to + 3 * i, ti, sizeof ti);
                        dub_propagate(n, x, y, cost, parent, i);
                    }
                }
            }
        }
        if (!p->search_until_max_iter && truthy) {
            gi = dub_best_goal(p, n, x, y, yaw, cost);
            if (gi > 0) { it++; done = 1; break; }
        }
    }
    if (!done) gi = dub_best_goal(p, n, x, y, yaw, cost);
    if (gi <= 0) gi = -1; /* `if last_index:` -- index 0 is falsy (rrt_05:1445, :1451) */
    *n_nodes = n; *iters_done = it; *goal_index = gi;
    free(buf); free(dl); free(near);
    return 0;
}

/* `sum([a, b, c])` of three Python floats as CPython >= 3.12 evaluates it (Neumaier's compensated addition after the first
 * item, compensation added at the end; bltinmodule.c builtin_sum_impl) -- rrt_03:1470 costs an edge this way. */
static double py312_sum3(double a, double b, double c) {
    double f = a, comp = 0.0, t;
    t = f + b; comp += fabs(f) >= fabs(b) ? (f - t) + b : (b - t) + f; f = t;
    t = f + c; comp += fabs(f) >= fabs(c) ? (f - t) + c : (c - t) + f; f = t;
    if (comp != 0.0 && isfinite(comp)) f += comp;
    return f;
}

/* RRT-Dubins, rrt_03's `RRT.planning` (:1402-1456): plain RRT whose steer is the whole Dubins course to the sample.
 * A new node is kept iff its end pose is inside the play area (:1437; play = xmin, xmax, ymin, ymax or NULL) and no course
 * point is inside a circle; its cost is the parent's plus sum |course lengths| (:1470).  p->expand_dis and
 * p->connect_circle_dist are unused.  Returns -2 where the reference raises AttributeError (steer returned None while a
 * play area is set, :1626). */
ORC_EXPORT int orc_rrt_dubins_run(const orc_dubins_params_t *p, const double *obs3, const double *stream3, const double *play,
                                  double *x, double *y, double *yaw, double *cost, int32_t *parent, double *edge_from,
                                  double *edge_to, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index) {
    const int mode = p->math_mode;
    int n = 1, it, gi = -1, done = 0, rc = 0;
    double *buf = (double *)malloc(sizeof(double) * 3 * DUB_MAXPTS);
    x[0] = p->sx; y[0] = p->sy; yaw[0] = p->syaw; cost[0] = 0.0; parent[0] = -1;
    for (it = 0; it < p->max_iter; it++) {
        const double *rnd = stream3 + 3 * it;
        int ni = 0;
        double dmin = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(mode, x[i] - rnd[0]) + sq(mode, y[i] - rnd[1]);
            if (d < dmin) { dmin = d; ni = i; }
        }
        double from[3] = {x[ni], y[ni], yaw[ni]};
        int32_t md;
        double lengths[3];
        int np = orc_dubins_plan(from[0], from[1], from[2], rnd[0], rnd[1], rnd[2], p->kappa, 0.1, mode, &md, lengths, NULL, 0);
        const int truthy = np > 1;
        if (!truthy && play) { rc = -2; it++; break; }
        if (truthy) {
            dub_edge_t e = dubins_edge(p, obs3, from, rnd, buf);
            const int inside = !play || !(e.end[0] < play[0] || e.end[0] > play[1] || e.end[1] < play[2] || e.end[1] > play[3]);
            if (inside && e.free_) {
                x[n] = e.end[0]; y[n] = e.end[1]; yaw[n] = e.end[2];
                cost[n] = cost[ni] + py312_sum3(fabs(lengths[0]), fabs(lengths[1]), fabs(lengths[2]));
                parent[n] = ni;
                memcpy(edge_from + 3 * n, from, sizeof from);
                memcpy(edge_to + 3 * n, rnd, 3 * sizeof(double));
                n++;
            }
        }
        if (!p->search_until_max_iter && truthy) {
            gi = dub_best_goal(p, n, x, y, yaw, cost);
            if (gi > 0) { it++; done = 1; break; }
        }
    }
    if (!done && rc == 0) gi = dub_best_goal(p, n, x, y, yaw, cost);
    if (gi <= 0) gi = -1; /* `if last_index:` -- index 0 is falsy (rrt_03:1446, :1452) */
    *n_nodes = n; *iters_done = it; *goal_index = gi;
    free(buf);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* Reeds-Shepp local planner (rs00:73-515 == rrt_06:1021-1437)                             */
/* ------------------------------------------------------------------------------------ */
static inline double m_asin(int m, double x) { return m == ORC_MATH_LIBM ? asin(x) : crm_asin(x); }
static inline double m_hyp(int m, double a, double b) { return m == ORC_MATH_LIBM ? orc_hypot(a, b) : crm_hypot(a, b); }
static double rs_mod2pi(double x) { /* rs00:130-139 */
    double m = copysign(2.0 * ORC_PI, x);
    double v = fmod(x, m);
    if (v == 0.0) v = copysign(0.0, m);
    if (v < -ORC_PI) v += 2.0 * ORC_PI;
    else if (v > ORC_PI) v -= 2.0 * ORC_PI;
    return v;
}
/* base segment types of the 12 path functions (0 = L, 1 = S, 2 = R) and their lengths */
static const int8_t RS_T[12][5] = {{0, 1, 0, -1, -1}, {0, 1, 2, -1, -1}, {0, 2, 0, -1, -1}, {0, 2, 0, -1, -1}, {0, 2, 0, -1, -1},
                                   {0, 2, 0, 2, -1}, {0, 2, 0, 2, -1},   {0, 2, 1, 0, -1},  {0, 2, 1, 2, -1},  {0, 1, 2, 0, -1},
                                   {0, 1, 0, 2, -1}, {0, 2, 1, 0, 2}};
static const int RS_N[12] = {3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5};
/* the path functions of rs00:166-363 in path_functions order; returns 1 and the travel distances, or 0 */
static int rs_word(int m, int f, double x, double y, double phi, double *d) {
    const double pi = ORC_PI;
    double u, t, v, u1, theta, A, zeta, eeta;
    if (f == 0) {
        u = m_hyp(m, x - m_sin(m, phi), y - 1.0 + m_cos(m, phi));
        t = m_atan2(m, y - 1.0 + m_cos(m, phi), x - m_sin(m, phi));
        if (0.0 <= t && t <= pi) {
            v = rs_mod2pi(phi - t);
            if (0.0 <= v && v <= pi) { d[0] = t; d[1] = u; d[2] = v; return 1; }
        }
        return 0;
    }
    if (f == 1) {
        u1 = m_hyp(m, x + m_sin(m, phi), y - 1.0 - m_cos(m, phi));
        double t1 = m_atan2(m, y - 1.0 - m_cos(m, phi), x + m_sin(m, phi));
        u1 = sq(m, u1);
        if (u1 >= 4.0) {
            u = sqrt(u1 - 4.0);
            theta = m_atan2(m, 2.0, u);
            t = rs_mod2pi(t1 + theta);
            v = rs_mod2pi(t - phi);
            if (t >= 0.0 && v >= 0.0) { d[0] = t; d[1] = u; d[2] = v; return 1; }
        }
        return 0;
    }
    if (f == 2 || f == 3 || f == 4 || f == 7 || f == 9) { zeta = x - m_sin(m, phi); eeta = y - 1 + m_cos(m, phi); }
    else { zeta = x + m_sin(m, phi); eeta = y - 1 - m_cos(m, phi); }
    u1 = m_hyp(m, zeta, eeta);
    theta = m_atan2(m, eeta, zeta);
    switch (f) {
        case 2:
            if (u1 <= 4.0) {
                A = m_acos(m, 0.25 * u1);
                t = rs_mod2pi(A + theta + pi / 2); u = rs_mod2pi(pi - 2 * A); v = rs_mod2pi(phi - t - u);
                d[0] = t; d[1] = -u; d[2] = v; return 1;
            }
            return 0;
        case 3:
            if (u1 <= 4.0) {
                A = m_acos(m, 0.25 * u1);
                t = rs_mod2pi(A + theta + pi / 2); u = rs_mod2pi(pi - 2 * A); v = rs_mod2pi(-phi + t + u);
                d[0] = t; d[1] = -u; d[2] = -v; return 1;
            }
            return 0;
        case 4:
            if (u1 <= 4.0) {
                u = m_acos(m, 1 - sq(m, u1) * 0.125);
                A = m_asin(m, 2 * m_sin(m, u) / u1);
                t = rs_mod2pi(-A + theta + pi / 2); v = rs_mod2pi(t - u - phi);
                d[0] = t; d[1] = u; d[2] = -v; return 1;
            }
            return 0;
        case 5:
            if (u1 <= 2) {
                A = m_acos(m, (u1 + 2) * 0.25);
                t = rs_mod2pi(theta + A + pi / 2); u = rs_mod2pi(A); v = rs_mod2pi(phi - t + 2 * u);
                if (t >= 0 && u >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = -u; d[3] = -v; return 1; }
            }
            return 0;
        case 6: {
            double u2 = (20 - sq(m, u1)) / 16;
            if (0 <= u2 && u2 <= 1) {
                u = m_acos(m, u2);
                A = m_asin(m, 2 * m_sin(m, u) / u1);
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -u; d[2] = -u; d[3] = v; return 1; }
            }
            return 0;
        }
        case 7:
            if (u1 >= 2.0) {
                u = sqrt(sq(m, u1) - 4) - 2;
                A = m_atan2(m, 2, sqrt(sq(m, u1) - 4));
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi + pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -v; return 1; }
            }
            return 0;
        case 8:
            if (u1 >= 2.0) {
                t = rs_mod2pi(theta + pi / 2); u = u1 - 2; v = rs_mod2pi(phi - t - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -v; return 1; }
            }
            return 0;
        case 9:
            if (u1 >= 2.0) {
                u = sqrt(sq(m, u1) - 4) - 2;
                A = m_atan2(m, sqrt(sq(m, u1) - 4), 2);
                t = rs_mod2pi(theta - A + pi / 2); v = rs_mod2pi(t - phi - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = pi / 2; d[3] = -v; return 1; }
            }
            return 0;
        case 10:
            if (u1 >= 2.0) {
                t = rs_mod2pi(theta); u = u1 - 2; v = rs_mod2pi(phi - t - pi / 2);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = u; d[2] = pi / 2; d[3] = -v; return 1; }
            }
            return 0;
        default:
            if (u1 >= 4.0) {
                u = sqrt(sq(m, u1) - 4) - 4;
                A = m_atan2(m, 2, sqrt(sq(m, u1) - 4));
                t = rs_mod2pi(theta + A + pi / 2); v = rs_mod2pi(t - phi);
                if (t >= 0 && v >= 0) { d[0] = t; d[1] = -pi / 2; d[2] = -u; d[3] = -pi / 2; d[4] = v; return 1; }
            }
            return 0;
    }
}
/* reeds_shepp_path_planning (rs00:496-515).  Returns the number of course points (0 = None); types [5] (0 L, 1 S, 2 R,
 * -1 pad), lengths [5] (already / maxc), pts [max_pts][4] = x, y, yaw, direction; *n_paths = len(paths) */
ORC_EXPORT int orc_reeds_shepp(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                               double step_size, int mode, int32_t *types, double *lengths, int32_t *n_seg, double *best_L,
                               int32_t *n_paths, double *pts, int32_t max_pts) {
    const double dx = gx - sx, dy = gy - sy, dth = gyaw - syaw;
    const double c = m_cos(mode, syaw), s = m_sin(mode, syaw);
    const double x = (c * dx + s * dy) * maxc, y = (-s * dx + c * dy) * maxc;
    const double step = step_size * maxc;
    double pl[48][5], pL[48];
    int pt[48][5], pn[48], np_ = 0;
    for (int f = 0; f < 12; f++)
        for (int k = 0; k < 4; k++) {
            double d[5];
            const double ax = (k & 1) ? -x : x, ay = (k & 2) ? -y : y, aphi = (k == 1 || k == 2) ? -dth : dth;
            if (!rs_word(mode, f, ax, ay, aphi, d)) continue;
            const int n = RS_N[f];
            double tot = 0;
            for (int i = 0; i < n; i++) tot += fabs(d[i]);
            for (int i = 0; i < n; i++)
                if (0.1 * tot < fabs(d[i]) && fabs(d[i]) < step) { *n_paths = 0; return 0; } /* "Step size too large" */
            int ty[5];
            for (int i = 0; i < n; i++) {
                if (k == 1 || k == 3) d[i] = -d[i];                                 /* timeflip */
                ty[i] = (k >= 2 && RS_T[f][i] != 1) ? 2 - RS_T[f][i] : RS_T[f][i];  /* reflect */
            }
            double L = 0.0;
            for (int i = 0; i < n; i++) L += fabs(d[i]);
            int same = 0;
            for (int j = 0; j < np_ && !same; j++) {
                if (pn[j] != n) continue;
                int eq = 1;
                for (int i = 0; i < n; i++) eq &= pt[j][i] == ty[i];
                if (eq && (pL[j] - L) <= step) same = 1;
            }
            if (same || L <= step) continue;
            for (int i = 0; i < n; i++) { pl[np_][i] = d[i]; pt[np_][i] = ty[i]; }
            pn[np_] = n; pL[np_] = L; np_++;
        }
    *n_paths = np_;
    if (np_ == 0) return 0;
    int best = 0;
    for (int j = 1; j < np_; j++)
        if (fabs(pL[j] / maxc) < fabs(pL[best] / maxc)) best = j;
    const int n = pn[best];
    *n_seg = n; *best_L = pL[best] / maxc;
    for (int i = 0; i < 5; i++) { types[i] = i < n ? pt[best][i] : -1; lengths[i] = i < n ? pl[best][i] / maxc : 0.0; }
    /* generate_local_course (rs00:431-447) + the world transform of calc_paths (:481-487) */
    const double cm = m_cos(mode, -syaw), sm = m_sin(mode, -syaw);
    double ox = 0.0, oy = 0.0, oyaw = 0.0;
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const double length = pl[best][i];
        const int type = pt[best][i];
        const double dd = length >= 0.0 ? step : -step;
        long na = length != 0.0 ? (long)ceil((length - 0.0) / dd) : 0; /* np.arange(0.0, length, d_dist) */
        if (na < 0) na = 0;
        double lx = 0, ly = 0, lyaw = 0;
        for (long j = 0; j <= na; j++) {
            const double dist = j < na ? 0.0 + (double)j * dd : length;
            if (type == 1) {
                lx = ox + dist / maxc * m_cos(mode, oyaw);
                ly = oy + dist / maxc * m_sin(mode, oyaw);
                lyaw = oyaw;
            } else {
                const double ldx = m_sin(mode, dist) / maxc;
                const double ldy = type == 0 ? (1.0 - m_cos(mode, dist)) / maxc : (1.0 - m_cos(mode, dist)) / -maxc;
                lyaw = type == 0 ? oyaw + dist : oyaw - dist;
                const double gdx = m_cos(mode, -oyaw) * ldx + m_sin(mode, -oyaw) * ldy;
                const double gdy = -m_sin(mode, -oyaw) * ldx + m_cos(mode, -oyaw) * ldy;
                lx = ox + gdx; ly = oy + gdy;
            }
            if (cnt < max_pts) {
                pts[4 * cnt] = cm * lx + sm * ly + sx;
                pts[4 * cnt + 1] = -sm * lx + cm * ly + sy;
                pts[4 * cnt + 2] = angle_mod_pi(lyaw + syaw);
                pts[4 * cnt + 3] = length > 0.0 ? 1.0 : -1.0;
            }
            cnt++;
        }
        ox = lx; oy = ly; oyaw = lyaw;
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------ */
/* RRT*-Reeds-Shepp planning loop (rrt_06:1444-1913)                                      */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double sx, sy, syaw, gx, gy, gyaw;
    double expand_dis, robot_radius, connect_circle_dist, kappa, goal_yaw_th, goal_xy_th, step_size;
    int32_t max_iter, n_obs, search_until_max_iter, math_mode;
    int32_t cost_mode, pad_; /* 0: rrt_06 (Euclidean costs), 1: rrt_10:1005-1207 (Reeds-Shepp length costs, calc_new_cost :1153-1161) */
} orc_rs_params_t;
#define RS_MAXPTS 8192
typedef struct { int npts, free_; double end[3], lsum; } rs_edge_t;

/* steer (rrt_06:1584-1604) + check_collision (:1749-1762) of one edge */
static rs_edge_t rs_edge(const orc_rs_params_t *p, const double *obs3, const double *f, const double *t, double *buf) {
    rs_edge_t e;
    int32_t types[5], nseg = 0, npaths = 0;
    double lengths[5], bl = 0.0;
    int n = orc_reeds_shepp(f[0], f[1], f[2], t[0], t[1], t[2], p->kappa, p->step_size, p->math_mode, types, lengths, &nseg,
                            &bl, &npaths, buf, RS_MAXPTS);
    e.npts = n; e.free_ = 0; e.lsum = 0.0;
    e.end[0] = e.end[1] = e.end[2] = 0.0;
    if (n == 0) return e;
    if (n > RS_MAXPTS) n = RS_MAXPTS;
    e.end[0] = buf[4 * (n - 1)]; e.end[1] = buf[4 * (n - 1) + 1]; e.end[2] = buf[4 * (n - 1) + 2];
    for (int i = 0; i < nseg; i++) e.lsum = e.lsum + fabs(lengths[i]); /* sum([abs(l) ...]) over np.float64 items: plain adds */
    int ok = 1;
    for (int o = 0; o < p->n_obs && ok; o++) {
        double ox = obs3[3 * o], oy = obs3[3 * o + 1], size = obs3[3 * o + 2], mn = INFINITY;
        for (int k = 0; k < n; k++) {
            double dx = ox - buf[4 * k], dy = oy - buf[4 * k + 1], dd = dx * dx + dy * dy;
            if (dd < mn) mn = dd;
        }
        if (mn <= sq_libm(size + p->robot_radius)) ok = 0;
    }
    e.free_ = ok;
    return e;
}
/* propagate_cost_to_leaves with rrt_10's calc_new_cost (:572-577, :1153-1161): parent cost + Reeds-Shepp length, inf when
 * there is no course */
static void rs_propagate(const orc_rs_params_t *p, const double *obs3, int n, const double *x, const double *y, const double *yaw,
                         double *cost, const int32_t *parent, int q, double *buf) {
    for (int c = 0; c < n; c++)
        if (parent[c] == q) {
            double f[3] = {x[q], y[q], yaw[q]}, t[3] = {x[c], y[c], yaw[c]};
            rs_edge_t e = rs_edge(p, obs3, f, t, buf);
            cost[c] = e.npts > 0 ? cost[q] + e.lsum : INFINITY;
            rs_propagate(p, obs3, n, x, y, yaw, cost, parent, c, buf);
        }
}
static int rs_best_goal(const orc_rs_params_t *p, int n, const double *x, const double *y, const double *yaw, const double *cost) {
    int gi = -1;
    double mc = INFINITY;
    for (int i = 0; i < n; i++)
        if (orc_hypot(x[i] - p->gx, y[i] - p->gy) <= p->goal_xy_th && fabs(yaw[i] - p->gyaw) <= p->goal_yaw_th && cost[i] < mc) {
            mc = cost[i]; gi = i;
        }
    return gi;
}
/* arrays sized 2 * max_iter + 1 (try_goal_path can append a second node per iteration) */
ORC_EXPORT int orc_rrtstar_rs_run(const orc_rs_params_t *p, const double *obs3, const double *stream3, double *x, double *y,
                                  double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                                  int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index) {
    const int mode = p->math_mode;
    int cap = 2 * p->max_iter + 1, n = 1, it, gi = -1, done = 0;
    double *buf = (double *)malloc(sizeof(double) * 4 * RS_MAXPTS);
    double *dl = (double *)malloc(sizeof(double) * cap);
    int *near = (int *)malloc(sizeof(int) * cap);
    const double goal[3] = {p->gx, p->gy, p->gyaw};
    x[0] = p->sx; y[0] = p->sy; yaw[0] = p->syaw; cost[0] = 0.0; parent[0] = -1;
    for (it = 0; it < p->max_iter; it++) {
        const double *rnd = stream3 + 3 * it;
        int ni = 0;
        double dmin = INFINITY;
        for (int i = 0; i < n; i++) {
            double d = sq(mode, x[i] - rnd[0]) + sq(mode, y[i] - rnd[1]);
            if (d < dmin) { dmin = d; ni = i; }
        }
        double from[3] = {x[ni], y[ni], yaw[ni]};
        rs_edge_t e0 = rs_edge(p, obs3, from, rnd, buf);
        int truthy = e0.npts > 0;
        if (e0.npts > 0 && e0.free_) {
            double nw[3] = {e0.end[0], e0.end[1], e0.end[2]};
            double nnode = (double)(n + 1);
            double r = p->connect_circle_dist * sqrt(log(nnode) / nnode);
            if (p->expand_dis < r) r = p->expand_dis;
            double r2 = sq_libm(r);
            int n_near = 0;
            for (int i = 0; i < n; i++) dl[i] = sq(mode, x[i] - nw[0]) + sq(mode, y[i] - nw[1]);
            for (int i = 0; i < n; i++)
                if (dl[i] <= r2) {
                    int f = 0;
                    while (dl[f] != dl[i]) f++;
                    near[n_near++] = f;
                }
            truthy = 0;
            int best = -1;
            double mc = INFINITY;
            for (int k = 0; k < n_near; k++) {
                int i = near[k];
                double fi[3] = {x[i], y[i], yaw[i]};
                rs_edge_t e = rs_edge(p, obs3, fi, nw, buf);
                double c = (e.npts > 0 && e.free_) ? cost[i] + (p->cost_mode ? e.lsum : orc_hypot(nw[0] - x[i], nw[1] - y[i])) : INFINITY;
                if (c < mc) { mc = c; best = i; }
            }
            if (best >= 0) {
                double fb[3] = {x[best], y[best], yaw[best]};
                rs_edge_t e = rs_edge(p, obs3, fb, nw, buf);
                int newi = n;
                x[newi] = e.end[0]; y[newi] = e.end[1]; yaw[newi] = e.end[2]; cost[newi] = mc; parent[newi] = best;
                memcpy(edge_from + 3 * newi, fb, sizeof fb);
                memcpy(edge_to + 3 * newi, nw, sizeof nw);
                n++;
                truthy = 1;
                double cp[3] = {e.end[0], e.end[1], e.end[2]};
                for (int k = 0; k < n_near; k++) { /* rewire (:1865-1899), after the append */
                    int i = near[k];
                    double ti[3] = {x[i], y[i], yaw[i]};
                    rs_edge_t ed = rs_edge(p, obs3, cp, ti, buf);
                    if (ed.npts == 0) continue;
                    double ecost = mc + (p->cost_mode ? ed.lsum : orc_hypot(x[i] - cp[0], y[i] - cp[1]));
                    if (ed.free_ && cost[i] > ecost) {
                        x[i] = ed.end[0]; y[i] = ed.end[1]; yaw[i] = ed.end[2]; cost[i] = ecost; parent[i] = newi;
                        memcpy(edge_from + 3 * i, cp, sizeof cp);
                        memcpy(edge_to + 3 * i, ti, sizeof ti);
                        if (p->cost_mode) rs_propagate(p, obs3, n, x, y, yaw, cost, parent, i, buf);
                        else dub_propagate(n, x, y, cost, parent, i);
                    }
                }
                /* try_goal_path (:1572-1582): from the node as it is now */
                double np_[3] = {x[newi], y[newi], yaw[newi]};
                rs_edge_t eg = rs_edge(p, obs3, np_, goal, buf);
                if (eg.npts > 0 && eg.free_) {
                    x[n] = eg.end[0]; y[n] = eg.end[1]; yaw[n] = eg.end[2]; cost[n] = cost[newi] + eg.lsum; parent[n] = newi;
                    memcpy(edge_from + 3 * n, np_, sizeof np_);
                    memcpy(edge_to + 3 * n, goal, sizeof goal);
                    n++;
                }
            }
        }
        if (!p->search_until_max_iter && truthy) {
            gi = rs_best_goal(p, n, x, y, yaw, cost);
            if (gi > 0) { it++; done = 1; break; }
        }
    }
    if (!done) gi = rs_best_goal(p, n, x, y, yaw, cost);
    if (gi <= 0) gi = -1; /* `if last_index:` -- index 0 is falsy */
    *n_nodes = n; *iters_done = it; *goal_index = gi;
    free(buf); free(dl); free(near);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Closed-loop RRT* (rrt_10:1215-1582): unicycle model + pure pursuit over one course and  */
/* the feasibility checks of check_tracking_path_is_feasible.  Constants: rrt_10:1592-1607 */
/* ------------------------------------------------------------------------------------ */
static inline double m_tan(int m, double x) { return m == ORC_MATH_LIBM ? tan(x) : crm_tan(x); }
static inline double m_nphyp(int m, double a, double b) { return m == ORC_MATH_LIBM ? hypot(a, b) : crm_hypot(a, b); } /* np.hypot */
#define CL_DT 0.05
#define CL_WB 0.9          /* L */
#define CL_ACCEL_MAX 5.0
#define CL_KP 2.0
#define CL_LF 0.5
#define CL_T 100.0
#define CL_GOAL_DIS 0.5
#define CL_STOP_SPEED 0.5
#define CL_EXTEND 6        /* int(Lf / 0.1) + 1 */
typedef struct {
    double target_speed, yaw_th, invalid_travel_ratio, robot_radius;
    int32_t n_obs, math_mode, traj_cap, pad_;
} orc_cl_params_t;

/* calc_target_index (:1285-1304) */
static int cl_target_index(int m, double sx, double sy, const double *cx, const double *cy, int n, double *mindis) {
    int ind = 0;
    double best = INFINITY;
    for (int i = 0; i < n; i++) {
        double d = m_nphyp(m, sx - cx[i], sy - cy[i]);
        if (d < best) { best = d; ind = i; }
    }
    *mindis = best;
    double le = 0.0;
    while (CL_LF > le && ind + 1 < n) {
        le += m_hyp(m, cx[ind + 1] - cx[ind], cy[ind + 1] - cy[ind]);
        ind++;
    }
    return ind;
}

/* course3: the final course in driving order (start -> goal), rows x, y, yaw, n_course >= 3.
 * traj7 [traj_cap][7] = x, y, yaw, v, t, a, d of closed_loop_prediction (yaw after the extra angle_mod of :1533);
 * bits: 1 goal not reached, 2 final angle, 4 too long, 8 collision (0 = feasible).  Returns 0, or -1 when traj_cap is too small. */
ORC_EXPORT int orc_closed_loop(const orc_cl_params_t *p, int32_t n_course, const double *course3, const double *obs3,
                               double *traj7, int32_t *n_traj, int32_t *bits_out) {
    const int m = p->math_mode;
    const double steer_max = 40.0 * (ORC_PI / 180.0); /* np.deg2rad(40.0) = 40 * (pi / 180) */
    int n = n_course + CL_EXTEND;
    double *cx = (double *)malloc(sizeof(double) * 4 * (size_t)n), *cy = cx + n, *cyaw = cy + n, *sp = cyaw + n;
    for (int i = 0; i < n_course; i++) { cx[i] = course3[3 * i]; cy[i] = course3[3 * i + 1]; cyaw[i] = course3[3 * i + 2]; }
    const double goal[3] = {cx[n_course - 1], cy[n_course - 1], cyaw[n_course - 1]};
    { /* extend_path (:1432-1447) */
        const int l = n_course - 1;
        const double md = m_atan2(m, cy[l] - cy[l - 2], cx[l] - cx[l - 2]);
        const int back = fabs(md - cyaw[l]) >= ORC_PI / 2.0;
        const double idl = back ? -0.1 : 0.1;
        for (int k = n_course; k < n; k++) {
            cx[k] = cx[k - 1] + idl * m_cos(m, cyaw[k - 1]);
            cy[k] = cy[k - 1] + idl * m_sin(m, cyaw[k - 1]);
            cyaw[k] = cyaw[k - 1];
        }
    }
    { /* set_stop_point (:1375-1419) */
        int forward = 1, back = 0;
        for (int i = 0; i < n; i++) sp[i] = p->target_speed;
        for (int i = 0; i < n - 1; i++) {
            const double dx = cx[i + 1] - cx[i], dy = cy[i + 1] - cy[i];
            back = fabs(m_atan2(m, dy, dx) - cyaw[i]) >= ORC_PI / 2.0;
            if (dx == 0.0 && dy == 0.0) continue;
            sp[i] = back ? -p->target_speed : p->target_speed;
            if (back && forward) { sp[i] = 0.0; forward = 0; }
            else if (!back && !forward) { sp[i] = 0.0; forward = 1; }
        }
        sp[0] = 0.0;
        sp[n - 1] = back ? -CL_STOP_SPEED : CL_STOP_SPEED;
    }
    /* closed_loop_prediction (:1307-1372) */
    double sx = -0.0, sy = -0.0, syaw = 0.0, sv = 0.0, time = 0.0, dis = 0.0, travel = 0.0;
    int cnt = 0, find_goal = 0, rc = 0;
    const double maxdis = 0.5, dcap = maxdis - 0.1;
    double *r = traj7;
    r[0] = sx; r[1] = sy; r[2] = syaw; r[3] = sv; r[4] = 0.0; r[5] = 0.0; r[6] = 0.0;
    cnt = 1;
    int target_ind = cl_target_index(m, sx, sy, cx, cy, n, &dis);
    while (CL_T >= time) {
        int ind = cl_target_index(m, sx, sy, cx, cy, n, &dis); /* pure_pursuit_control (:1255-1283) */
        if (target_ind >= ind) ind = target_ind;
        double tx, ty;
        if (ind < n) { tx = cx[ind]; ty = cy[ind]; }
        else { tx = cx[n - 1]; ty = cy[n - 1]; ind = n - 1; }
        double alpha = m_atan2(m, ty - sy, tx - sx) - syaw;
        if (sv <= 0.0) alpha = ORC_PI - alpha;
        double di = m_atan2(m, 2.0 * CL_WB * m_sin(m, alpha) / CL_LF, 1.0);
        if (di > steer_max) di = steer_max;
        else if (di < -steer_max) di = -steer_max;
        target_ind = ind;
        double ts = sp[target_ind];
        ts = ts * (maxdis - (dcap < dis ? dcap : dis)) / maxdis;
        double ai = CL_KP * (ts - sv); /* PIDControl (:1243-1252) */
        if (ai > CL_ACCEL_MAX) ai = CL_ACCEL_MAX;
        else if (ai < -CL_ACCEL_MAX) ai = -CL_ACCEL_MAX;
        const double nx = sx + sv * m_cos(m, syaw) * CL_DT; /* update (:1224-1232) */
        const double ny = sy + sv * m_sin(m, syaw) * CL_DT;
        const double nyaw = angle_mod_pi(syaw + sv / CL_WB * m_tan(m, di) * CL_DT);
        sv = sv + ai * CL_DT;
        sx = nx; sy = ny; syaw = nyaw;
        if (fabs(sv) <= CL_STOP_SPEED && target_ind <= n - 2) target_ind++;
        time = time + CL_DT;
        if (m_hyp(m, sx - goal[0], sy - goal[1]) <= CL_GOAL_DIS) { find_goal = 1; break; }
        if (cnt >= p->traj_cap) { rc = -1; break; }
        r = traj7 + 7 * (size_t)cnt;
        r[0] = sx; r[1] = sy; r[2] = syaw; r[3] = sv; r[4] = time; r[5] = ai; r[6] = di;
        cnt++;
    }
    /* check_tracking_path_is_feasible (:1521-1559) */
    int bits = find_goal ? 0 : 1;
    for (int k = 0; k < cnt; k++) traj7[7 * (size_t)k + 2] = angle_mod_pi(traj7[7 * (size_t)k + 2]);
    if (fabs(traj7[7 * (size_t)(cnt - 1) + 2] - goal[2]) >= p->yaw_th * 10.0) bits |= 2;
    for (int k = 0; k < cnt; k++) travel = travel + fabs(traj7[7 * (size_t)k + 3]);
    travel = CL_DT * travel;
    double origin = 0.0;
    for (int i = 0; i < n - 1; i++) origin = origin + m_nphyp(m, cx[i + 1] - cx[i], cy[i + 1] - cy[i]);
    if (travel / origin >= p->invalid_travel_ratio) bits |= 4;
    for (int o = 0; o < p->n_obs; o++) { /* check_collision (rrt_10:279-293) */
        const double ox = obs3[3 * o], oy = obs3[3 * o + 1], lim = sq_libm(obs3[3 * o + 2] + p->robot_radius);
        double mn = INFINITY;
        for (int k = 0; k < cnt; k++) {
            const double dx = ox - traj7[7 * (size_t)k], dy = oy - traj7[7 * (size_t)k + 1], dd = dx * dx + dy * dy;
            if (dd < mn) mn = dd;
        }
        if (mn <= lim) { bits |= 8; break; }
    }
    *n_traj = cnt; *bits_out = bits;
    free(cx);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* BIT* (rrt_08:138-611, RTree :29-135): batch informed trees as the reference implements  */
/* them -- ids are cells of a 0.01 grid over randArea, every cost is taken between the      */
/* QUANTISED coordinates of two ids (only expand_vertex's radius test reads a sample's raw  */
/* coordinates), dicts and lists keep Python's insertion order.  Quirks kept: informed_sample */
/* draws m + 1 samples; best_edge_queue_value is the MAXIMUM (sort(reverse=True)[0], :455-463); */
/* add_vertex_to_edge_queue pairs vid with itself and never appends (:503-522); the two      */
/* `continue`s in plan skip `iterations += 1` (:286, :293); remove_queue removes from the list */
/* it iterates (:343-351); cMin = |start - goal| / 1.5 (:199-200).                           */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double sx, sy, gx, gy, min_rand, max_rand;
    double lower, resolution, num_cells; /* RTree: lowerLimit = randArea[0], 0.01, np.ceil((upper - lower) / resolution) */
    double rot[4];                       /* 2 x 2 block of C (:205-213), host-evaluated */
    int32_t max_iter, n_obs, math_mode, n_draws;
} orc_bit_params_t;

typedef struct { double *id, *x, *y; int n, cap; } bit_samples_t;
typedef struct { const orc_bit_params_t *p; } bit_ctx_t;

static double bit_id_of(const orc_bit_params_t *p, double x, double y) { /* real_world_to_node_id (:65-101) */
    const double c0 = (double)(long long)rint((x - p->lower) / p->resolution);
    const double c1 = (double)(long long)rint((y - p->lower) / p->resolution);
    return (0.0 + c1 * p->num_cells) + c0;
}
static void bit_coord_of(const orc_bit_params_t *p, double id, double *x, double *y) { /* node_id_to_real_world_coord (:115-135) */
    const double c1 = floor(id / p->num_cells);
    id = id - c1 * p->num_cells;
    const double c0 = floor(id / 1.0);
    *x = p->lower + p->resolution * c0;
    *y = p->lower + p->resolution * c1;
}
/* np.linalg.norm(v, 2) of a 2-vector = sqrt(v.dot(v)): BLAS ddot, whose scalar tail loop is compiled with FMA contraction on
 * this platform -> sqrt(fma(y, y, x * x)) (checked against numpy on 20 000 random vectors: 20 000 / 20 000, the plain
 * x*x + y*y form matches 16 669) */
static inline double bit_norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }
static double bit_dist(const orc_bit_params_t *p, double a, double b) { /* np.linalg.norm(coord(b) - coord(a), 2) */
    double ax, ay, bx, by;
    bit_coord_of(p, a, &ax, &ay);
    bit_coord_of(p, b, &bx, &by);
    const double dx = bx - ax, dy = by - ay;
    return bit_norm2(dx, dy);
}
static void bit_samples_set(bit_samples_t *s, double id, double x, double y) {
    for (int i = 0; i < s->n; i++)
        if (s->id[i] == id) { s->x[i] = x; s->y[i] = y; return; }
    if (s->n == s->cap) {
        s->cap *= 2;
        s->id = (double *)realloc(s->id, sizeof(double) * s->cap);
        s->x = (double *)realloc(s->x, sizeof(double) * s->cap);
        s->y = (double *)realloc(s->y, sizeof(double) * s->cap);
    }
    s->id[s->n] = id; s->x[s->n] = x; s->y[s->n] = y; s->n++;
}
static void bit_samples_del(bit_samples_t *s, double id) {
    for (int i = 0; i < s->n; i++)
        if (s->id[i] == id) {
            for (int j = i; j + 1 < s->n; j++) { s->id[j] = s->id[j + 1]; s->x[j] = s->x[j + 1]; s->y[j] = s->y[j + 1]; }
            s->n--;
            return;
        }
}
/* informed_sample (:385-419) merged into self.samples; returns 0, or -1 when the draws run out */
static int bit_informed_sample(const orc_bit_params_t *p, bit_samples_t *s, int m, double c_max, double c_min, double xc,
                               double yc, const double *draws, int *used) {
    const int mode = p->math_mode;
    for (int i = 0; i < m + 1; i++) {
        double rx, ry;
        if (*used + 2 > p->n_draws) return -1;
        const double u0 = draws[*used], u1 = draws[*used + 1];
        *used += 2;
        if (c_max < INFINITY) {
            const double r0 = c_max / 2.0, r1 = sqrt(sq(mode, c_max) - sq(mode, c_min)) / 2.0;
            double a = u0, b = u1; /* sample_unit_ball (:421-430) */
            if (b < a) { double t = a; a = b; b = t; }
            const double ang = 2 * 3.141592653589793 * a / b;
            const double bx = b * m_cos(mode, ang), by = b * m_sin(mode, ang);
            const double m00 = p->rot[0] * r0, m01 = p->rot[1] * r1, m10 = p->rot[2] * r0, m11 = p->rot[3] * r1;
            rx = fma(m00, bx, m01 * by) + xc; /* np.dot(np.dot(C, L), xBall) + xCenter, as in orc_informed_run */
            ry = fma(m10, bx, m11 * by) + yc;
        } else { /* sample_free_space (:432-435): random.uniform(a, b) = a + (b - a) * u */
            rx = p->min_rand + (p->max_rand - p->min_rand) * u0;
            ry = p->min_rand + (p->max_rand - p->min_rand) * u1;
        }
        bit_samples_set(s, bit_id_of(p, rx, ry), rx, ry);
    }
    return 0;
}

/* outputs: vertices [vcap] ids in insertion order, g_vertices, edges [vcap][2], parent_of [vcap][2] (id, parent id; in
 * the dict's insertion order), samples (ids + xy) [scap], vertex_queue [vcap], edge_queue [ecap][2], path [pcap][2];
 * counts[8] = n_vertices, n_edges, n_parent, n_samples, n_vq, n_eq, path_len, draws_used;  g_goal.
 * returns 0, 1 = the reference raises IndexError (both queues empty), 2 = the reference loops forever (every edge of the
 * first batch is skipped, so `iterations` stays 0 and the same batch is replayed), -1 = a capacity or the draw stream ran out. */
ORC_EXPORT int orc_bitstar_plan(const orc_bit_params_t *p, const double *obs3, const double *draws, int32_t vcap, int32_t scap,
                                int32_t ecap, int32_t pcap, double *vertices, double *g_vertices, double *edges, double *parent_of,
                                double *sample_ids, double *sample_xy, double *vertex_queue, double *edge_queue, double *path,
                                int32_t *counts, double *g_goal_out) {
    const int mode = p->math_mode;
    int rc = 0, used = 0;
    bit_samples_t S;
    S.cap = 512; S.n = 0;
    S.id = (double *)malloc(sizeof(double) * S.cap); S.x = (double *)malloc(sizeof(double) * S.cap); S.y = (double *)malloc(sizeof(double) * S.cap);
    /* score table: slot 0 = goal, slot 1 = start, then the tree vertices as they appear */
    const int kcap = vcap + 2;
    double *k_id = (double *)malloc(sizeof(double) * 4 * kcap), *k_g = k_id + kcap, *k_f = k_g + kcap, *k_par = k_f + kcap;
    int *k_haspar = (int *)calloc(kcap, sizeof(int)), *par_order = (int *)malloc(sizeof(int) * kcap), n_k = 0, n_par = 0;
    int *old = (int *)calloc(kcap, sizeof(int));
    /* tree: vertex list (slot indexes into k_*), add_edge calls in order */
    int *tv = (int *)malloc(sizeof(int) * kcap), n_v = 0;
    int *te_v = (int *)malloc(sizeof(int) * 2 * kcap), *te_x = te_v + kcap, n_te = 0;
    int *vq = (int *)malloc(sizeof(int) * kcap), n_vq = 0;
    int eqcap = 4096, n_eq = 0;
    int *eq_v = (int *)malloc(sizeof(int) * eqcap);
    double *eq_x = (double *)malloc(sizeof(double) * eqcap); /* edge = (tree vertex slot, target id) */
    int *open_ = (int *)malloc(sizeof(int) * 2 * kcap), *closed = open_ + kcap;

    const double start_id = bit_id_of(p, p->sx, p->sy), goal_id = bit_id_of(p, p->gx, p->gy);
#define BIT_SLOT(idv, out) do { out = -1; for (int q_ = 0; q_ < n_k; q_++) if (k_id[q_] == (idv)) { out = q_; break; } } while (0)
    /* setup_planning (:186-216) */
    bit_samples_set(&S, goal_id, p->gx, p->gy);
    k_id[0] = goal_id; k_g[0] = INFINITY; k_f[0] = 0.0; n_k = 1;
    int s_slot;
    BIT_SLOT(start_id, s_slot);
    if (s_slot < 0) { s_slot = n_k; k_id[n_k] = start_id; n_k++; }
    tv[n_v++] = s_slot;
    k_g[s_slot] = 0.0; k_f[s_slot] = bit_dist(p, start_id, goal_id);
    const double c_min = m_hyp(mode, p->sx - p->gx, p->sy - p->gy) / 1.5;
    const double xc = (p->sx + p->gx) / 2.0, yc = (p->sy + p->gy) / 2.0;
    if (bit_informed_sample(p, &S, 200, INFINITY, c_min, xc, yc, draws, &used)) rc = -1;
    double r = INFINITY;
    int iterations = 0, found_goal = 0, n_batches = 0, n_reset = 0, n_skipped = 0, n_expand = 0;
    while (rc == 0 && iterations < p->max_iter) {
        /* setup_sample (:218-234) */
        if (n_vq == 0 && n_eq == 0) {
            r = 2.0;
            n_batches++;
            if (n_batches >= 2 && iterations == 0) { rc = 2; break; } /* same samples, same edges, all skipped again: the reference never returns */
            if (iterations != 0) {
                int m = 100;
                if (found_goal) { m = 200; S.n = 0; bit_samples_set(&S, goal_id, p->gx, p->gy); }
                if (bit_informed_sample(p, &S, m, k_g[0], c_min, xc, yc, draws, &used)) { rc = -1; break; }
            }
            for (int i = 0; i < n_v; i++) old[tv[i]] = 1;
            for (int i = 0; i < n_v; i++) {
                int in = 0;
                for (int j = 0; j < n_vq; j++) if (vq[j] == tv[i]) { in = 1; break; }
                if (!in) vq[n_vq++] = tv[i];
            }
        }
        /* expand while best_vertex_queue_value() <= best_edge_queue_value() (:244-246) */
        for (;;) {
            double vmin = INFINITY, emax = INFINITY;
            int vbest = -1;
            for (int j = 0; j < n_vq; j++) {
                const double v = k_g[vq[j]] + bit_dist(p, k_id[vq[j]], goal_id);
                if (vbest < 0 || v < vmin) { vmin = v; vbest = j; }
            }
            if (n_vq == 0) vmin = INFINITY;
            if (n_eq > 0) {
                emax = -INFINITY;
                for (int j = 0; j < n_eq; j++) {
                    const double v = k_g[eq_v[j]] + bit_dist(p, k_id[eq_v[j]], eq_x[j]) + bit_dist(p, eq_x[j], goal_id);
                    if (v > emax) emax = v;
                }
            }
            if (!(vmin <= emax)) break;
            if (n_vq == 0) { rc = 1; break; } /* best_in_vertex_queue on an empty list: IndexError */
            /* expand_vertex (:476-501) */
            const int vs = vq[vbest];
            const double vid = k_id[vs];
            n_expand++;
            for (int j = vbest; j + 1 < n_vq; j++) vq[j] = vq[j + 1];
            n_vq--;
            double cx, cy;
            bit_coord_of(p, vid, &cx, &cy);
            const double d_sv = bit_dist(p, start_id, vid);
            for (int i = 0; i < S.n; i++) {
                const double dx = S.x[i] - cx, dy = S.y[i] - cy;
                if (bit_norm2(dx, dy) <= r && S.id[i] != vid) {
                    const double est = d_sv + bit_dist(p, S.id[i], goal_id) + bit_dist(p, vid, S.id[i]);
                    if (est < k_g[0]) {
                        if (n_eq == eqcap) {
                            eqcap *= 2;
                            eq_v = (int *)realloc(eq_v, sizeof(int) * eqcap);
                            eq_x = (double *)realloc(eq_x, sizeof(double) * eqcap);
                        }
                        eq_v[n_eq] = vs; eq_x[n_eq] = S.id[i]; n_eq++;
                    }
                }
            }
            /* add_vertex_to_edge_queue (:503-522) pairs vid with itself: g[vid] + 0 < g[vid] never holds -> no append */
        }
        if (rc) break;
        /* best_in_edge_queue (:469-474): first minimum */
        int eb = 0;
        double ebv = INFINITY;
        for (int j = 0; j < n_eq; j++) {
            const double v = k_g[eq_v[j]] + bit_dist(p, k_id[eq_v[j]], eq_x[j]) + bit_dist(p, eq_x[j], goal_id);
            if (j == 0 || v < ebv) { ebv = v; eb = j; }
        }
        const int e0s = eq_v[eb];
        const double e0 = k_id[e0s], e1 = eq_x[eb];
        for (int j = eb; j + 1 < n_eq; j++) { eq_v[j] = eq_v[j + 1]; eq_x[j] = eq_x[j + 1]; }
        n_eq--;
        const double d01 = bit_dist(p, e0, e1), h1 = bit_dist(p, e1, goal_id);
        const double est_v = k_g[e0s] + d01 + h1;
        const double est_e = bit_dist(p, start_id, e0) + bit_dist(p, e0, e1) + h1;
        const double actual = k_g[e0s] + d01;
        if (est_v < k_g[0] && est_e < k_g[0] && actual < k_g[0]) {
            double fx, fy, tx, ty;
            bit_coord_of(p, e0, &fx, &fy);
            bit_coord_of(p, e1, &tx, &ty);
            /* connect (:359-374): np.linspace samples, stop before the first colliding one */
            const long steps = (long)(bit_dist(p, bit_id_of(p, fx, fy), bit_id_of(p, tx, ty)) * 10);
            const double last_edge = bit_id_of(p, tx, ty);
            long n_free = 0;
            double lx = 0.0, ly = 0.0;
            int none = 0;
            if (steps > 0) {
                const double div = (double)(steps - 1), ddx = tx - fx, ddy = ty - fy;
                const double stx = steps > 1 ? ddx / div : 0.0, sty = steps > 1 ? ddy / div : 0.0;
                for (long i = 0; i < steps; i++) {
                    double px, py;
                    if (steps > 1 && i == steps - 1) { px = tx; py = ty; }
                    else {
                        px = (steps > 1 ? (stx == 0.0 ? ((double)i / div) * ddx : (double)i * stx) : 0.0 * ddx) + fx;
                        py = (steps > 1 ? (sty == 0.0 ? ((double)i / div) * ddy : (double)i * sty) : 0.0 * ddy) + fy;
                    }
                    int hit = 0;
                    for (int o = 0; o < p->n_obs; o++) {
                        const double ex = obs3[3 * o] - px, ey = obs3[3 * o + 1] - py;
                        if (ex * ex + ey * ey <= sq_libm(obs3[3 * o + 2])) { hit = 1; break; }
                    }
                    if (hit) { if (i == 0) none = 1; break; }
                    lx = px; ly = py; n_free++;
                }
            }
            if (none || n_free == 0) { n_skipped++; continue; } /* `continue` before iterations += 1 */
            const double nid = bit_id_of(p, lx, ly);
            int exists;
            BIT_SLOT(nid, exists);
            int in_tree = 0;
            if (exists >= 0) for (int i = 0; i < n_v; i++) if (tv[i] == exists) { in_tree = 1; break; }
            if (in_tree) { n_skipped++; continue; }
            bit_samples_del(&S, nid);
            if (n_v >= vcap) { rc = -1; break; }
            int ns = exists;
            if (ns < 0) { ns = n_k; k_id[n_k] = nid; k_haspar[n_k] = 0; n_k++; }
            tv[n_v++] = ns;
            vq[n_vq++] = ns;
            if (nid == goal_id || e0 == goal_id) found_goal = 1;
            te_v[n_te] = e0s; te_x[n_te] = ns; n_te++;
            const double gsc = bit_dist(p, e0, nid);
            k_g[ns] = gsc + k_g[e0s];
            k_f[ns] = gsc + bit_dist(p, nid, goal_id);
            { /* update_graph (:524-552) */
                int n_open = 0, n_closed = 0;
                open_[n_open++] = s_slot;
                while (n_open) {
                    int bi = 0;
                    for (int j = 1; j < n_open; j++) if (k_f[open_[j]] < k_f[open_[bi]]) bi = j;
                    const int cur = open_[bi];
                    for (int j = bi; j + 1 < n_open; j++) open_[j] = open_[j + 1];
                    n_open--;
                    if (k_id[cur] == goal_id) break;
                    int inc = 0;
                    for (int j = 0; j < n_closed; j++) if (closed[j] == cur) { inc = 1; break; }
                    if (!inc) closed[n_closed++] = cur;
                    for (int c = 0; c < n_te; c++) {
                        int suc;
                        if (te_v[c] == cur) suc = te_x[c];
                        else if (te_x[c] == cur) suc = te_v[c];
                        else continue;
                        int isc = 0;
                        for (int j = 0; j < n_closed; j++) if (closed[j] == suc) { isc = 1; break; }
                        if (isc) continue;
                        const double gs = k_g[cur] + bit_dist(p, k_id[cur], k_id[suc]);
                        int ino = 0;
                        for (int j = 0; j < n_open; j++) if (open_[j] == suc) { ino = 1; break; }
                        if (!ino) open_[n_open++] = suc;
                        else if (gs >= k_g[suc]) continue;
                        k_g[suc] = gs;
                        k_f[suc] = gs + bit_dist(p, k_id[suc], goal_id);
                        if (!k_haspar[suc]) { k_haspar[suc] = 1; par_order[n_par++] = suc; }
                        k_par[suc] = k_id[cur];
                    }
                }
            }
            /* remove_queue(lastEdge, bestEdge) (:343-351): the list is mutated while it is iterated */
            for (int i = 0; i < n_eq; i++) {
                if (eq_x[i] == nid && k_g[ns] + bit_dist(p, eq_x[i], nid) >= k_g[0]) {
                    int le;
                    BIT_SLOT(last_edge, le);
                    if (le >= 0)
                        for (int j = 0; j < n_eq; j++)
                            if (eq_v[j] == le && eq_x[j] == nid) {
                                for (int t = j; t + 1 < n_eq; t++) { eq_v[t] = eq_v[t + 1]; eq_x[t] = eq_x[t + 1]; }
                                n_eq--;
                                break;
                            }
                }
            }
        } else {
            n_eq = 0; n_vq = 0; /* "Nothing good" */
            n_reset++;
        }
        iterations++;
    }
    /* find_final_path (:333-341) */
    int plen = 0;
    if (rc == 0) {
        double cur = goal_id;
        int ok = 1;
        if (plen < pcap) { path[0] = p->gx; path[1] = p->gy; }
        plen = 1;
        while (cur != start_id) {
            double x, y;
            bit_coord_of(p, cur, &x, &y);
            if (plen < pcap) { path[2 * plen] = x; path[2 * plen + 1] = y; }
            plen++;
            int sl;
            BIT_SLOT(cur, sl);
            if (sl < 0 || !k_haspar[sl]) { ok = 0; break; }
            cur = k_par[sl];
        }
        if (!ok) plen = 0;
        else {
            if (plen < pcap) { path[2 * plen] = p->sx; path[2 * plen + 1] = p->sy; }
            plen++;
            if (plen <= pcap)
                for (int i = 0, j = plen - 1; i < j; i++, j--) {
                    double t0 = path[2 * i], t1 = path[2 * i + 1];
                    path[2 * i] = path[2 * j]; path[2 * i + 1] = path[2 * j + 1];
                    path[2 * j] = t0; path[2 * j + 1] = t1;
                }
            else rc = -1;
        }
    }
    for (int i = 0; i < n_v; i++) { vertices[i] = k_id[tv[i]]; g_vertices[i] = k_g[tv[i]]; }
    for (int i = 0; i < n_te; i++) { edges[2 * i] = k_id[te_v[i]]; edges[2 * i + 1] = k_id[te_x[i]]; }
    for (int i = 0; i < n_par; i++) { parent_of[2 * i] = k_id[par_order[i]]; parent_of[2 * i + 1] = k_par[par_order[i]]; }
    int ns_out = S.n < scap ? S.n : scap, ne_out = n_eq < ecap ? n_eq : ecap;
    if (S.n > scap || n_eq > ecap) rc = rc ? rc : -1;
    for (int i = 0; i < ns_out; i++) { sample_ids[i] = S.id[i]; sample_xy[2 * i] = S.x[i]; sample_xy[2 * i + 1] = S.y[i]; }
    for (int i = 0; i < n_vq; i++) vertex_queue[i] = k_id[vq[i]];
    for (int i = 0; i < ne_out; i++) { edge_queue[2 * i] = k_id[eq_v[i]]; edge_queue[2 * i + 1] = eq_x[i]; }
    counts[0] = n_v; counts[1] = n_te; counts[2] = n_par; counts[3] = S.n; counts[4] = n_vq; counts[5] = n_eq; counts[6] = plen;
    counts[7] = used; counts[8] = n_batches; counts[9] = n_reset; counts[10] = n_skipped; counts[11] = n_expand;
    *g_goal_out = k_g[0];
#undef BIT_SLOT
    free(S.id); free(S.x); free(S.y); free(k_id); free(k_haspar); free(par_order); free(old); free(tv); free(te_v); free(vq);
    free(eq_v); free(eq_x); free(open_);
    return rc;
}
