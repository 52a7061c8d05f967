// rrtk_bitstar.cu -- BIT* as rrt_08 implements it (BITStar :138-611 over RTree :29-135), Q independent queries per launch,
// one warp per query.  Every identifier is a cell id of the 0.01 grid over randArea (a double, like the reference's
// np.float64 ids); all costs are taken between the QUANTISED coordinates of two ids, only expand_vertex's radius test
// reads a sample's raw coordinates.  Python's dict / list order is part of the result (first minimum of the queues,
// the order of neighbours), so the containers are ordered arrays in global memory and the lanes split the scans:
//   * vertex / edge queue scoring (best_vertex_queue_value :439-446, best_edge_queue_value :448-457 -- the MAXIMUM, as
//     sort(reverse=True)[0] returns --, best_in_*_queue :459-474): lane-strided, warp argmin / max, first position on ties;
//   * expand_vertex (:476-501): radius search over the sample batch with ballot compaction in dict order;
//   * connect (:359-374) + _collision_check (:376-383): np.linspace points over the lanes, first colliding index;
//   * update_graph (:524-552): the A*-like relabelling stays sequential over the pops, neighbours of a pop in parallel.
// np.linalg.norm(v, 2) of a 2-vector is sqrt(fma(y, y, x*x)) on the reference platform (BLAS ddot, see oracle).
// Quirks kept: m + 1 samples per batch, add_vertex_to_edge_queue never appends (:503-522), the `continue`s that skip
// `iterations += 1` (:286, :293), remove_queue (:343-351), cMin = |start - goal| / 1.5.  FP64, -fmad=false.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"

namespace rrtk {

constexpr int BIT_WARPS_PER_CTA = 4;
constexpr double BIT_DEAD_ID = -1.0e300;

struct BitGrid { double lower, res, nc; };

static __device__ __forceinline__ double bit_id_of(const BitGrid &g, double x, double y) {  // real_world_to_node_id (:65-101)
    const double c0 = (double)(long long)rint((x - g.lower) / g.res);
    const double c1 = (double)(long long)rint((y - g.lower) / g.res);
    return (0.0 + c1 * g.nc) + c0;
}
static __device__ __forceinline__ void bit_coord_of(const BitGrid &g, double id, double *x, double *y) {  // :115-135
    const double c1 = floor(id / g.nc);
    id = id - c1 * g.nc;
    const double c0 = floor(id / 1.0);
    *x = g.lower + g.res * c0;
    *y = g.lower + g.res * c1;
}
static __device__ __forceinline__ double bit_norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }
static __device__ __forceinline__ double bit_dist(const BitGrid &g, double a, double b) {
    double ax, ay, bx, by;
    bit_coord_of(g, a, &ax, &ay);
    bit_coord_of(g, b, &bx, &by);
    return bit_norm2(bx - ax, by - ay);
}

// first position in [0, n) whose id equals `id`, or -1 (uniform result)
static __device__ __noinline__ int bit_find(const double *ids, int n, double id, int lane) {
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int i = b0 + lane;
        const unsigned m = __ballot_sync(FULL, i < n && ids[i] == id);
        if (m) return b0 + __ffs(m) - 1;
    }
    return -1;
}

// Every cost is a function of two ids, so the decoded (quantised) coordinates and the heuristic to the goal are cached next
// to each id when it enters a container: same operations on the same values, evaluated once.
struct BitState {
    double *s_id, *s_x, *s_y;            // samples, dict order; deleted entries: id = BIT_DEAD_ID, x = inf
    double *s_qx, *s_qy, *s_h;           //   quantised coordinates of the id, h(id, goal)
    double *k_id, *k_g, *k_f, *k_par;    // score table: slot 0 = goal, 1 = start, then tree vertices
    double *k_x, *k_y, *k_h;             //   quantised coordinates, h(id, goal)
    int *k_haspar, *par_order, *tv, *te_v, *te_x, *vq, *eq_v, *open_, *flag;
    int *ch_first, *ch_last, *ch_next;   // children of a tree vertex in add_edge order (= its adjacency list minus the parent)
    double *eq_x, *eq_c, *eq_h;          // edge queue: target id, dist(e0, e1), h(e1, goal)
    int n_s, n_k, n_par, n_v, n_te, n_vq, n_eq, n_eq_live;
};

static __device__ __noinline__ void bit_samples_set(BitState &S, const BitGrid &g, int scap, double id, double x, double y,
                                                    int lane, int *status) {
    const int f = bit_find(S.s_id, S.n_s, id, lane);
    if (f >= 0) {
        if (lane == 0) { S.s_x[f] = x; S.s_y[f] = y; }
    } else if (S.n_s < scap) {
        if (lane == 0) {
            double qx, qy;
            bit_coord_of(g, id, &qx, &qy);
            S.s_id[S.n_s] = id; S.s_x[S.n_s] = x; S.s_y[S.n_s] = y;
            S.s_qx[S.n_s] = qx; S.s_qy[S.n_s] = qy; S.s_h[S.n_s] = bit_norm2(qx - S.k_x[0], qy - S.k_y[0]);
        }
        S.n_s++;
    } else {
        *status |= RRTK_BIT_SAMPLE_OVERFLOW;
    }
    __syncwarp();
}

// informed_sample (:385-419) merged into the sample dict; false when the draw stream ran out
static __device__ __noinline__ bool bit_informed_sample(BitState &S, const rrtk_bitstar_params &p, const BitGrid &g, int m,
                                                        double c_max, double c_min, double xc, double yc, const double *rot,
                                                        double min_rand, double max_rand, const double *draws, int *used,
                                                        int lane, int *status) {
    for (int i = 0; i < m + 1; i++) {
        if (*used + 2 > p.n_draws) return false;
        const double u0 = draws[*used], u1 = draws[*used + 1];
        *used += 2;
        double rx, ry;
        if (c_max < CUDART_INF) {
            const double r0 = c_max / 2.0, r1 = sqrt(c_max * c_max - c_min * c_min) / 2.0;
            double a = u0, b = u1;  // sample_unit_ball (:421-430)
            if (b < a) { const double t = a; a = b; b = t; }
            const double ang = 2 * 3.141592653589793 * a / b;
            const double bx = b * crm_cos(ang), by = b * crm_sin(ang);
            const double m00 = rot[0] * r0, m01 = rot[1] * r1, m10 = rot[2] * r0, m11 = rot[3] * r1;
            rx = fma(m00, bx, m01 * by) + xc;
            ry = fma(m10, bx, m11 * by) + yc;
        } else {  // sample_free_space (:432-435)
            rx = min_rand + (max_rand - min_rand) * u0;
            ry = min_rand + (max_rand - min_rand) * u1;
        }
        bit_samples_set(S, g, p.sample_cap, bit_id_of(g, rx, ry), rx, ry, lane, status);
    }
    return true;
}

// value of edge j of the queue: g[e0] + dist(e0, e1) + h(e1, goal)
static __device__ __forceinline__ double bit_edge_value(const BitState &S, int j) {
    return S.k_g[S.eq_v[j]] + S.eq_c[j] + S.eq_h[j];
}

// remove position `pos` from an int list / (int, double) list of length n, keeping the order
static __device__ __noinline__ void bit_erase_int(int *a, int n, int pos, int lane) {
    for (int b0 = pos; b0 + 1 < n; b0 += 32) {
        const int j = b0 + lane;
        int t = 0;
        if (j + 1 < n) t = a[j + 1];
        __syncwarp();
        if (j + 1 < n) a[j] = t;
        __syncwarp();
    }
}
// edge_queue.remove(edge): the slot becomes a tombstone (eq_v = -1) -- order is kept without shifting the queue; n_eq counts
// slots, n_eq_live the edges.  Every scan skips tombstones; "Nothing good" empties the slots; the slots are compacted once at
// the end of the run.
static __device__ __forceinline__ void bit_erase_edge(BitState &S, int pos, int lane) {
    if (lane == 0) S.eq_v[pos] = -1;
    S.n_eq_live--;
    __syncwarp();
}

// update_graph (:524-552).  flag[slot]: bit 0 = in closedSet, bit 1 = in openSet.
static __device__ __noinline__ void bit_update_graph(BitState &S, double goal_id, int s_slot, int lane) {
    for (int i = lane; i < S.n_k; i += 32) S.flag[i] = 0;
    __syncwarp();
    int n_open = 1;
    if (lane == 0) { S.open_[0] = s_slot; S.flag[s_slot] = 2; }
    __syncwarp();
    while (n_open) {
        double bf = CUDART_INF;
        int bi = 0x7fffffff;
        for (int j = lane; j < n_open; j += 32) {
            const double f = S.k_f[S.open_[j]];
            if (bi == 0x7fffffff || f < bf) { bf = f; bi = j; }
        }
        // min(openSet, key=f): first minimum; f >= 0 or +inf, a lane without entries holds (inf, INT_MAX)
        warp_argmin(bf, bi);
        const int cur = S.open_[bi];
        __syncwarp();
        bit_erase_int(S.open_, n_open, bi, lane);
        n_open--;
        if (lane == 0) S.flag[cur] &= ~2;
        __syncwarp();
        if (S.k_id[cur] == goal_id) break;
        if (lane == 0) S.flag[cur] |= 1;
        __syncwarp();
        const double gcur = S.k_g[cur], idcur = S.k_id[cur], xcur = S.k_x[cur], ycur = S.k_y[cur];
        // adjacency of `cur` in add_edge order = its parent (closed: `cur` was reached through it) followed by its children
        for (int suc = S.ch_first[cur]; suc >= 0; suc = S.ch_next[suc]) {
            const int fl = S.flag[suc];
            if (fl & 1) continue;
            const double gs = gcur + bit_norm2(S.k_x[suc] - xcur, S.k_y[suc] - ycur);
            bool set_ = false;
            if (!(fl & 2)) set_ = true;
            else if (!(gs >= S.k_g[suc])) set_ = true;
            const bool first_par = set_ && !S.k_haspar[suc];
            __syncwarp();
            if (lane == 0) {
                if (!(fl & 2)) { S.open_[n_open] = suc; S.flag[suc] = fl | 2; }
                if (set_) {
                    if (first_par) { S.par_order[S.n_par] = suc; S.k_haspar[suc] = 1; }
                    S.k_g[suc] = gs;
                    S.k_f[suc] = gs + S.k_h[suc];
                    S.k_par[suc] = idcur;
                }
            }
            if (!(fl & 2)) n_open++;
            if (first_par) S.n_par++;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(BIT_WARPS_PER_CTA * 32)
bitstar_kernel(rrtk_bitstar_params p, const double *__restrict__ start_goal, const double *__restrict__ rot_all,
               const double4 *__restrict__ obstacles, const int32_t *__restrict__ n_obs_arr, const double *__restrict__ draws_all,
               double *ws_d, int32_t *ws_i, double *path_all, int32_t *counts_all, double *g_goal_out, int32_t *status_out) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * BIT_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (q >= p.n_queries) return;
    const int vcap = p.vertex_cap, kcap = vcap + 2, scap = p.sample_cap, ecap = p.edge_cap;
    BitState S;
    {
        double *d = ws_d + (size_t)q * RRTK_BITSTAR_WS_DOUBLES(vcap, scap, ecap);
        S.s_id = d; d += scap; S.s_x = d; d += scap; S.s_y = d; d += scap;
        S.k_id = d; d += kcap; S.k_g = d; d += kcap; S.k_f = d; d += kcap; S.k_par = d; d += kcap;
        S.eq_x = d; d += ecap;
        S.s_qx = d; d += scap; S.s_qy = d; d += scap; S.s_h = d; d += scap;
        S.k_x = d; d += kcap; S.k_y = d; d += kcap; S.k_h = d; d += kcap;
        S.eq_c = d; d += ecap; S.eq_h = d;
        int *w = ws_i + (size_t)q * RRTK_BITSTAR_WS_INTS(vcap, scap, ecap);
        S.k_haspar = w; w += kcap; S.par_order = w; w += kcap; S.tv = w; w += kcap; S.te_v = w; w += kcap; S.te_x = w; w += kcap;
        S.vq = w; w += kcap; S.open_ = w; w += kcap; S.flag = w; w += kcap; S.eq_v = w; w += ecap;
        S.ch_first = w; w += kcap; S.ch_last = w; w += kcap; S.ch_next = w;
    }
    S.n_s = S.n_k = S.n_par = S.n_v = S.n_te = S.n_vq = S.n_eq = S.n_eq_live = 0;
    const double sx = start_goal[4 * q], sy = start_goal[4 * q + 1], gx = start_goal[4 * q + 2], gy = start_goal[4 * q + 3];
    const double *rot = rot_all + 4 * (size_t)q;
    const double4 *obs = obstacles + (size_t)q * p.obs_stride;
    const int n_obs = n_obs_arr[q];
    const double *draws = draws_all + (size_t)q * p.n_draws;
    BitGrid g;
    g.lower = p.min_rand; g.res = 0.01; g.nc = p.num_cells;
    int status = 0, used = 0;
    for (int i = lane; i < kcap; i += 32) { S.k_haspar[i] = 0; S.ch_first[i] = -1; S.ch_last[i] = -1; S.ch_next[i] = -1; }
    __syncwarp();

    const double start_id = bit_id_of(g, sx, sy), goal_id = bit_id_of(g, gx, gy);
    // setup_planning (:186-216)
    if (lane == 0) {
        S.k_id[0] = goal_id; S.k_g[0] = CUDART_INF; S.k_f[0] = 0.0;
        bit_coord_of(g, goal_id, &S.k_x[0], &S.k_y[0]);
        S.k_h[0] = bit_norm2(S.k_x[0] - S.k_x[0], S.k_y[0] - S.k_y[0]);
    }
    S.n_k = 1;
    __syncwarp();
    bit_samples_set(S, g, scap, goal_id, gx, gy, lane, &status);
    int s_slot = start_id == goal_id ? 0 : 1;
    if (lane == 0) {
        if (s_slot == 1) {
            S.k_id[1] = start_id;
            bit_coord_of(g, start_id, &S.k_x[1], &S.k_y[1]);
            S.k_h[1] = bit_norm2(S.k_x[1] - S.k_x[0], S.k_y[1] - S.k_y[0]);
        }
        S.tv[0] = s_slot;
        S.k_g[s_slot] = 0.0; S.k_f[s_slot] = bit_dist(g, start_id, goal_id);
    }
    S.n_k = s_slot + 1; S.n_v = 1;
    __syncwarp();
    const double c_min = crm_hypot(sx - gx, sy - gy) / 1.5;
    const double xc = (sx + gx) / 2.0, yc = (sy + gy) / 2.0;
    bool ok = bit_informed_sample(S, p, g, 200, CUDART_INF, c_min, xc, yc, rot, p.min_rand, p.max_rand, draws, &used, lane, &status);
    double r = CUDART_INF;
    int iterations = 0, found_goal = 0, n_batches = 0, n_reset = 0, n_skipped = 0, n_expand = 0;
    bool index_error = false, goal_in_tree = s_slot == 0;
    while (ok && !status && iterations < p.max_iter) {
        if (S.n_vq == 0 && S.n_eq_live == 0) {  // setup_sample (:218-234)
            S.n_eq = 0;
            r = 2.0;
            n_batches++;
            if (n_batches >= 2 && iterations == 0) { status |= RRTK_BIT_LIVELOCK; break; }  // the reference never returns
            if (iterations != 0) {
                int m = 100;
                if (found_goal) { m = 200; S.n_s = 0; bit_samples_set(S, g, scap, goal_id, gx, gy, lane, &status); }
                ok = bit_informed_sample(S, p, g, m, S.k_g[0], c_min, xc, yc, rot, p.min_rand, p.max_rand, draws, &used, lane, &status);
                if (!ok || status) break;
            }
            for (int i = lane; i < S.n_v; i += 32) S.vq[i] = S.tv[i];  // the vertex queue is empty here: append all, in order
            S.n_vq = S.n_v;
            __syncwarp();
        }
        for (;;) {  // expand while best_vertex_queue_value() <= best_edge_queue_value() (:244-246)
            double vmin = CUDART_INF;
            int vbest = 0x7fffffff;
            for (int j = lane; j < S.n_vq; j += 32) {
                const int v = S.vq[j];
                const double val = S.k_g[v] + S.k_h[v];
                if (vbest == 0x7fffffff || val < vmin) { vmin = val; vbest = j; }
            }
            warp_argmin(vmin, vbest);
            if (S.n_vq == 0) vmin = CUDART_INF;
            double emax = CUDART_INF;
            if (S.n_eq_live > 0) {
                emax = -CUDART_INF;
                for (int j = lane; j < S.n_eq; j += 32) {
                    if (S.eq_v[j] < 0) continue;
                    const double val = bit_edge_value(S, j);
                    emax = val > emax ? val : emax;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double t = __shfl_xor_sync(FULL, emax, o);
                    emax = t > emax ? t : emax;
                }
            }
            if (!(vmin <= emax)) break;
            if (S.n_vq == 0) { index_error = true; break; }  // best_in_vertex_queue on an empty list raises IndexError
            // expand_vertex (:476-501)
            const int vs = S.vq[vbest];
            const double vid = S.k_id[vs];
            n_expand++;
            __syncwarp();
            bit_erase_int(S.vq, S.n_vq, vbest, lane);
            S.n_vq--;
            const double cx = S.k_x[vs], cy = S.k_y[vs];
            const double d_sv = bit_norm2(cx - S.k_x[s_slot], cy - S.k_y[s_slot]), g_goal = S.k_g[0];
            for (int b0 = 0; b0 < S.n_s; b0 += 32) {
                const int i = b0 + lane;
                bool take = false;
                double sid = 0.0, dvs = 0.0, hs = 0.0;
                if (i < S.n_s) {
                    sid = S.s_id[i];
                    const double dx = S.s_x[i] - cx, dy = S.s_y[i] - cy;
                    if (bit_norm2(dx, dy) <= r && sid != vid && sid != BIT_DEAD_ID) {
                        hs = S.s_h[i];
                        dvs = bit_norm2(S.s_qx[i] - cx, S.s_qy[i] - cy);
                        take = d_sv + hs + dvs < g_goal;
                    }
                }
                const unsigned m = __ballot_sync(FULL, take);
                const int pos = S.n_eq + __popc(m & ((1u << lane) - 1u));
                if (take && pos < ecap) { S.eq_v[pos] = vs; S.eq_x[pos] = sid; S.eq_c[pos] = dvs; S.eq_h[pos] = hs; }
                S.n_eq += __popc(m);
                S.n_eq_live += __popc(m);
            }
            __syncwarp();
            if (S.n_eq > ecap) { status |= RRTK_BIT_EDGE_OVERFLOW; break; }
        }
        if (index_error || status) break;
        // best_in_edge_queue (:469-474): first minimum; then edge_queue.remove(bestEdge)
        double ebv = CUDART_INF;
        int eb = 0x7fffffff;
        for (int j = lane; j < S.n_eq; j += 32) {
            if (S.eq_v[j] < 0) continue;
            const double val = bit_edge_value(S, j);
            if (eb == 0x7fffffff || val < ebv) { ebv = val; eb = j; }
        }
        warp_argmin(ebv, eb);
        const int e0s = S.eq_v[eb];
        const double e0 = S.k_id[e0s], e1 = S.eq_x[eb];
        const double d01 = S.eq_c[eb], h1 = S.eq_h[eb];
        __syncwarp();
        bit_erase_edge(S, eb, lane);
        const double est_v = S.k_g[e0s] + d01 + h1;
        const double est_e = bit_norm2(S.k_x[e0s] - S.k_x[s_slot], S.k_y[e0s] - S.k_y[s_slot]) + d01 + h1;
        const double actual = S.k_g[e0s] + d01, gg = S.k_g[0];
        if (est_v < gg && est_e < gg && actual < gg) {
            double fx, fy, tx, ty;
            bit_coord_of(g, e0, &fx, &fy);
            bit_coord_of(g, e1, &tx, &ty);
            // connect (:359-374): np.linspace(start, end, steps), cut before the first colliding sample
            const double last_edge = bit_id_of(g, tx, ty);
            const long long steps = (long long)(bit_dist(g, bit_id_of(g, fx, fy), last_edge) * 10);
            const double div = (double)(steps - 1), ddx = tx - fx, ddy = ty - fy;
            const double stx = steps > 1 ? ddx / div : 0.0, sty = steps > 1 ? ddy / div : 0.0;
            long long first_hit = steps;
            for (long long b0 = 0; b0 < steps && first_hit == steps; b0 += 32) {
                const long long i = b0 + lane;
                bool hit = false;
                if (i < steps) {
                    double px, py;
                    if (steps > 1 && i == steps - 1) { px = tx; py = ty; }
                    else if (steps > 1) {
                        px = (stx == 0.0 ? ((double)i / div) * ddx : (double)i * stx) + fx;
                        py = (sty == 0.0 ? ((double)i / div) * ddy : (double)i * sty) + fy;
                    } else { px = 0.0 * ddx + fx; py = 0.0 * ddy + fy; }
                    for (int o = 0; o < n_obs && !hit; o++) {
                        const double4 ob = obs[o];
                        const double ex = ob.x - px, ey = ob.y - py;
                        if (ex * ex + ey * ey <= ob.w) hit = true;
                    }
                }
                const unsigned m = __ballot_sync(FULL, hit);
                if (m) first_hit = b0 + __ffs(m) - 1;
            }
            const long long n_free = first_hit;  // points [0, first_hit) are returned; 0 -> None / empty
            if (n_free == 0) { n_skipped++; continue; }  // `continue` before iterations += 1
            double lx, ly;
            {
                const long long i = n_free - 1;
                if (steps > 1 && i == steps - 1) { lx = tx; ly = ty; }
                else if (steps > 1) {
                    lx = (stx == 0.0 ? ((double)i / div) * ddx : (double)i * stx) + fx;
                    ly = (sty == 0.0 ? ((double)i / div) * ddy : (double)i * sty) + fy;
                } else { lx = 0.0 * ddx + fx; ly = 0.0 * ddy + fy; }
            }
            const double nid = bit_id_of(g, lx, ly);
            const int exists = bit_find(S.k_id, S.n_k, nid, lane);
            bool in_tree = false;
            if (exists >= 0) {
                bool f = false;
                for (int i = lane; i < S.n_v; i += 32) f |= S.tv[i] == exists;
                in_tree = __any_sync(FULL, f);
            }
            if (in_tree) { n_skipped++; continue; }
            {   // del self.samples[nid]
                const int f = bit_find(S.s_id, S.n_s, nid, lane);
                if (f >= 0 && lane == 0) { S.s_id[f] = BIT_DEAD_ID; S.s_x[f] = CUDART_INF; }
                __syncwarp();
            }
            if (S.n_v >= vcap) { status |= RRTK_BIT_VERTEX_OVERFLOW; break; }
            int ns = exists;
            if (ns < 0) {
                ns = S.n_k;
                if (lane == 0) {
                    S.k_id[ns] = nid; S.k_haspar[ns] = 0;
                    bit_coord_of(g, nid, &S.k_x[ns], &S.k_y[ns]);
                    S.k_h[ns] = bit_norm2(S.k_x[ns] - S.k_x[0], S.k_y[ns] - S.k_y[0]);
                }
                S.n_k++;
            }
            const double gsc = bit_dist(g, e0, nid);
            if (lane == 0) {
                S.tv[S.n_v] = ns; S.vq[S.n_vq] = ns;
                S.te_v[S.n_te] = e0s; S.te_x[S.n_te] = ns;
                if (S.ch_first[e0s] < 0) S.ch_first[e0s] = ns; else S.ch_next[S.ch_last[e0s]] = ns;
                S.ch_last[e0s] = ns;
                S.k_g[ns] = gsc + S.k_g[e0s];
                S.k_f[ns] = gsc + bit_dist(g, nid, goal_id);
            }
            S.n_v++; S.n_vq++; S.n_te++;
            if (nid == goal_id || e0 == goal_id) found_goal = 1;
            __syncwarp();
            // update_graph (:524-552) relabels the TREE from the start: every vertex is reached through its only tree
            // parent, so g = g[parent] + dist -- the value it already has -- f = g + h and nodes[v] = parent.  Until the goal is a
            // tree vertex nothing stops the pass, it reaches every vertex, and all of them but the new one already hold exactly
            // these values: the pass reduces to labelling the new vertex.  Once the goal is in the tree the pass stops when the
            // goal is popped (vertices behind it keep their stale f / miss their nodes entry), so it is run as written.
            if (ns == 0) goal_in_tree = true;
            if (goal_in_tree) {
                bit_update_graph(S, goal_id, s_slot, lane);
            } else if (lane == 0) {
                S.k_f[ns] = S.k_g[ns] + S.k_h[ns];
                S.k_par[ns] = e0;
                S.k_haspar[ns] = 1;
                S.par_order[S.n_par] = ns;
            }
            if (!goal_in_tree) S.n_par++;
            __syncwarp();
            // remove_queue (:343-351).  Edges of the queue are unique (a vertex is expanded once per queue lifetime), so
            // the loop over the mutating list reduces to: if g[nid] (+ 0) >= g[goal], drop (lastEdge, nid) when present
            if (S.k_g[ns] + bit_dist(g, nid, nid) >= S.k_g[0]) {
                const int le = bit_find(S.k_id, S.n_k, last_edge, lane);
                if (le >= 0) {
                    int pos = -1;
                    for (int b0 = 0; b0 < S.n_eq && pos < 0; b0 += 32) {
                        const int j = b0 + lane;
                        const unsigned m = __ballot_sync(FULL, j < S.n_eq && S.eq_v[j] == le && S.eq_x[j] == nid);
                        if (m) pos = b0 + __ffs(m) - 1;
                    }
                    if (pos >= 0) bit_erase_edge(S, pos, lane);
                }
            }
        } else {  // "Nothing good"
            S.n_eq = 0; S.n_eq_live = 0; S.n_vq = 0;
            n_reset++;
        }
        iterations++;
    }
    if (!ok) status |= RRTK_BIT_DRAWS_EXHAUSTED;
    if (index_error) status |= RRTK_BIT_INDEX_ERROR;
    // find_final_path (:333-341)
    int plen = 0;
    double *path = path_all + (size_t)q * p.path_cap * 2;
    if (!status) {
        if (lane == 0) {
            double cur = goal_id;
            bool found = true;
            if (plen < p.path_cap) { path[0] = gx; path[1] = gy; }
            plen = 1;
            while (cur != start_id) {
                double x, y;
                bit_coord_of(g, cur, &x, &y);
                if (plen < p.path_cap) { path[2 * plen] = x; path[2 * plen + 1] = y; }
                plen++;
                int sl = -1;
                for (int i = 0; i < S.n_k; i++) if (S.k_id[i] == cur) { sl = i; break; }
                if (sl < 0 || !S.k_haspar[sl] || plen > p.path_cap) { found = false; break; }
                cur = S.k_par[sl];
            }
            if (!found) plen = 0;
            else {
                if (plen < p.path_cap) { path[2 * plen] = sx; path[2 * plen + 1] = sy; }
                plen++;
                if (plen <= p.path_cap)
                    for (int i = 0, j = plen - 1; i < j; i++, j--) {
                        const double t0 = path[2 * i], t1 = path[2 * i + 1];
                        path[2 * i] = path[2 * j]; path[2 * i + 1] = path[2 * j + 1];
                        path[2 * j] = t0; path[2 * j + 1] = t1;
                    }
                else { plen = 0; status |= RRTK_BIT_PATH_OVERFLOW; }
            }
        }
        plen = __shfl_sync(FULL, plen, 0);
        status = __shfl_sync(FULL, status, 0);
    }
    {   // compact the edge queue (drop the tombstones, keep the order) so that the workspace holds the list as it is
        int w = 0;
        for (int b0 = 0; b0 < S.n_eq && b0 < ecap; b0 += 32) {
            const int j = b0 + lane;
            const bool live = j < S.n_eq && j < ecap && S.eq_v[j] >= 0;
            int v = 0;
            double x = 0.0;
            if (live) { v = S.eq_v[j]; x = S.eq_x[j]; }
            const unsigned m = __ballot_sync(FULL, live);
            __syncwarp();
            if (live) { const int d = w + __popc(m & ((1u << lane) - 1u)); S.eq_v[d] = v; S.eq_x[d] = x; }
            w += __popc(m);
            __syncwarp();
        }
        S.n_eq = w;
    }
    if (lane == 0) {
        int32_t *c = counts_all + (size_t)q * 12;
        c[0] = S.n_v; c[1] = S.n_te; c[2] = S.n_par; c[3] = S.n_s; c[4] = S.n_vq; c[5] = S.n_eq; c[6] = plen; c[7] = used;
        c[8] = n_batches; c[9] = n_reset; c[10] = n_skipped; c[11] = n_expand;
        g_goal_out[q] = S.k_g[0];
        status_out[q] = status;
    }
}

int launch_bitstar(const rrtk_bitstar_params &p, const double *start_goal, const double *rot, const double *obstacles,
                   const int32_t *n_obs, const double *draws, double *ws_d, int32_t *ws_i, double *path, int32_t *counts,
                   double *g_goal, int32_t *status, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.n_queries + BIT_WARPS_PER_CTA - 1) / BIT_WARPS_PER_CTA);
    bitstar_kernel<<<grid, BIT_WARPS_PER_CTA * 32, 0, s>>>(p, start_goal, rot, reinterpret_cast<const double4 *>(obstacles), n_obs,
                                                           draws, ws_d, ws_i, path, counts, g_goal, status);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "bitstar_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
