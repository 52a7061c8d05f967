// rrtk_rrtstar_dubins.cu -- batched RRT*-Dubins: the `planning()` loop of rrt_05:1416-1456 for Q independent
// queries, one warp per query, persistent grid.  Reference quirks kept:
//   * steer = full Dubins course to the sample, no expand_dis clamp (:1458-1479); near / nearest use xy only;
//   * no play-area checks (:1430-1432); a node is appended only when choose_parent succeeds, and BEFORE rewire;
//   * choose_parent / rewire / propagate costs are EUCLIDEAN (the second calc_new_cost, :1777-1779, wins);
//   * every successful rewire moves the node to the end pose of the new Dubins edge (:1771);
//   * goal test by xy and yaw thresholds, minimum cost, index 0 counts as "not found" (:1445, :1691-1712).
// Each edge (6 words + sampled course + collision) is evaluated by one lane (rrtk_dubins.cuh); the first steer
// of an iteration is uniform, choose_parent and rewire run one lane per near candidate.
// TW = 1: one warp per query (the batch fills the GPU).  TW = 4: one CTA per query, for batches of at most one wave of
// CTAs -- the near lists here hold most of the tree (the radius is ~10 in a 17 x 17 area), so an iteration is
// ceil(count / 32) rounds of edge evaluations per warp, and four warps cut that chain.  Every warp of the team runs the
// same control flow on the same data (nearest, the first edge, the near list: recomputed per warp, no exchange); the
// candidates of choose_parent and of rewire's edge phase are split over the 128 threads; warp 0 alone appends and applies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_dubins.cuh"
#include "rrtk_rs.cuh"

namespace rrtk {

constexpr int DUB_WARPS_PER_CTA = 4;
// Reeds-Shepp candidate rounds of at most this many edges share their 48 words out over the warp (rs_pick_coop: 12 word
// evaluations per group of eight edges); fuller rounds keep one lane per edge
constexpr int RS_COOP_MAX = 16;

__host__ __device__ inline size_t dub_warp_smem_bytes(int near_cap, int node_cap) {
    (void)node_cap;
    // near_idx int | flags int | nd double (d2, then edge cost) | end pose 3 doubles | node cost snapshot double
    // | frontier length of propagate_lists
    size_t b = (size_t)near_cap * (4 + 4 + 8 + 24 + 8) + 16;
    b = (b + 15) & ~(size_t)15;
    return b + ((sizeof(RsWarp) + 15) & ~(size_t)15);   // + the 48 candidate words of a warp-cooperative Reeds-Shepp edge
}

// One steering edge by one lane, for either local planner.  STEER = 0: Dubins (rrt_05, steer returns None when
// len(px) <= 1); STEER = 1: Reeds-Shepp (rrt_06, None when no course, `step` = its step_size).
struct PEdge {
    double ex, ey, eyaw, lsum;
    bool valid, free_;
};
template <int STEER>
__device__ __forceinline__ PEdge plan_edge(double sx, double sy, double syaw, double gx, double gy, double gyaw, double kappa,
                                           double step, const double4 *obs, int n_obs) {
    PEdge r;
    if (STEER == 0) {
        const DubEdge e = dubins_edge_lane(sx, sy, syaw, gx, gy, gyaw, kappa, step, obs, n_obs);
        r.ex = e.ex; r.ey = e.ey; r.eyaw = e.eyaw; r.lsum = 0.0; r.valid = e.npts > 1; r.free_ = e.free_;
    } else {
        const RsEdge e = rs_edge_lane(sx, sy, syaw, gx, gy, gyaw, kappa, step, obs, n_obs);
        r.ex = e.ex; r.ey = e.ey; r.eyaw = e.eyaw; r.lsum = e.lsum; r.valid = e.npts > 0; r.free_ = e.free_;
    }
    return r;
}

// The same edge by the WHOLE warp (uniform arguments and result): used where an iteration has a single edge to evaluate.
template <int STEER>
__device__ __forceinline__ PEdge plan_edge_warp(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                double kappa, double step, const double4 *obs, int n_obs, int lane, RsWarp &W) {
    PEdge r;
    if (STEER == 0) {
        const DubEdge e = dubins_edge_warp(sx, sy, syaw, gx, gy, gyaw, kappa, step, obs, n_obs, lane);
        r.ex = e.ex; r.ey = e.ey; r.eyaw = e.eyaw; r.lsum = 0.0; r.valid = e.npts > 1; r.free_ = e.free_;
    } else {
        const RsEdge e = rs_edge_warp(sx, sy, syaw, gx, gy, gyaw, kappa, step, obs, n_obs, lane, W);
        r.ex = e.ex; r.ey = e.ey; r.eyaw = e.eyaw; r.lsum = e.lsum; r.valid = e.npts > 0; r.free_ = e.free_;
    }
    return r;
}

// propagate_cost_to_leaves with rrt_10's calc_new_cost (:572-577, :1153-1161): cost = parent cost + Reeds-Shepp length of
// the course parent pose -> child pose (inf when there is none).  Breadth-first over the children lists like
// propagate_lists; one lane per frontier node.
static __device__ __noinline__ void propagate_lists_rs(int root, const double2 *xy, const double *yaw, double *cost, int4 *links,
                                                       int *tail, int lane, double kappa, double step) {
    if (links[root].x < 0) return;
    if (lane == 0) { links[0].w = root; *tail = 1; }
    __syncwarp();
    for (int head = 0;;) {
        const int end = *tail;
        if (head >= end) break;
        const int k = head + lane;
        __syncwarp();
        if (k < end) {
            const int p = links[k].w;
            const double2 a = xy[p];
            const double ayaw = yaw[p], cp = cost[p];
            for (int c = links[p].x; c >= 0;) {
                const int4 lc = links[c];
                const double2 b = xy[c];
                const RsEdge e = rs_edge_lane(a.x, a.y, ayaw, b.x, b.y, yaw[c], kappa, step, nullptr, 0, true);
                cost[c] = e.npts > 0 ? cp + e.lsum : CUDART_INF;
                if (lc.x >= 0) links[atomicAdd(tail, 1)].w = c;
                c = lc.y;
            }
        }
        __syncwarp();
        head = end < head + 32 ? end : head + 32;
    }
}

// The course of the first edge of an iteration ends on the sample (to ~1e-13 for coordinates below 1e4), and every course
// point is tested against the circles: a sample inside a circle by more than a band of 1e-9 * (1 + R^2) in squared
// distance makes that edge blocked whatever the course is -- ~20 % of the samples of the built-in scenes, each worth a whole
// edge evaluation (62 % of an iteration's chain goes into that first edge).  Warp-collective, the lanes split the circles.
// Only when the caller does not need to know whether a (blocked) course exists at all.
__device__ __forceinline__ bool sample_certainly_blocked(double fx, double fy, double rx, double ry, const double4 *obs, int n_obs,
                                                         int lane) {
    if (!(fabs(fx) < 1e4 && fabs(fy) < 1e4 && fabs(rx) < 1e4 && fabs(ry) < 1e4)) return false;
    bool in = false;
    for (int o = lane; o < n_obs; o += 32) {
        const double4 ob = obs[o];
        const double dx = ob.x - rx, dy = ob.y - ry;
        if (dx * dx + dy * dy <= ob.w - 1e-9 * (1.0 + ob.w)) in = true;
    }
    return __any_sync(FULL, in);
}

// STEER = 2: rrt_10's RRTStarReedsShepp (rrt_10:1005-1207) -- the STEER = 1 loop with Reeds-Shepp-length costs in
// choose_parent / rewire / propagate (the host passes an unclipped near radius table, rrt_10:521-523).
// STEER = 1 also runs rrt_06's try_goal_path after every append (:1572-1582): the new node is steered to the goal and
// that node is appended too when its course is free, costing the Reeds-Shepp length (:1601).
template <int STEER, int TW>
__global__ void __launch_bounds__(DUB_WARPS_PER_CTA * 32, 3)
rrtstar_dubins_kernel(rrtk_dubins_params p, const double *__restrict__ start_goal6,
                      const double4 *__restrict__ obstacles, const int32_t *__restrict__ n_obs_arr,
                      const double *__restrict__ near_r2, const double *__restrict__ stream3, double2 *xy_all,
                      double *yaw_all, double *cost_all, int32_t *parent_all, double *edge_from_all,
                      double *edge_to_all, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
                      int32_t *status_out, int32_t *workspace, unsigned int *counter) {
    static_assert(TW == 1 || TW == DUB_WARPS_PER_CTA, "a team is one warp or the whole CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double t_red[DUB_WARPS_PER_CTA][4];   // team reduction of choose_parent: cost, end pose of each warp's best
    __shared__ int t_bk[DUB_WARPS_PER_CTA];
    __shared__ unsigned int t_q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lead = TW == 1 || warp == 0;           // the warp that writes the tree
    const int kfirst = TW == 1 ? lane : (int)threadIdx.x, kstride = 32 * TW;   // candidate split
    const int near_cap = p.near_cap;
    unsigned char *base = smem_raw + (size_t)warp * dub_warp_smem_bytes(near_cap, p.node_cap);
    double *nd = reinterpret_cast<double *>(base);
    double *s_end = nd + near_cap;  // [near_cap][3]
    double *s_c = s_end + 3 * near_cap;
    int *near_idx = reinterpret_cast<int *>(s_c + near_cap);
    int *flags = near_idx + near_cap;
    int *qtail = flags + near_cap;
    // rewire's edge phase writes warp 0's arrays (the warp that applies); every warp keeps its own copy of the near list
    double *nd0 = TW == 1 ? nd : reinterpret_cast<double *>(smem_raw);
    double *s_end0 = nd0 + near_cap, *s_c0 = s_end0 + 3 * near_cap;
    int *flags0 = reinterpret_cast<int *>(s_c0 + near_cap) + near_cap;
    auto team_sync = [&]() { if (TW > 1) __syncthreads(); else __syncwarp(); };
    RsWarp &rsw = *reinterpret_cast<RsWarp *>(base + dub_warp_smem_bytes(near_cap, p.node_cap) - ((sizeof(RsWarp) + 15) & ~(size_t)15));
    const double INF = CUDART_INF;
    const double kappa = p.curvature, step = p.step_size;

    for (;;) {
        unsigned int q = 0;
        if (TW > 1) {
            __syncthreads();   // (the previous query's t_q has been read)
            if (threadIdx.x == 0) t_q = atomicAdd(counter, 1u);
            __syncthreads();
            q = t_q;
        } else {
            if (lane == 0) q = atomicAdd(counter, 1u);
            q = __shfl_sync(FULL, q, 0);
        }
        if (q >= (unsigned)p.n_queries) break;
        const double *sg = start_goal6 + 6 * (size_t)q;
        const double sx = sg[0], sy = sg[1], syaw = sg[2], gx = sg[3], gy = sg[4], gyaw = sg[5];
        const double4 *obs = obstacles + (size_t)q * p.obs_stride;
        const int n_obs = n_obs_arr[q];
        double2 *xy = xy_all + (size_t)q * p.node_cap;
        double *yaw = yaw_all + (size_t)q * p.node_cap;
        double *cost = cost_all + (size_t)q * p.node_cap;
        int32_t *parent = parent_all + (size_t)q * p.node_cap;
        double *efrom = edge_from_all + (size_t)q * p.node_cap * 3;
        double *eto = edge_to_all + (size_t)q * p.node_cap * 3;
        int4 *links = reinterpret_cast<int4 *>(workspace + (size_t)q * 4 * p.node_cap);   // children lists (rrtk_planner.cuh)
        const double *stream = stream3 + (size_t)q * p.max_iter * 3;
        if (lead && lane == 0) {
            xy[0] = make_double2(sx, sy); yaw[0] = syaw; cost[0] = 0.0; parent[0] = -1; links[0] = make_int4(-1, -1, -1, 0);
            for (int k = 0; k < 3; k++) { efrom[k] = 0.0; eto[k] = 0.0; }
        }
        int n = 1, status = RRTK_Q_OK, gi = -1, it = 0;
        bool done = false;

        // search_best_goal_node (rrt_05:1691-1712): uniform result, -1 = none
        auto best_goal = [&]() {
            double bc = INF;
            int bi = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                double2 a = xy[i];
                if (crm_hypot(a.x - gx, a.y - gy) <= p.goal_xy_th && fabs(yaw[i] - gyaw) <= p.goal_yaw_th) {
                    double c = cost[i];
                    if (c < bc) { bc = c; bi = i; }
                }
            }
            warp_argmin(bc, bi);
            return bi == 0x7fffffff ? -1 : bi;
        };

        for (it = 0; it < p.max_iter; it++) {
            team_sync();   // what the lead warp wrote in the previous iteration is visible; t_red may be reused
            const double rx = stream[3 * it], ry = stream[3 * it + 1], ryaw = stream[3 * it + 2];
            // nearest on xy (rrt_05:1605-1610)
            double bd = INF;
            int bi = 0x7fffffff;
#pragma unroll 1
            for (int i = lane; i < n; i += 32) {
                double2 a = xy[i];
                double ddx = a.x - rx, ddy = a.y - ry;
                double d = ddx * ddx + ddy * ddy;
                if (d < bd) { bd = d; bi = i; }
            }
            warp_argmin(bd, bi);
            const int ni = bi;
            const double2 from = xy[ni];
            const double fyaw = yaw[ni];
            // (with search_until_max_iter a blocked first edge ends the iteration: `truthy` is only read together with free_)
            if (p.search_until_max_iter && sample_certainly_blocked(from.x, from.y, rx, ry, obs, n_obs, lane)) continue;
            const PEdge e0 = plan_edge_warp<STEER>(from.x, from.y, fyaw, rx, ry, ryaw, kappa, step, obs, n_obs, lane, rsw);
            bool truthy = e0.valid;
            if (truthy && e0.free_) {
                const double nx = e0.ex, ny = e0.ey, nyaw = e0.eyaw;
                truthy = false;
                if (n + (STEER >= 1 ? 1 : 0) >= p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; break; }
                // find_near_nodes (rrt_05:1715-1739)
                const double r2 = near_r2[n + 1];
                int count = 0;
#pragma unroll 1
                for (int b0 = 0; b0 < n; b0 += 32) {
                    int i = b0 + lane;
                    bool hit = false;
                    double d = 0.0;
                    if (i < n) {
                        double2 a = xy[i];
                        double ddx = a.x - nx, ddy = a.y - ny;
                        d = ddx * ddx + ddy * ddy;
                        hit = d <= r2;
                    }
                    unsigned mask = __ballot_sync(FULL, hit);
                    int pos = count + __popc(mask & ((1u << lane) - 1u));
                    if (hit && pos < near_cap) { near_idx[pos] = i; nd[pos] = d; }
                    count += __popc(mask);
                }
                __syncwarp();
                if (count > near_cap) { status |= RRTK_Q_NEAR_OVERFLOW; break; }
                for (int k = lane; k < count; k += 32) {  // `.index()` first-occurrence mapping
                    double dk = nd[k];
                    int f = k;
                    for (int j = 0; j < k; j++)
                        if (nd[j] == dk) { f = j; break; }
                    flags[k] = near_idx[f];
                }
                __syncwarp();
                for (int k = lane; k < count; k += 32) near_idx[k] = flags[k];
                __syncwarp();
                // choose_parent (rrt_05:1648-1689): one lane per candidate
                double mc = INF, bex = 0.0, bey = 0.0, beyaw = 0.0;
                int bk = 0x7fffffff;
#pragma unroll 1
                for (int k0 = kfirst - lane; k0 < count; k0 += kstride) {   // (uniform trip count: see RS_COOP_MAX)
                    const int k = k0 + lane;
                    const bool in = k < count;
                    const int i = in ? near_idx[k] : 0;
                    const double2 a = xy[i];
                    const int nact = count - k0 < 32 ? count - k0 : 32;
                    PEdge e;
                    e.valid = false; e.free_ = false; e.ex = e.ey = e.eyaw = e.lsum = 0.0;
                    if (STEER >= 1 && nact <= RS_COOP_MAX) {   // few candidates: their words shared out over the warp
                        const RsEdge re = rs_edges_coop(nact, lane, a.x, a.y, yaw[i], nx, ny, nyaw, kappa, step, obs, n_obs);
                        e.ex = re.ex; e.ey = re.ey; e.eyaw = re.eyaw; e.lsum = re.lsum; e.valid = re.npts > 0; e.free_ = re.free_;
                    } else if (in) {
                        e = plan_edge<STEER>(a.x, a.y, yaw[i], nx, ny, nyaw, kappa, step, obs, n_obs);
                    }
                    if (in && e.valid && e.free_) {
                        const double c = cost[i] + (STEER == 2 ? e.lsum : crm_hypot(nx - a.x, ny - a.y));
                        if (c < mc) { mc = c; bk = k; bex = e.ex; bey = e.ey; beyaw = e.eyaw; }
                    }
                }
                warp_argmin(mc, bk);
                double cx = 0.0, cy = 0.0, cyaw = 0.0;
                if (bk != 0x7fffffff) {   // (candidate k sits on lane k & 31 of its warp)
                    const int src = bk & 31;
                    cx = __shfl_sync(FULL, bex, src); cy = __shfl_sync(FULL, bey, src); cyaw = __shfl_sync(FULL, beyaw, src);
                }
                if (TW > 1) {   // the first minimum over the team's warps
                    if (lane == 0) { t_red[warp][0] = mc; t_red[warp][1] = cx; t_red[warp][2] = cy; t_red[warp][3] = cyaw; t_bk[warp] = bk; }
                    __syncthreads();
                    mc = t_red[0][0]; bk = t_bk[0];
                    int bw = 0;
                    for (int w = 1; w < TW; w++) {
                        const double cw = t_red[w][0];
                        const int kw = t_bk[w];
                        if (cw < mc || (cw == mc && kw < bk)) { mc = cw; bk = kw; bw = w; }
                    }
                    cx = t_red[bw][1]; cy = t_red[bw][2]; cyaw = t_red[bw][3];
                }
                if (bk != 0x7fffffff) {
                    const int best = near_idx[bk];
                    const int newi = n;
                    __syncwarp();
                    if (lead && lane == 0) {
                        const double2 b = xy[best];
                        efrom[3 * newi] = b.x; efrom[3 * newi + 1] = b.y; efrom[3 * newi + 2] = yaw[best];
                        eto[3 * newi] = nx; eto[3 * newi + 1] = ny; eto[3 * newi + 2] = nyaw;
                        xy[newi] = make_double2(cx, cy); yaw[newi] = cyaw; cost[newi] = mc; parent[newi] = best;
                        links[newi] = make_int4(-1, -1, -1, 0);
                        link_child(links, best, newi);
                    }
                    n++;
                    truthy = true;
                    __syncwarp();
                    // rewire (rrt_05:1741-1775), phase A: the entries' edges in parallel.  With Euclidean costs (STEER 0 / 1) the
                    // cost an entry would get does not depend on its edge, and planning an edge has no side effect: an entry
                    // whose node does not cost more than that NOW is not planned (flag 8); phase B looks at it again with the
                    // node's cost at its turn (a re-parented node moves, so costs can go either way).
#pragma unroll 1
                    for (int k0 = kfirst - lane; k0 < count; k0 += kstride) {
                        const int k = k0 + lane;
                        const bool in = k < count;
                        const int i = in ? near_idx[k] : 0;
                        const double2 a = xy[i];
                        const double ayaw = yaw[i];
                        const double ci = cost[i];
                        double ec = STEER == 2 ? 0.0 : mc + crm_hypot(a.x - cx, a.y - cy);
                        int fl = 8;
                        double ex = 0.0, ey = 0.0, eyw = 0.0;
                        const bool need = in && (STEER == 2 || ci > ec);
                        const unsigned nm = __ballot_sync(FULL, need);
                        const int nact = __popc(nm);
                        if (STEER >= 1 && nact > 0 && nact <= RS_COOP_MAX) {
                            // the entries to plan move to lanes 0..nact-1 (rs_edges_coop shares their words out)
                            const int srcl = lane < nact ? (int)__fns(nm, 0, lane + 1) : 0;
                            const double tx = __shfl_sync(FULL, a.x, srcl), ty = __shfl_sync(FULL, a.y, srcl);
                            const double tyaw = __shfl_sync(FULL, ayaw, srcl);
                            const RsEdge re = rs_edges_coop(nact, lane, cx, cy, cyaw, tx, ty, tyaw, kappa, step, obs, n_obs);
                            // ... and their results go back to the lanes that hold the entries
                            const int rank = __popc(nm & ((1u << lane) - 1u));
                            const int rfl = (re.npts > 0 ? 1 : 0) | (re.free_ ? 2 : 0);
                            const int bfl = __shfl_sync(FULL, rfl, rank);
                            const double bex_ = __shfl_sync(FULL, re.ex, rank), bey_ = __shfl_sync(FULL, re.ey, rank);
                            const double beyw_ = __shfl_sync(FULL, re.eyaw, rank), bls_ = __shfl_sync(FULL, re.lsum, rank);
                            if (need) {
                                fl = bfl; ex = bex_; ey = bey_; eyw = beyw_;
                                if (STEER == 2) ec = mc + bls_;
                            }
                        } else if (need) {
                            const PEdge e = plan_edge<STEER>(cx, cy, cyaw, a.x, a.y, ayaw, kappa, step, obs, n_obs);
                            fl = (e.valid ? 1 : 0) | (e.free_ ? 2 : 0);
                            if (STEER == 2) ec = mc + e.lsum;
                            ex = e.ex; ey = e.ey; eyw = e.eyaw;
                        }
                        if (in) {
                            flags0[k] = fl;
                            nd0[k] = ec;
                            s_c0[k] = ci;
                            s_end0[3 * k] = ex; s_end0[3 * k + 1] = ey; s_end0[3 * k + 2] = eyw;
                        }
                    }
                    team_sync();
                    // phase B: apply in list order.  s_c is the current cost of every entry's node (refreshed after each
                    // re-parenting), so the lanes can find the next entry that can act without walking the list one by one.
                    for (int k0 = 0; lead && k0 < count;) {
                        bool need = false;
                        if (k0 + lane < count) {
                            const int fl = flags[k0 + lane];
                            need = (fl & 4) || (((fl & 3) == 3 || (fl & 8)) && s_c[k0 + lane] > nd[k0 + lane]);
                        }
                        const unsigned nm = __ballot_sync(FULL, need);
                        if (!nm) { k0 += 32; continue; }
                        const int k = k0 + __ffs(nm) - 1;
                        k0 = k + 1;
                        const int i = near_idx[k];
                        int fl = flags[k];
                        double ecost = nd[k], ex = s_end[3 * k], ey = s_end[3 * k + 1], eyw = s_end[3 * k + 2];
                        const double2 a = xy[i];
                        const double ayaw = yaw[i];
                        if (fl & (4 | 8)) {  // node i was re-parented (moved) earlier in this call, or its edge was not planned yet
                            const PEdge e = plan_edge_warp<STEER>(cx, cy, cyaw, a.x, a.y, ayaw, kappa, step, obs, n_obs, lane, rsw);
                            fl = (e.valid ? 1 : 0) | (e.free_ ? 2 : 0);
                            ecost = mc + (STEER == 2 ? e.lsum : crm_hypot(a.x - cx, a.y - cy));
                            ex = e.ex; ey = e.ey; eyw = e.eyaw;
                        }
                        if ((fl & 3) != 3) continue;
                        if (cost[i] > ecost) {
                            __syncwarp();
                            if (lane == 0) {
                                unlink_child(links, parent[i], i);
                                link_child(links, newi, i);
                                efrom[3 * i] = cx; efrom[3 * i + 1] = cy; efrom[3 * i + 2] = cyaw;
                                eto[3 * i] = a.x; eto[3 * i + 1] = a.y; eto[3 * i + 2] = ayaw;
                                xy[i] = make_double2(ex, ey); yaw[i] = eyw; cost[i] = ecost; parent[i] = newi;
                            }
                            __syncwarp();
                            for (int k2 = k + 1 + lane; k2 < count; k2 += 32)
                                if (near_idx[k2] == i) flags[k2] |= 4;
                            __syncwarp();
                            if (STEER == 2) propagate_lists_rs(i, xy, yaw, cost, links, qtail, lane, kappa, step);
                            else propagate_lists(i, xy, cost, links, qtail, lane);
                            __syncwarp();
                            for (int k2 = k + 1 + lane; k2 < count; k2 += 32) s_c[k2] = cost[near_idx[k2]];
                            __syncwarp();
                        }
                    }
                    if (STEER >= 1) {   // try_goal_path (rrt_06:1572-1582), from the new node as it is now
                        team_sync();
                        const double2 a = xy[newi];
                        const double ayaw = yaw[newi];
                        const PEdge eg = plan_edge_warp<STEER>(a.x, a.y, ayaw, gx, gy, gyaw, kappa, step, obs, n_obs, lane, rsw);
                        if (eg.valid && eg.free_) {
                            if (lead && lane == 0) {
                                efrom[3 * n] = a.x; efrom[3 * n + 1] = a.y; efrom[3 * n + 2] = ayaw;
                                eto[3 * n] = gx; eto[3 * n + 1] = gy; eto[3 * n + 2] = gyaw;
                                xy[n] = make_double2(eg.ex, eg.ey); yaw[n] = eg.eyaw; cost[n] = cost[newi] + eg.lsum; parent[n] = newi;
                                links[n] = make_int4(-1, -1, -1, 0);
                                link_child(links, newi, n);
                            }
                            n++;
                            __syncwarp();
                        }
                    }
                }
            }
            if (!p.search_until_max_iter && truthy) {
                team_sync();
                gi = best_goal();
                if (gi > 0) { it++; done = true; break; }
            }
        }
        team_sync();
        if (!done) gi = best_goal();
        if (gi <= 0) gi = -1;  // `if last_index:` -- index 0 is falsy (rrt_05:1445, :1451)
        if (lead && lane == 0) {
            n_nodes[q] = n;
            iters_done[q] = it;
            goal_index[q] = gi;
            status_out[q] = status;
        }
        __syncwarp();
    }
}

// `sum([a, b, c])` of three Python floats as CPython >= 3.12 evaluates it: the first item plainly, the others through
// Neumaier's compensated addition, the compensation added at the end (bltinmodule.c builtin_sum_impl; rrt_03:1470)
static __device__ __forceinline__ double py312_sum3(double a, double b, double c) {
    double f = a, comp = 0.0, t;
    t = f + b; comp += fabs(f) >= fabs(b) ? (f - t) + b : (b - t) + f; f = t;
    t = f + c; comp += fabs(f) >= fabs(c) ? (f - t) + c : (c - t) + f; f = t;
    if (comp != 0.0 && isfinite(comp)) f += comp;
    return f;
}

// RRT-Dubins (rrt_03:1402-1456), one warp per query: nearest on xy, ONE Dubins edge per iteration evaluated by the whole
// warp (words across lanes, course points across lanes), play-area test of the end pose, append.
__global__ void __launch_bounds__(DUB_WARPS_PER_CTA * 32, 3)
rrt_dubins_kernel(rrtk_dubins_params p, const double *__restrict__ start_goal6, const double4 *__restrict__ obstacles,
                  const int32_t *__restrict__ n_obs_arr, const double *__restrict__ play, const double *__restrict__ stream3,
                  double2 *xy_all, double *yaw_all, double *cost_all, int32_t *parent_all, double *edge_from_all,
                  double *edge_to_all, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status_out,
                  unsigned int *counter) {
    const int lane = threadIdx.x & 31;
    const double INF = CUDART_INF;
    const double kappa = p.curvature, step = p.step_size;
    double pa[4] = {0.0, 0.0, 0.0, 0.0};
    if (play) { pa[0] = play[0]; pa[1] = play[1]; pa[2] = play[2]; pa[3] = play[3]; }
    for (;;) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(counter, 1u);
        q = __shfl_sync(FULL, q, 0);
        if (q >= (unsigned)p.n_queries) break;
        const double *sg = start_goal6 + 6 * (size_t)q;
        const double gx = sg[3], gy = sg[4], gyaw = sg[5];
        const double4 *obs = obstacles + (size_t)q * p.obs_stride;
        const int n_obs = n_obs_arr[q];
        double2 *xy = xy_all + (size_t)q * p.node_cap;
        double *yaw = yaw_all + (size_t)q * p.node_cap;
        double *cost = cost_all + (size_t)q * p.node_cap;
        int32_t *parent = parent_all + (size_t)q * p.node_cap;
        double *efrom = edge_from_all + (size_t)q * p.node_cap * 3;
        double *eto = edge_to_all + (size_t)q * p.node_cap * 3;
        const double *stream = stream3 + (size_t)q * p.max_iter * 3;
        if (lane == 0) {
            xy[0] = make_double2(sg[0], sg[1]); yaw[0] = sg[2]; cost[0] = 0.0; parent[0] = -1;
            for (int k = 0; k < 3; k++) { efrom[k] = 0.0; eto[k] = 0.0; }
        }
        __syncwarp();
        int n = 1, status = RRTK_Q_OK, gi = -1, it = 0;
        bool done = false;
        auto best_goal = [&]() {   // search_best_goal_node (rrt_03:1491-1512): first minimum of the cost, -1 = none
            double bc = INF;
            int bi = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                const double2 a = xy[i];
                if (crm_hypot(a.x - gx, a.y - gy) <= p.goal_xy_th && fabs(yaw[i] - gyaw) <= p.goal_yaw_th) {
                    const double c = cost[i];
                    if (c < bc) { bc = c; bi = i; }
                }
            }
            warp_argmin(bc, bi);
            return bi == 0x7fffffff ? -1 : bi;
        };
        for (it = 0; it < p.max_iter; it++) {
            const double rx = stream[3 * it], ry = stream[3 * it + 1], ryaw = stream[3 * it + 2];
            double bd = INF;
            int bi = 0x7fffffff;
#pragma unroll 1
            for (int i = lane; i < n; i += 32) {   // get_nearest_node_index (rrt_03:1612-1618)
                const double2 a = xy[i];
                const double ddx = a.x - rx, ddy = a.y - ry;
                const double d = ddx * ddx + ddy * ddy;
                if (d < bd) { bd = d; bi = i; }
            }
            warp_argmin(bd, bi);
            const int ni = bi;
            const double2 from = xy[ni];
            const double fyaw = yaw[ni];
            double len3[3] = {0.0, 0.0, 0.0};
            // (`truthy` of a blocked edge is read only for the play-area abort and the early goal test)
            if (!play && p.search_until_max_iter && sample_certainly_blocked(from.x, from.y, rx, ry, obs, n_obs, lane)) continue;
            const DubEdge e = dubins_edge_warp(from.x, from.y, fyaw, rx, ry, ryaw, kappa, step, obs, n_obs, lane, len3);
            const bool truthy = e.npts > 1;
            if (!truthy && play) { status |= RRTK_Q_NONE_STEER; it++; done = true; break; }
            if (truthy && e.free_ && (!play || !(e.ex < pa[0] || e.ex > pa[1] || e.ey < pa[2] || e.ey > pa[3]))) {
                if (n >= p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; it++; done = true; break; }
                if (lane == 0) {
                    efrom[3 * n] = from.x; efrom[3 * n + 1] = from.y; efrom[3 * n + 2] = fyaw;
                    eto[3 * n] = rx; eto[3 * n + 1] = ry; eto[3 * n + 2] = ryaw;
                    xy[n] = make_double2(e.ex, e.ey); yaw[n] = e.eyaw; parent[n] = ni;
                    cost[n] = cost[ni] + py312_sum3(fabs(len3[0]), fabs(len3[1]), fabs(len3[2]));
                }
                n++;
                __syncwarp();
            }
            if (!p.search_until_max_iter && truthy) {
                gi = best_goal();
                if (gi > 0) { it++; done = true; break; }
            }
        }
        if (!done) gi = best_goal();
        if (gi <= 0) gi = -1;   // `if last_index:` -- index 0 is falsy (rrt_03:1446, :1452)
        if (lane == 0) {
            n_nodes[q] = n;
            iters_done[q] = it;
            goal_index[q] = gi;
            status_out[q] = status;
        }
        __syncwarp();
    }
}

int launch_rrt_dubins(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles, const int32_t *n_obs,
                      const double *play, const double *stream3, double *xy, double *yaw, double *cost, int32_t *parent,
                      double *edge_from, double *edge_to, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
                      int32_t *status, unsigned int *counter, cudaStream_t s) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rrt_dubins_kernel, DUB_WARPS_PER_CTA * 32, 0);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) per_sm = 1;
    long long want = ((long long)p.n_queries + DUB_WARPS_PER_CTA - 1) / DUB_WARPS_PER_CTA;
    long long grid = (long long)sms * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counter)");
    rrt_dubins_kernel<<<(unsigned)grid, DUB_WARPS_PER_CTA * 32, 0, s>>>(
        p, start_goal6, reinterpret_cast<const double4 *>(obstacles), n_obs, play, stream3, reinterpret_cast<double2 *>(xy), yaw,
        cost, parent, edge_from, edge_to, n_nodes, iters_done, goal_index, status, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rrt_dubins_kernel launch");
    return RRTK_OK;
}

int launch_rrtstar_steer(int steer, const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles,
                          const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                          double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                          int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                          int32_t *workspace, unsigned int *counter, cudaStream_t s) {
    size_t smem = dub_warp_smem_bytes(p.near_cap, p.node_cap) * DUB_WARPS_PER_CTA;
    if (smem > 227 * 1024) return set_error(RRTK_ERR_INVALID, "near_cap/node_cap need more than 227 KB of shared memory");
    typedef void (*kernel_t)(rrtk_dubins_params, const double *, const double4 *, const int32_t *, const double *, const double *,
                             double2 *, double *, double *, int32_t *, double *, double *, int32_t *, int32_t *, int32_t *,
                             int32_t *, int32_t *, unsigned int *);
    const kernel_t kern_w = steer == 2 ? rrtstar_dubins_kernel<2, 1> : steer == 1 ? rrtstar_dubins_kernel<1, 1> : rrtstar_dubins_kernel<0, 1>;
    const kernel_t kern_c = steer == 2 ? rrtstar_dubins_kernel<2, DUB_WARPS_PER_CTA>
                          : steer == 1 ? rrtstar_dubins_kernel<1, DUB_WARPS_PER_CTA> : rrtstar_dubins_kernel<0, DUB_WARPS_PER_CTA>;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // execution: a CTA per query while the batch is at most two CTAs per SM (the chain of a query is what a launch of
    // that size takes), a warp per query beyond
    bool cta = p.exec_mode == RRTK_EXEC_CTA;
    if (p.exec_mode != RRTK_EXEC_CTA && p.exec_mode != RRTK_EXEC_WARP) {
        cudaError_t e0 = cudaFuncSetAttribute(kern_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e0 != cudaSuccess) return set_cuda_error(e0, "cudaFuncSetAttribute(rrtstar_dubins_kernel)");
        int per_c = 0;
        e0 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_c, kern_c, DUB_WARPS_PER_CTA * 32, smem);
        if (e0 != cudaSuccess) return set_cuda_error(e0, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        if (per_c > 2) per_c = 2;   // (measured: at three CTAs per SM the teams contend and a warp per query is faster)
        cta = (long long)p.n_queries <= (long long)sms * (per_c < 1 ? 1 : per_c);
    }
    const kernel_t kern = cta ? kern_c : kern_w;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(rrtstar_dubins_kernel)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DUB_WARPS_PER_CTA * 32, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) per_sm = 1;
    long long want = cta ? (long long)p.n_queries : ((long long)p.n_queries + DUB_WARPS_PER_CTA - 1) / DUB_WARPS_PER_CTA;
    long long grid = (long long)sms * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counter)");
    kern<<<(unsigned)grid, DUB_WARPS_PER_CTA * 32, smem, s>>>(
        p, start_goal6, reinterpret_cast<const double4 *>(obstacles), n_obs, near_r2, stream3,
        reinterpret_cast<double2 *>(xy), yaw, cost, parent, edge_from, edge_to, n_nodes, iters_done, goal_index,
        status, workspace, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rrtstar_dubins_kernel launch");
    return RRTK_OK;
}

int launch_rrtstar_dubins(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles,
                          const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                          double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                          int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                          int32_t *workspace, unsigned int *counter, cudaStream_t s) {
    return launch_rrtstar_steer(0, p, start_goal6, obstacles, n_obs, near_r2, stream3, xy, yaw, cost, parent, edge_from, edge_to,
                                n_nodes, iters_done, goal_index, status, workspace, counter, s);
}
int launch_rrtstar_rs(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles,
                      const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                      double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                      int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                      int32_t *workspace, unsigned int *counter, cudaStream_t s) {
    return launch_rrtstar_steer(p.rs_cost ? 2 : 1, p, start_goal6, obstacles, n_obs, near_r2, stream3, xy, yaw, cost, parent, edge_from, edge_to,
                                n_nodes, iters_done, goal_index, status, workspace, counter, s);
}

}  // namespace rrtk
