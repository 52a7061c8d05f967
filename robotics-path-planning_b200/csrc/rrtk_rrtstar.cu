// rrtk_rrtstar.cu -- batched RRT / RRT* planning loop, one warp per query, persistent grid.
//
// Replaces the whole `planning()` loop of the reference's RRT* (rrt_04:1036-1084) and basic RRT
// (rrt_01:71-101) for Q independent queries.  Design:
//   * one warp owns one query for all its iterations; warps pull query ids from an atomic counter
//     (persistent grid sized to the SM count), so ragged per-query work balances itself;
//   * the tree lives in the caller's output arrays (xy as double2, cost, parent) -- L2-resident for
//     the active set -- and is scanned with coalesced 16-byte loads (nearest, near, propagate);
//   * all decisions are taken in FP64 with the reference's operation order and correctly rounded
//     hypot/atan2/cos/sin (crmath.h); compiled with -fmad=false so no multiply-add is contracted;
//   * per iteration the obstacles are culled once (exact, conservative disc test around the new
//     node) into shared memory; every edge of that iteration (1 + 2|near| steers) is tested only
//     against the survivors -- collision verdicts are unchanged because a culled obstacle cannot
//     touch any path point of that iteration;
//   * choose_parent / rewire: one lane per near candidate (steer + collision + cost in parallel),
//     warp-shuffle first-min, then the order-dependent apply phase of rewire runs in list order with
//     level-synchronous cost propagation over the parent array.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
#include "rrtk_rrtstar_common.cuh"

namespace rrtk {

constexpr int WARPS_PER_CTA = 4;
// rare branches (exact fall-backs, overflows): told to the compiler so that their code leaves the hot straight-line path
// (the kernel is bound by instruction fetch; the L1.5 instruction cache holds 32 KB)
#define RRTK_RARE(c) __builtin_expect(!!(c), 0)
#define RRTK_PRAGMA(x) _Pragma(#x)
#define RRTK_UNROLL(n) RRTK_PRAGMA(unroll n)
#ifndef RRTK_UNROLL_NEAREST
#define RRTK_UNROLL_NEAREST 1
#endif
#ifndef RRTK_UNROLL_NEAR
#define RRTK_UNROLL_NEAR 1
#endif
#ifndef RRTK_UNROLL_PROP
#define RRTK_UNROLL_PROP 1
#endif

// find_near_nodes (rrt_04:1314-1338) around (cx, cy): ballot compaction in ascending index order.  Only used when the
// new node is not the sample (the merged scan of the main loop covers the common case).
__device__ __noinline__ int near_scan(const double2 *xy, int n, double cx, double cy, double r2, int *near_idx, double *nd,
                                     int near_cap, int lane) {
    int count = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
        int i = b0 + lane;
        bool hit = false;
        double d = 0.0;
        if (i < n) {
            double2 a = xy[i];
            double ddx = a.x - cx, ddy = a.y - cy;
            d = ddx * ddx + ddy * ddy;
            hit = d <= r2;
        }
        unsigned mask = __ballot_sync(FULL, hit);
        int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < near_cap) { near_idx[pos] = i; nd[pos] = d; }
        count += __popc(mask);
    }
    return count;
}

// rewire entries [from, count) one at a time against the current tree (after a re-parented node MOVED, rrt_04:1365-1371:
// the parallel pass's view of positions and costs is stale).  Out of line: rare.
__device__ __noinline__ void rewire_serial(const PlanConsts p, int from, int count, const int *near_idx, double2 *xy,
                                          double *cost, int32_t *parent, int4 *links, double *elen, int n, double cx, double cy,
                                          double ccost, const ObsList &L, int *qtail, int lane, int &t_rwok, int &t_rwap) {
    const double res = p.res;
    for (int k = from; k < count; k++) {
        const int i = near_idx[k];
        const double2 a = xy[i];
        Steer st = steer(cx, cy, a.x, a.y, CUDART_INF, res);
        const bool ok = edge_free_warp(cx, cy, st, a.x, a.y, L, lane) && inside_play(p, st.ex, st.ey);
        const double ec = ccost + st.d;
        t_rwok += ok ? 1 : 0;
        if (ok && cost[i] > ec) {
            __syncwarp();
            if (lane == 0) {
                unlink_child(links, parent[i], i);
                link_child(links, n, i);
                xy[i] = make_double2(st.ex, st.ey);
                cost[i] = ec;
                parent[i] = n;
                elen[i] = __longlong_as_double(0x7ff8000000000000ll);   // (the node may have moved: recomputed on demand)
            }
            __syncwarp();
            t_rwap++;
#ifdef RRTK_V_NOELEN
            propagate_lists(i, xy, cost, links, qtail, lane);
#else
            propagate_lists_elen(i, true, xy, cost, elen, links, qtail, lane);
#endif
        }
    }
}

// per-warp shared memory
struct WarpSmem {
    // obstacles near the new node of the current iteration (re-used for the goal by best_goal)
    double cull_x[CULL_CAP], cull_y[CULL_CAP], cull_r2[CULL_CAP];
    int qtail;      // frontier length of propagate_lists
    // per-query values used once or twice per iteration: read here at the point of use instead of being carried in registers
    // through the whole loop (the kernel sits at the 128-register cap and was spilling ~250 B per thread)
    unsigned int sob_q0, sob_q1;
    long long sob_n;
    const double2 *q_stream;
    int32_t *q_gcnt;
    uint16_t *q_glists;
    double q_ginv;
    double *q_elen;
};

// layout of the dynamic shared memory of one warp:
//   WarpSmem | nd[near_cap] double (d2, then edge length) | nc[near_cap] double (node cost)
//   | near_idx[near_cap] int | near_ok[near_cap] int
__host__ __device__ inline size_t warp_smem_bytes(int near_cap, int node_cap) {
    (void)node_cap;
    size_t b = sizeof(WarpSmem) + (size_t)near_cap * (4 + 4 + 8 * 2);
    return (b + 15) & ~(size_t)15;
}

#ifndef RRTK_MIN_BLOCKS
#define RRTK_MIN_BLOCKS 4
#endif
// RRT_ONLY / TRACE are compile-time so the planning launch carries neither the basic-RRT branches nor the trace
// bookkeeping in its instruction stream (the loop is instruction-fetch sensitive)
// NC > 0: the per-warp shared-memory arrays are laid out for a near list of NC entries (p.near_cap <= NC stays the
// logical capacity), so every shared-memory address is an immediate; NC = 0: laid out for p.near_cap at run time.
// RESUME: continue the trees of a previous call (p.resume); a separate instantiation so that the default path keeps its
// register allocation (the kernel sits at the 128-register cap: the few extra live values cost 4 % when they were a
// run-time branch).
template <bool RRT_ONLY, bool TRACE, int NC, bool RESUME = false>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, RRTK_MIN_BLOCKS)
rrtstar_kernel(rrtk_rrtstar_params p, const double4 *__restrict__ start_goal,
               const double4 *__restrict__ obstacles, const int32_t *__restrict__ n_obs_arr,
               const double *__restrict__ near_r2, const double2 *__restrict__ sample_stream,
               const int64_t *__restrict__ sobol_offset, double2 *xy_all, double *cost_all,
               int32_t *parent_all, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
               int32_t *status_out, int32_t *trace_all, int32_t *workspace, unsigned int *counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int near_cap = p.near_cap;
    const int lay = NC > 0 ? NC : near_cap;   // layout capacity
    unsigned char *base = smem_raw + (size_t)wib * warp_smem_bytes(lay, p.node_cap);
    WarpSmem *ws = reinterpret_cast<WarpSmem *>(base);
    double *nd = reinterpret_cast<double *>(base + sizeof(WarpSmem));
    double *s_nc = nd + lay;
    int *near_idx = reinterpret_cast<int *>(s_nc + lay);
    int *near_ok = near_idx + lay;
    const double res = p.path_resolution;
    const double INF = CUDART_INF;

    for (;;) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(counter, 1u);
        q = __shfl_sync(FULL, q, 0);
        if (q >= (unsigned)p.n_queries) break;

        const double4 sg = start_goal[q];
        const double gx = sg.z, gy = sg.w;
        const double4 *obs = obstacles + (size_t)q * p.obs_stride;
        const int n_obs = n_obs_arr[q];
        double2 *xy = xy_all + (size_t)q * p.node_cap;
        double *cost = cost_all + (size_t)q * p.node_cap;
        int32_t *parent = parent_all + (size_t)q * p.node_cap;
        const double2 *stream = sample_stream ? sample_stream + (size_t)q * p.max_iter : nullptr;
        int32_t *trace = TRACE ? trace_all + (size_t)q * p.max_iter * 8 : nullptr;
        const int64_t sobol_base = sobol_offset ? sobol_offset[q] : 0;

        // scratch of the query: children lists + the propagate frontier (4 * node_cap ints), the cached edge lengths
        // elen[i] = hypot(node i - its parent) (node_cap doubles; NaN = not known, recomputed on demand), obstacle cells
        const int grid_cells = p.grid_nx * p.grid_ny;
        int4 *links = reinterpret_cast<int4 *>(workspace + (size_t)q * RRTK_RRTSTAR_WS_INTS(p.node_cap, p.grid_nx, p.grid_ny));
        double *elen = reinterpret_cast<double *>(links + p.node_cap);
        const double NaN = __longlong_as_double(0x7ff8000000000000ll);
        ObsGrid grid;
        grid.nx = p.grid_nx; grid.ny = p.grid_ny; grid.x0 = p.grid_x0; grid.y0 = p.grid_y0;
        grid.cell = p.grid_cell; grid.inv_cell = grid_cells > 0 ? 1.0 / p.grid_cell : 0.0;
        grid.cnt = reinterpret_cast<int32_t *>(links + p.node_cap) + 4 * (size_t)((p.node_cap + 1) / 2);
        grid.lists = reinterpret_cast<uint16_t *>(grid.cnt + grid_cells);
        if (lane == 0) { ws->q_stream = stream; ws->q_gcnt = grid.cnt; ws->q_glists = grid.lists; ws->q_ginv = grid.inv_cell; ws->q_elen = elen; }
        // every path point of an iteration's edges lies within `reach` of its new node (a snapped first edge starts up to
        // expand_dis + res away; near nodes are within the near radius, <= expand_dis unless near_r_max says otherwise)
        const double reach = (p.near_r_max > p.expand_dis ? p.near_r_max : p.expand_dis) + res;
        int n = 1, status = RRTK_Q_OK, gi = -1, it = 0, it_prev = 0;
        SobolState sob;
        sob.n = sobol_base < 0 ? 0 : sobol_base;
        if (!RESUME) {
            if (lane == 0) {
                xy[0] = make_double2(sg.x, sg.y);
                cost[0] = 0.0;
                parent[0] = -1;
                links[0] = make_int4(-1, -1, -1, 0);
                elen[0] = 0.0;
            }
        } else {
            // continue the tree a previous call left in xy / cost / parent (rows 0 .. n_nodes[q] - 1): rebuild the children
            // lists from the parent array (their order only fixes the traversal order of propagate, not its values) and
            // skip the Sobol points the earlier iterations consumed (one per non-goal coin)
            // A query that had finished (goal found in early-exit mode, or an overflow) stays as it is.
            n = n_nodes[q];
            it_prev = iters_done[q];
            if ((status_out[q] & (RRTK_Q_NEAR_OVERFLOW | RRTK_Q_NODE_OVERFLOW)) ||
                ((RRT_ONLY || !p.search_until_max_iter) && goal_index[q] >= 0))
                continue;   // uniform
            if (n < 1 || n > p.node_cap) { n = 1; status |= RRTK_Q_NODE_OVERFLOW; }
            for (int i = lane; i < n; i += 32) { links[i] = make_int4(-1, -1, -1, 0); elen[i] = NaN; }
            __syncwarp();
            if (lane == 0)
                for (int i = 1; i < n; i++) { const int pp = parent[i]; if (pp >= 0 && pp < n) link_child(links, pp, i); }
            if (p.sampler == RRTK_SAMPLER_SOBOL) {
                int used = 0;
                for (int k = lane; k < p.iter_offset; k += 32)
                    used += (int)(splitmix64(rng_key(p.seed, (uint64_t)(q + p.query_base), (uint64_t)k)) % 101ull) > p.goal_sample_rate;
                sob.n += (int64_t)__reduce_add_sync(FULL, (unsigned)used);
            }
        }
        __syncwarp();
        if (grid_cells > 0) build_obstacle_grid(grid, obs, n_obs, reach, lane);
        const double goal_reach = p.expand_dis > res ? p.expand_dis : res;
        const double inv_res = 1.0 / res, q_expand = floor(p.expand_dis / res);   // steer's n_expand at full extension
        sobol2(sob.n, sob.q0, sob.q1);
        if (lane == 0) { ws->sob_n = sob.n; ws->sob_q0 = sob.q0; ws->sob_q1 = sob.q1; }
        __syncwarp();
        bool done = false;
        auto mk_grid = [&]() {   // the obstacle cell grid, rebuilt from the parameter block + the warp's shared memory
            ObsGrid g;
            g.nx = p.grid_nx; g.ny = p.grid_ny; g.x0 = p.grid_x0; g.y0 = p.grid_y0; g.cell = p.grid_cell;
            g.inv_cell = ws->q_ginv; g.cnt = ws->q_gcnt; g.lists = ws->q_glists;
            return g;
        };

        for (it = 0; it < p.max_iter; it++) {
            SobolState sb;
            sb.n = ws->sob_n; sb.q0 = ws->sob_q0; sb.q1 = ws->sob_q1;
            Sample smp = draw_sample(p, (int)q, it, RESUME ? it + p.iter_offset : it, gx, gy, ws->q_stream, sb);
            __syncwarp();
            if (lane == 0) { ws->sob_n = sb.n; ws->sob_q0 = sb.q0; ws->sob_q1 = sb.q1; }
            const double rx = smp.x, ry = smp.y;
            // ---- get_nearest_node_index (rrt_04:1196-1202), merged with a SPECULATIVE find_near_nodes around the
            // sample: when the steered node snaps onto the sample (the common case once the tree is dense) the near
            // scan would compute exactly these d^2 again ----
            double bd = INF;
            int bi = 0x7fffffff, count = 0;
            const double r2 = RRT_ONLY ? -1.0 : near_r2[n + 1];
            // two chunks of node positions are kept in flight ahead of the one being processed (the scan is bound by
            // load latency: the live trees of the ~16 resident queries exceed L1)
            double2 a1 = lane < n ? xy[lane] : make_double2(0.0, 0.0);
            double2 a2 = lane + 32 < n ? xy[lane + 32] : make_double2(0.0, 0.0);
RRTK_UNROLL(RRTK_UNROLL_NEAREST)
            for (int b0 = 0; b0 < n; b0 += 32) {
                const int i = b0 + lane;
                const double2 a = a1;
                a1 = a2;
                if (i + 64 < n) a2 = xy[i + 64];
                bool hit = false;
                double d = 0.0;
                if (i < n) {
                    double ddx = a.x - rx, ddy = a.y - ry;
                    d = ddx * ddx + ddy * ddy;
                    if (d < bd) { bd = d; bi = i; }
                    hit = d <= r2;
                }
                unsigned mask = __ballot_sync(FULL, hit);
                int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (hit && pos < near_cap) { near_idx[pos] = i; nd[pos] = d; }
                count += __popc(mask);
            }
            warp_argmin(bd, bi);
            const int ni = bi;
            const double2 from = xy[ni];
            int t_status = 0, t_near = 0, t_par = -1, t_cpok = 0, t_rwok = 0, t_rwap = 0;
            bool accept = false, near_valid = false;
            ObsList L;
            L.ox = L.oy = L.r2 = nullptr; L.stride = 1; L.m = 0;
            // ---- steer towards the sample (rrt_04:1051-1052).  Fast form: if the edge certainly snaps onto the
            // sample the new node IS the sample and only the collision verdict is needed (edge_verdict_fast);
            // otherwise, or when the verdict is too close to call, the exact steer runs ----
            double nx = rx, ny = ry;
            // (the verdict keeps a 1e-9 margin on every use of the edge length: a plain sqrt is as good as the correctly
            // rounded hypot here; the exact steer computes its own)
#ifdef RRTK_V_HYPOT0
            const double d0 = crm_hypot(rx - from.x, ry - from.y);
#else
            const double ddx0 = rx - from.x, ddy0 = ry - from.y;
            const double d0 = sqrt(ddx0 * ddx0 + ddy0 * ddy0);
#endif
            int v = -1;
            {
                if (snap_certain(d0, false, p.expand_dis, q_expand, res, inv_res)) {
                    if (inside_play(p, nx, ny)) {
                        L = cull_obstacles_grid_mem(mk_grid(), obs, n_obs, nx, ny, reach, ws->cull_x, ws->cull_y, ws->cull_r2, lane);
                        const int vl = edge_verdict_fast<false>(from.x, from.y, rx, ry, d0, false, p.expand_dis, q_expand, res, inv_res, L,
                                                                lane, 32, ~0ull).v;
                        const unsigned blocked = __ballot_sync(FULL, vl == 0), unsure = __ballot_sync(FULL, vl < 0);
                        v = blocked ? 0 : (unsure ? -1 : 1);
                        if (v >= 0) { t_status = 1; accept = v == 1; near_valid = true; }
                    } else {
                        v = 0;  // outside the play area: rejected before the collision check (rrt_04:1054)
                    }
                }
            }
            if (RRTK_RARE(v < 0)) {
                Steer e0 = steer(from.x, from.y, rx, ry, p.expand_dis, res);
                nx = e0.ex; ny = e0.ey;
                if (inside_play(p, nx, ny)) {
                    t_status = 1;
                    L = cull_obstacles_grid_mem(mk_grid(), obs, n_obs, nx, ny, reach, ws->cull_x, ws->cull_y, ws->cull_r2, lane);
                    accept = edge_free_warp(from.x, from.y, e0, rx, ry, L, lane);
                }
            }
            {
                if (RRTK_RARE(accept && n >= p.node_cap)) { status |= RRTK_Q_NODE_OVERFLOW; accept = false; done = true; }
                if (accept && RRT_ONLY) {
                    if (lane == 0) { xy[n] = make_double2(nx, ny); cost[n] = 0.0; parent[n] = ni; }
                    t_status = 2; t_par = ni;
                    n++;
                    __syncwarp();
                } else if (accept) {
                    if (lane == 0) links[n] = make_int4(-1, -1, -1, 0);  // children arrive through rewire, before the append
                    const double nlen = crm_hypot(nx - from.x, ny - from.y);
                    const double ncost = cost[ni] + nlen;
                    // ---- find_near_nodes (rrt_04:1314-1338): ballot compaction, ascending index ----
                    if (RRTK_RARE(!near_valid)) count = near_scan(xy, n, nx, ny, r2, near_idx, nd, near_cap, lane);
                    __syncwarp();
                    if (RRTK_RARE(count > near_cap)) {
                        status |= RRTK_Q_NEAR_OVERFLOW;
                        done = true;
                    } else {
                        t_near = count;
                        // `.index()` quirk: every hit is replaced by the first hit with the same d2
                        // (the first 32 hits -- usually all of them -- by one warp MATCH on the d2 bit patterns: d2 is a finite
                        // non-negative double, so equal values are equal bits; idle lanes offer distinct NaN patterns)
                        {
                            const double d0k = lane < count ? nd[lane] : 0.0;
                            const unsigned long long key = lane < count ? (unsigned long long)__double_as_longlong(d0k)
                                                                        : (0xfff8000000000000ull | (unsigned)lane);
                            const unsigned same = __match_any_sync(FULL, key);
                            if (lane < count) near_ok[lane] = near_idx[__ffs(same) - 1];  // staged: resolved node index
                        }
                        for (int k = 32 + lane; k < count; k += 32) {
                            double dk = nd[k];
                            int f = k;
                            for (int j = 0; j < k; j++)
                                if (nd[j] == dk) { f = j; break; }
                            near_ok[k] = near_idx[f];
                        }
                        __syncwarp();
                        for (int k = lane; k < count; k += 32) near_idx[k] = near_ok[k];
                        __syncwarp();
                        // ---- choose_parent (rrt_04:1242-1282): one lane per candidate ----
                        double bc = INF, bex = 0.0, bey = 0.0;
                        int bk = 0x7fffffff;
                        for (int k = lane; k < count; k += 32) {
                            int i = near_idx[k];
                            double2 a = xy[i];
                            double ci = cost[i];
                            const double dk = crm_hypot(nx - a.x, ny - a.y);  // what steer's calc_distance_and_angle returns
                            double ex = nx, ey = ny;                          // a snapped edge ends on the new node
                            const int v = edge_verdict_fast<false>(a.x, a.y, nx, ny, dk, true, INF, INF, res, inv_res, L, 0, 1, ~0ull).v;
                            bool ok = v == 1;                                 // (the new node is inside the play area)
                            if (RRTK_RARE(v < 0)) {
                                Steer st = steer(a.x, a.y, nx, ny, INF, res);
                                ok = edge_free_lane(a.x, a.y, st, nx, ny, L) && inside_play(p, st.ex, st.ey);
                                ex = st.ex; ey = st.ey;
                            }
                            nd[k] = dk;     // = hypot(new - node), calc_new_cost's distance (rrt_04:1375-1377)
                            s_nc[k] = ci;
                            if (ok) {
                                t_cpok++;
                                double c = ci + dk;
                                if (c < bc) { bc = c; bk = k; bex = ex; bey = ey; }
                            }
                        }
                        warp_argmin(bc, bk);  // first minimum of the cost list
                        if (bk != 0x7fffffff) {
                            bex = __shfl_sync(FULL, bex, bk & 31);  // k = lane (mod 32): the winner's lane
                            bey = __shfl_sync(FULL, bey, bk & 31);
                        }
                        if (TRACE) {
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1) t_cpok += __shfl_xor_sync(FULL, t_cpok, off);
                        }
                        if (bk != 0x7fffffff) {
                            const int best = near_idx[bk];
                            // the node is re-steered from the winner (rrt_04:1279): same edge as above
                            const double cx = bex, cy = bey, ccost = bc;
                            const bool c_is_new = (cx == nx) && (cy == ny);  // the winner's edge snapped
                            // ---- rewire (rrt_04:1340-1373) ----
                            // An entry can only be re-parented if node.cost > new.cost + d (:1362); costs
                            // never increase while this loop runs (unless a node MOVES, handled below), so
                            // the steer + collision of an entry that fails the test now is dead work: only
                            // lanes whose entry passes evaluate their edge.  With a trace every edge is
                            // evaluated (the trace counts collision-free rewire edges).
                            bool dirty = false;      // a propagate ran: node costs must be re-read
                            int fallback_from = -1;  // >= 0: a node moved; redo entries from here serially
                            for (int b0 = 0; b0 < count && fallback_from < 0; b0 += 32) {
                                const int k = b0 + lane;
                                int i = -1;
                                double2 a = make_double2(0.0, 0.0);
                                double ecost = 0.0, snc = 0.0;
                                bool want = false, ok = false;
                                Steer st;
                                st.ex = st.ey = 0.0;
                                double dk = 0.0;
                                if (k < count) {
                                    i = near_idx[k];
                                    a = xy[i];
                                    snc = s_nc[k];
                                    // hypot(node - c) == the forward edge's d when c is the sample point itself
                                    dk = c_is_new ? nd[k] : crm_hypot(a.x - cx, a.y - cy);
                                    ecost = ccost + dk;
                                    want = TRACE || (snc > ecost);
                                }
                                if (want) {
                                    const int v = edge_verdict_fast<false>(cx, cy, a.x, a.y, dk, true, INF, INF, res, inv_res, L, 0, 1, ~0ull).v;
                                    if (RRTK_RARE(v < 0)) {
                                        st = steer(cx, cy, a.x, a.y, INF, res);
                                        ok = edge_free_lane(cx, cy, st, a.x, a.y, L) && inside_play(p, st.ex, st.ey);
                                    } else {   // snapped: the edge ends on the node itself (it does not move)
                                        st.ex = a.x; st.ey = a.y;
                                        ok = v == 1 && inside_play(p, a.x, a.y);
                                    }
                                }
                                const unsigned okmask = __ballot_sync(FULL, ok);
                                t_rwok += __popc(okmask);
                                unsigned m = __ballot_sync(FULL, ok && (snc > ecost));
                                while (m) {  // apply in list order
                                    const int b = __ffs(m) - 1;
                                    m &= m - 1;
                                    const int ii = __shfl_sync(FULL, i, b);
                                    const double ec = __shfl_sync(FULL, ecost, b);
                                    const double ex = __shfl_sync(FULL, st.ex, b), ey = __shfl_sync(FULL, st.ey, b);
                                    const double ax = __shfl_sync(FULL, a.x, b), ay = __shfl_sync(FULL, a.y, b);
                                    const double c0 = __shfl_sync(FULL, snc, b);
                                    const double dkb = __shfl_sync(FULL, dk, b);
                                    const double ci = dirty ? cost[ii] : c0;
                                    if (ci > ec) {
                                        const bool moved = (ax != ex) || (ay != ey);
                                        __syncwarp();
                                        if (lane == 0) {
                                            unlink_child(links, parent[ii], ii);
                                            link_child(links, n, ii);
                                            xy[ii] = make_double2(ex, ey);
                                            cost[ii] = ec;
                                            parent[ii] = n;
                                            elen[ii] = moved ? NaN : dkb;   // hypot(node - new node), what propagate would compute
                                        }
                                        __syncwarp();
                                        t_rwap++;
#ifdef RRTK_V_NOELEN
                                        propagate_lists(ii, xy, cost, links, &ws->qtail, lane);
#else
                                        propagate_lists_elen(ii, moved, xy, cost, elen, links, &ws->qtail, lane);
#endif
                                        dirty = true;
                                        if (RRTK_RARE(moved)) {
                                            // the node no longer sits where the parallel pass saw it, and the
                                            // costs of its descendants may have gone UP: every later entry is
                                            // re-evaluated from the current tree, one at a time
                                            fallback_from = b0 + b + 1;
                                            t_rwok -= __popc(okmask >> b >> 1);  // recounted below
                                            break;
                                        }
                                    }
                                }
                            }
                            if (RRTK_RARE(fallback_from >= 0))
                                rewire_serial(plan_consts(p), fallback_from, count, near_idx, xy, cost, parent, links, elen, n, cx, cy, ccost, L,
                                              &ws->qtail, lane, t_rwok, t_rwap);
                            if (lane == 0) {
                                xy[n] = make_double2(cx, cy); cost[n] = ccost; parent[n] = best; link_child(links, best, n);
                                elen[n] = c_is_new ? nd[bk] : NaN;   // hypot(new node - parent) as long as the node sits on the sample
                            }
                            t_status = 3; t_par = best;
                        } else {
                            if (lane == 0) {
                                xy[n] = make_double2(nx, ny); cost[n] = ncost; parent[n] = ni; link_child(links, ni, n);
                                elen[n] = nlen;
                            }
                            t_status = 2; t_par = ni;
                        }
                        n++;
                        __syncwarp();
                    }
                }
            }
            if (TRACE && lane == 0) {
                int32_t *tr = trace + (size_t)it * 8;
                tr[0] = ni; tr[1] = t_status; tr[2] = t_near; tr[3] = t_par; tr[4] = t_cpok;
                tr[5] = t_rwok; tr[6] = t_rwap; tr[7] = n;
            }
            if (done) { it++; break; }
            if (RRT_ONLY) {
                // goal test on the last node (rrt_01:90-96)
                double2 last = xy[n - 1];
                if (crm_hypot(last.x - gx, last.y - gy) <= p.expand_dis) {
                    ObsList G = cull_obstacles_mem(obs, n_obs, gx, gy, goal_reach, ws->cull_x, ws->cull_y,
                                               ws->cull_r2, lane);
                    Steer st = steer(last.x, last.y, gx, gy, p.expand_dis, res);
                    if (edge_free_warp(last.x, last.y, st, gx, gy, G, lane)) { gi = n - 1; it++; done = true; break; }
                }
            } else if (!p.search_until_max_iter) {
                bool ovf = false;
                // obstacles that can touch an edge into the goal (search_best_goal_node steers end there)
                ObsList G = cull_obstacles_mem(obs, n_obs, gx, gy, goal_reach, ws->cull_x, ws->cull_y,
                                           ws->cull_r2, lane);
                gi = best_goal(plan_consts(p), n, xy, cost, gx, gy, G, near_idx, nd, near_cap, lane, ovf);
                if (ovf) status |= RRTK_Q_NEAR_OVERFLOW;
                if (gi >= 0) { it++; done = true; break; }
            }
        }
        if (!done && !RRT_ONLY) {
            bool ovf = false;
            ObsList G = cull_obstacles_mem(obs, n_obs, gx, gy, goal_reach, ws->cull_x, ws->cull_y, ws->cull_r2, lane);
            gi = best_goal(plan_consts(p), n, xy, cost, gx, gy, G, near_idx, nd, near_cap, lane, ovf);
            if (ovf) status |= RRTK_Q_NEAR_OVERFLOW;
        }
        if (lane == 0) {
            n_nodes[q] = n;
            iters_done[q] = it_prev + it;
            goal_index[q] = gi;
            status_out[q] = status;
        }
        __syncwarp();
    }
}

// generate_final_course (rrt_04:1117-1125): one thread per query walks the parent chain
extern "C" __global__ void extract_paths_kernel(int n_queries, int node_cap, int path_cap,
                                                const double4 *__restrict__ start_goal,
                                                const double2 *__restrict__ xy,
                                                const int32_t *__restrict__ parent,
                                                const int32_t *__restrict__ goal_index,
                                                double2 *path, int32_t *path_len) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    int gi = goal_index[q];
    if (gi < 0) { path_len[q] = 0; return; }
    const double2 *t = xy + (size_t)q * node_cap;
    const int32_t *par = parent + (size_t)q * node_cap;
    double2 *out = path + (size_t)q * path_cap;
    int len = 0;
    double4 sg = start_goal[q];
    if (len < path_cap) out[len] = make_double2(sg.z, sg.w);
    len++;
    int i = gi;
    for (int guard = 0; guard <= node_cap; guard++) {
        if (len < path_cap) out[len] = t[i];
        len++;
        if (par[i] < 0) break;
        i = par[i];
    }
    path_len[q] = len;
}

// materialise the in-kernel sampler as a stream [Q][max_iter][2]
extern "C" __global__ void sample_stream_kernel(rrtk_rrtstar_params p, const double4 *__restrict__ start_goal,
                                                const int64_t *__restrict__ sobol_offset, double2 *out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n_queries) return;
    double4 sg = start_goal[q];
    SobolState sob;
    sob.n = sobol_offset ? (sobol_offset[q] < 0 ? 0 : sobol_offset[q]) : 0;
    sobol2(sob.n, sob.q0, sob.q1);
    for (int it = 0; it < p.max_iter; it++) {
        Sample s = draw_sample(p, q, it, it + p.iter_offset, sg.z, sg.w, nullptr, sob);
        out[(size_t)q * p.max_iter + it] = make_double2(s.x, s.y);
    }
}

extern "C" __global__ void crmath_probe_kernel(int kind, int64_t n, const double *a, const double *b,
                                               double *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c, r = 0.0;
    switch (kind) {
        case 0: r = crm_hypot(a[i], b[i]); break;
        case 1: r = crm_atan2(a[i], b[i]); break;
        case 2: r = crm_sin(a[i]); break;
        case 3: r = crm_cos(a[i]); break;
        case 4: (void)crm_atan2_sincos(a[i], b[i], &s, &c); r = s; break;
        case 5: (void)crm_atan2_sincos(a[i], b[i], &s, &c); r = c; break;
        case 6: r = crm_acos(a[i]); break;
        case 7: r = crm_asin(a[i]); break;
        case 8: r = crm_tan(a[i]); break;
    }
    out[i] = r;
}

}  // namespace rrtk

// ------------------------------------------------------------------------------------------------
// host side: launch wrappers (called by rrtk_api.cu)
// ------------------------------------------------------------------------------------------------
namespace rrtk {

bool rrtstar_cta_fits(const rrtk_rrtstar_params &p);
int rrtstar_cta_resident(const rrtk_rrtstar_params &p);   // CTAs (= queries) resident on the device at once
int launch_rrtstar_cta(const rrtk_rrtstar_params &p, const double *start_goal, const double *obstacles,
                       const int32_t *n_obs, const double *near_r2, const double *sample_stream,
                       const int64_t *sobol_offset, double *xy, double *cost, int32_t *parent,
                       int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                       int32_t *trace, int32_t *workspace, unsigned int *counter, cudaStream_t s);

int launch_rrtstar(const rrtk_rrtstar_params &p, const double *start_goal, const double *obstacles,
                   const int32_t *n_obs, const double *near_r2, const double *sample_stream,
                   const int64_t *sobol_offset, double *xy, double *cost, int32_t *parent,
                   int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                   int32_t *trace, int32_t *workspace, unsigned int *counter, cudaStream_t s) {
    // one CTA per query with the tree in shared memory (rrtk_rrtstar_cta.cu) when it fits AND the batch is small enough for
    // every query to be resident at once (then the launch time is one query's serial chain, which the CTA shortens);
    // larger batches are throughput-bound and one warp per query packs the SMs better (measured on B200, config 2:
    // 128 queries 17.9 vs 24.5 ms, 512 queries 25.0 vs 25.9 ms, 1024 queries 48.2 vs 32.2 ms, 4096 queries 175 vs 85 ms)
    if (p.exec_mode == RRTK_EXEC_CTA || (p.exec_mode == RRTK_EXEC_AUTO && rrtstar_cta_fits(p) && p.n_queries <= rrtstar_cta_resident(p)))
        return launch_rrtstar_cta(p, start_goal, obstacles, n_obs, near_r2, sample_stream, sobol_offset, xy, cost, parent,
                                  n_nodes, iters_done, goal_index, status, trace, workspace, counter, s);
    const bool fixed_nc = p.near_cap <= 256;
    size_t per_warp = warp_smem_bytes(fixed_nc ? 256 : p.near_cap, p.node_cap);
    size_t smem = per_warp * WARPS_PER_CTA;
    if (smem > 227 * 1024) return set_error(RRTK_ERR_INVALID, "near_cap/node_cap need more than 227 KB of shared memory");
    typedef void (*kernel_t)(rrtk_rrtstar_params, const double4 *, const double4 *, const int32_t *, const double *,
                             const double2 *, const int64_t *, double2 *, double *, int32_t *, int32_t *, int32_t *,
                             int32_t *, int32_t *, int32_t *, int32_t *, unsigned int *);
    const kernel_t kern =
        p.resume
            ? (fixed_nc ? (p.rrt_only ? rrtstar_kernel<true, false, 256, true> : rrtstar_kernel<false, false, 256, true>)
                        : (p.rrt_only ? rrtstar_kernel<true, false, 0, true> : rrtstar_kernel<false, false, 0, true>))
            : fixed_nc ? (p.rrt_only ? (trace ? rrtstar_kernel<true, true, 256> : rrtstar_kernel<true, false, 256>)
                                     : (trace ? rrtstar_kernel<false, true, 256> : rrtstar_kernel<false, false, 256>))
                       : (p.rrt_only ? (trace ? rrtstar_kernel<true, true, 0> : rrtstar_kernel<true, false, 0>)
                                     : (trace ? rrtstar_kernel<false, true, 0> : rrtstar_kernel<false, false, 0>));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(rrtstar_kernel)");
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS_PER_CTA * 32, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) per_sm = 1;
    long long want = ((long long)p.n_queries + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    long long grid = (long long)sms * per_sm;  // persistent: a multiple of the SM count
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counter)");
    kern<<<(unsigned)grid, WARPS_PER_CTA * 32, smem, s>>>(
        p, reinterpret_cast<const double4 *>(start_goal), reinterpret_cast<const double4 *>(obstacles),
        n_obs, near_r2, reinterpret_cast<const double2 *>(sample_stream), sobol_offset,
        reinterpret_cast<double2 *>(xy), cost, parent, n_nodes, iters_done, goal_index, status, trace,
        workspace, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rrtstar_kernel launch");
    return RRTK_OK;
}

int launch_extract_paths(int32_t nq, int32_t node_cap, int32_t path_cap, const double *start_goal,
                         const double *xy, const int32_t *parent, const int32_t *goal_index,
                         double *path, int32_t *path_len, cudaStream_t s) {
    int threads = 128;
    extract_paths_kernel<<<(nq + threads - 1) / threads, threads, 0, s>>>(
        nq, node_cap, path_cap, reinterpret_cast<const double4 *>(start_goal),
        reinterpret_cast<const double2 *>(xy), parent, goal_index, reinterpret_cast<double2 *>(path),
        path_len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "extract_paths_kernel launch");
    return RRTK_OK;
}

int launch_sample_stream(const rrtk_rrtstar_params &p, const double *start_goal,
                         const int64_t *sobol_offset, double *out, cudaStream_t s) {
    int threads = 64;
    sample_stream_kernel<<<(p.n_queries + threads - 1) / threads, threads, 0, s>>>(
        p, reinterpret_cast<const double4 *>(start_goal), sobol_offset, reinterpret_cast<double2 *>(out));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "sample_stream_kernel launch");
    return RRTK_OK;
}

// steer (rrt_04:1086-1115) + check_collision (:1216-1230) + check_if_outside_play_area (:1204-1214) of N independent edges,
// one lane each: the stand-alone form of what the planner kernel does inside an iteration.
__global__ void steer_collide_kernel(long long n_req, const double2 *__restrict__ from_xy, const double2 *__restrict__ to_xy,
                                     double extend, double res, const int32_t *__restrict__ obs_set,
                                     const double4 *__restrict__ obstacles, int obs_stride, const int32_t *__restrict__ n_obs_arr,
                                     const double *__restrict__ play, double2 *new_xy, double *dist, int32_t *n_points,
                                     uint8_t *free_flag, uint8_t *inside_flag) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_req) return;
    const double2 f = from_xy[r], t = to_xy[r];
    const Steer st = steer(f.x, f.y, t.x, t.y, extend, res);
    const int set = obs_set ? obs_set[r] : 0;
    const double4 *obs = obstacles + (size_t)set * obs_stride;
    ObsList L;
    L.ox = &obs->x; L.oy = &obs->y; L.r2 = &obs->w; L.stride = 4; L.m = n_obs_arr ? n_obs_arr[set] : 0;
    new_xy[r] = make_double2(st.ex, st.ey);
    dist[r] = st.d;
    n_points[r] = 1 + st.n + (st.snap ? 1 : 0);   // len(path_x): the start, n_expand steps, the target when snapped
    free_flag[r] = edge_free_lane(f.x, f.y, st, t.x, t.y, L) ? 1 : 0;
    // check_if_outside_play_area: True (ok) without a play area, else xmin <= x <= xmax and ymin <= y <= ymax
    inside_flag[r] = (!play || !(st.ex < play[0] || st.ex > play[1] || st.ey < play[2] || st.ey > play[3])) ? 1 : 0;
}

int launch_steer_collide(long long n_req, const double *from_xy, const double *to_xy, double extend, double res,
                         const int32_t *obs_set, const double *obstacles, int obs_stride, const int32_t *n_obs, const double *play,
                         double *new_xy, double *dist, int32_t *n_points, uint8_t *free_flag, uint8_t *inside_flag, cudaStream_t s) {
    const int threads = 128;
    steer_collide_kernel<<<(unsigned)((n_req + threads - 1) / threads), threads, 0, s>>>(
        n_req, reinterpret_cast<const double2 *>(from_xy), reinterpret_cast<const double2 *>(to_xy), extend, res, obs_set,
        reinterpret_cast<const double4 *>(obstacles), obs_stride, n_obs, play, reinterpret_cast<double2 *>(new_xy), dist,
        n_points, free_flag, inside_flag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "steer_collide_kernel launch");
    return RRTK_OK;
}

int launch_crmath_probe(int kind, int64_t n, const double *a, const double *b, double *out, cudaStream_t s) {
    int threads = 128;
    crmath_probe_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(kind, n, a, b, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "crmath_probe_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
