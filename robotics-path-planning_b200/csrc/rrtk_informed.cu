// rrtk_informed.cu -- Informed RRT* (rrt_07:1027-1285): informed_rrt_star_search for Q independent queries,
// one warp per query, persistent grid.  Differences from the RRT* kernel that matter for the layout:
//   * fixed-length extension (get_new_node :1216-1224), int parents, NO cost propagation (:1232-1246);
//   * continuous segment-vs-circle collision (distance_squared_point_to_segment :1249-1261) with numpy's
//     fma dot, against obstacles culled to the near disc;
//   * un-clipped near radius 50*sqrt(log n / n) (:1137-1143): the near list can hold most of the tree, so it
//     lives in a global-memory workspace, not in shared memory;
//   * ellipsoidal sampling once a solution exists (informed_sample :1145-1159): the sample depends on c_best,
//     so it is drawn in-kernel from per-iteration raw draws (free-space sample, two unit uniforms);
//   * the best path is a SNAPSHOT taken when c_best improves (:1094-1103): later rewires do not change it.
// All arithmetic is FP64 in the reference's order with crmath.h leaf functions.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"

namespace rrtk {

constexpr int INF_WARPS_PER_CTA = 4;

// numpy's 2-vector dot on the reference platform: fma(a1, b1, a0 * b0)
__device__ __forceinline__ double idot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

// distance_squared_point_to_segment(v, w, p) <= size**2  (rrt_07:1249-1269) for one circle
__device__ __forceinline__ bool seg_hits(double x1, double y1, double x2, double y2, double ox, double oy, double r2) {
    double dd;
    if (x1 == x2 && y1 == y2) {
        dd = idot2(ox - x1, oy - y1, ox - x1, oy - y1);
    } else {
        double wx = x2 - x1, wy = y2 - y1;
        double l2 = idot2(wx, wy, wx, wy);
        double t = idot2(ox - x1, oy - y1, wx, wy) / l2;
        t = t < 1.0 ? t : 1.0;  // max(0, min(1, t)) with Python's NaN behaviour
        t = t > 0.0 ? t : 0.0;
        double px = x1 + t * wx, py = y1 + t * wy;
        dd = idot2(ox - px, oy - py, ox - px, oy - py);
    }
    return dd <= r2;
}

// check_segment_collision by ONE lane over an obstacle list. true = free
__device__ __noinline__ bool seg_free_lane(double x1, double y1, double x2, double y2, const ObsList &L) {
    for (int j = 0; j < L.m; j++)
        if (seg_hits(x1, y1, x2, y2, L.ox[j * L.stride], L.oy[j * L.stride], L.r2[j * L.stride])) return false;
    return true;
}

// the same verdict by the whole warp (lanes split the obstacles); uniform result
__device__ __forceinline__ bool seg_free_warp(double x1, double y1, double x2, double y2, const ObsList &L, int lane) {
    bool hit = false;
    for (int j = lane; j < L.m && !hit; j += 32)
        hit = seg_hits(x1, y1, x2, y2, L.ox[j * L.stride], L.oy[j * L.stride], L.r2[j * L.stride]);
    return __ballot_sync(FULL, hit) == 0u;
}

// distance_squared_point_to_segment for one circle, value only
__device__ __forceinline__ double seg_dd(double x1, double y1, double x2, double y2, double ox, double oy) {
    if (x1 == x2 && y1 == y2) return idot2(ox - x1, oy - y1, ox - x1, oy - y1);
    double wx = x2 - x1, wy = y2 - y1;
    double l2 = idot2(wx, wy, wx, wy);
    double t = idot2(ox - x1, oy - y1, wx, wy) / l2;
    t = t < 1.0 ? t : 1.0;
    t = t > 0.0 ? t : 0.0;
    double px = x1 + t * wx, py = y1 + t * wy;
    return idot2(ox - px, oy - py, ox - px, oy - py);
}

// the reference's verdict for the edge a -> new node: check_collision(node, theta, d) (rrt_07:1271-1276) ends the
// segment at a + (cos, sin)(theta) * d, a few ulp from the new node itself
__device__ __noinline__ bool near_edge_free_exact(double ax, double ay, double nx, double ny, double d, const ObsList &L) {
    double s, c;
    (void)crm_atan2_sincos(ny - ay, nx - ax, &s, &c);
    return seg_free_lane(ax, ay, ax + c * d, ay + s * d, L);
}

// Same verdict without the trigonometry whenever it is certain: the segment a -> new node differs from the reference's
// by a few ulp at one end, so |dd - size^2| above an error band decides; inside the band the exact form runs.
// band_k = 1e-12 * (coordinate bound)^2.
__device__ __forceinline__ bool near_edge_free(double ax, double ay, double nx, double ny, double d, double band_k,
                                               const ObsList &L) {
    bool hit = false, unsure = !(d > 1e-9);
    for (int j = 0; j < L.m; j++) {
        const double r2 = L.r2[j * L.stride];
        const double dd = seg_dd(ax, ay, nx, ny, L.ox[j * L.stride], L.oy[j * L.stride]);
        hit |= dd <= r2;
        unsure |= fabs(dd - r2) <= band_k + 1e-12 * (dd + r2);
    }
    if (unsure) return near_edge_free_exact(ax, ay, nx, ny, d, L);
    return !hit;
}

// slots of the per-warp duplicate detector (a power of two; lists longer than 3/4 of it take the exact mapping)
constexpr int INF_HASH = 1024;
struct InfWarpSmem {
    double cull_x[CULL_CAP], cull_y[CULL_CAP], cull_r2[CULL_CAP];
    unsigned long long dup[INF_HASH];
};

// TW = 1: one warp per query.  TW = 4: one CTA per query, for batches that are resident all at once (a launch of up to a
// few hundred queries lasts as long as ONE query's chain, and rrt_07's near lists -- a quarter of the tree -- are several
// rounds of 32 candidates): every warp runs the same control flow on the same data (sample, nearest, steer, cull, first
// segment: recomputed per warp), warp 0 writes the near list and the tree, the candidates of choose_parent and rewire
// are split over the 128 threads.  Same trees bit for bit.
template <int TW>
__global__ void __launch_bounds__(INF_WARPS_PER_CTA * 32, 4)
informed_kernel(rrtk_informed_params p, const double4 *__restrict__ start_goal, const double4 *__restrict__ rot,
                const double4 *__restrict__ obstacles, const int32_t *__restrict__ n_obs_arr,
                const double2 *__restrict__ near_rr2, const double2 *__restrict__ free_samples,
                const double2 *__restrict__ ball_draws, double2 *xy_all, double *cost_all, int32_t *parent_all,
                int32_t *n_nodes, double2 *path_all, int32_t *path_len, double *c_best_out, int32_t *status_out,
                int32_t *ws_idx_all, double *ws_d_all, unsigned int *counter) {
    static_assert(TW == 1 || TW == INF_WARPS_PER_CTA, "a team is one warp or the whole CTA");
    __shared__ InfWarpSmem smem[INF_WARPS_PER_CTA];
    __shared__ double t_mc[INF_WARPS_PER_CTA];    // team reduction of choose_parent
    __shared__ int t_bk[INF_WARPS_PER_CTA], t_bnode[INF_WARPS_PER_CTA];
    __shared__ unsigned int t_q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lead = TW == 1 || warp == 0;       // the warp that writes the near list and the tree
    const int kfirst = TW == 1 ? lane : (int)threadIdx.x, kstride = 32 * TW;
    InfWarpSmem *ws = &smem[warp];
    const double INF = CUDART_INF;
    const double ed = p.expand_dis;
    auto team_sync = [&]() { if (TW > 1) __syncthreads(); else __syncwarp(); };

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
        const double4 sg = start_goal[q];
        const double4 R = rot[q];  // c00, c01, c10, c11
        const double sx = sg.x, sy = sg.y, gx = sg.z, gy = sg.w;
        const double4 *obs = obstacles + (size_t)q * p.obs_stride;
        const int n_obs = n_obs_arr[q];
        double2 *xy = xy_all + (size_t)q * p.node_cap;
        double *cost = cost_all + (size_t)q * p.node_cap;
        int32_t *parent = parent_all + (size_t)q * p.node_cap;
        int32_t *near_idx = ws_idx_all + (size_t)q * p.node_cap;
        double *near_d = ws_d_all + (size_t)q * p.node_cap;
        double2 *path = path_all + (size_t)q * p.path_cap;
        const double2 *fs = free_samples + (size_t)q * p.max_iter;
        const double2 *bd = ball_draws + (size_t)q * p.max_iter;
        if (lead && lane == 0) { xy[0] = make_double2(sx, sy); cost[0] = 0.0; parent[0] = -1; }
        int n = 1, status = RRTK_Q_OK, plen_best = 0;
        double c_best = INF;
        const double band_k = 1e-12 * p.coord_bound * p.coord_bound;
        const double c_min = crm_hypot(sx - gx, sy - gy);
        const double xc = (sx + gx) / 2.0, yc = (sy + gy) / 2.0;

        for (int it = 0; it < p.max_iter; it++) {
            team_sync();   // the tree as the previous iteration left it (append, rewire) is visible to every warp
            // ---- informed_sample (rrt_07:1145-1159) ----
            double rx, ry;
            if (c_best < INF) {
                const double r0 = c_best / 2.0;
                const double r1 = sqrt(c_best * c_best - c_min * c_min) / 2.0;
                double2 ab = bd[it];
                double a = ab.x, b = ab.y;
                if (b < a) { double t = a; a = b; b = t; }
                const double ang = 2 * 3.141592653589793 * a / b;  // 2 * math.pi * a / b
                double sn, cs;
                crm_sincos(ang, &sn, &cs);
                const double bx = b * cs, by = b * (ang == 0.0 ? ang : sn);
                const double m00 = R.x * r0, m01 = R.y * r1, m10 = R.z * r0, m11 = R.w * r1;
                rx = fma(m00, bx, m01 * by) + xc;  // numpy (3x3)@(3x1) on the reference platform
                ry = fma(m10, bx, m11 * by) + yc;
            } else {
                double2 f = fs[it];
                rx = f.x; ry = f.y;
            }
            // ---- get_nearest_list_index (rrt_07:1210-1214) ----
            double bdist = INF;
            int bi = 0x7fffffff;
#pragma unroll 1
            for (int i = lane; i < n; i += 32) {
                double2 a = xy[i];
                double ddx = a.x - rx, ddy = a.y - ry;
                double d = ddx * ddx + ddy * ddy;
                if (d < bdist) { bdist = d; bi = i; }
            }
            warp_argmin(bdist, bi);
            const int ni = bi;
            const double2 from = xy[ni];
            double st, ct;
            (void)crm_atan2_sincos(ry - from.y, rx - from.x, &st, &ct);
            const double nx = from.x + ed * ct, ny = from.y + ed * st;  // get_new_node (:1216-1224)
            double ncost = cost[ni] + ed;
            int npar = ni;
            const double d0 = crm_hypot(from.x - nx, from.y - ny);
            if (n >= p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; break; }
            // obstacles that can touch any segment of this iteration: all of them lie in the disc of radius
            // max(near radius, d0) around the new node
            const double2 rr2 = near_rr2[n];  // (r, r**2) for n_node = n (rrt_07:1139), host-evaluated
            const double reach = rr2.x > d0 ? rr2.x : d0;
            ObsList L = cull_obstacles(obs, n_obs, nx, ny, reach, ws->cull_x, ws->cull_y, ws->cull_r2, lane);
            // ---- check_collision(nearest, theta, d) (:1271-1276) ----
            if (!seg_free_warp(from.x, from.y, from.x + ct * d0, from.y + st * d0, L, lane)) continue;
            // ---- find_near_nodes (:1137-1143) ----
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
                    hit = d <= rr2.y;
                }
                unsigned mask = __ballot_sync(FULL, hit);
                int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (hit && lead) { near_idx[pos] = i; near_d[pos] = d; }
                count += __popc(mask);
            }
            team_sync();
            // ---- the `.index()` mapping of find_near_nodes (:1140-1143): every hit stands for the FIRST hit with the same
            // d2.  The near lists of rrt_07 hold a quarter of the tree (the radius is not clipped), and looking for an equal
            // d2 among the earlier hits is quadratic: ~175 dependent loads per candidate, twice per iteration, were 40 % of an
            // iteration.  Equal d2 of different nodes are rare, so a shared-memory hash set of the d2 bit patterns first says
            // whether ANY two hits are equal; only then does the exact mapping run (once, in place: an entry that changes is
            // not a first occurrence, and only first occurrences are read) ----
            {
                bool dup = count > INF_HASH * 3 / 4;
                if (!dup && count > 1) {
                    for (int t = lane; t < INF_HASH; t += 32) ws->dup[t] = ~0ull;
                    __syncwarp();
                    for (int k = lane; k < count && !dup; k += 32) {
                        const unsigned long long key = (unsigned long long)__double_as_longlong(near_d[k]);
                        unsigned slot = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 54) & (INF_HASH - 1);
                        for (;;) {
                            const unsigned long long old = atomicCAS(&ws->dup[slot], ~0ull, key);
                            if (old == ~0ull) break;
                            if (old == key) { dup = true; break; }
                            slot = (slot + 1) & (INF_HASH - 1);
                        }
                    }
                    dup = __any_sync(FULL, dup);
                }
                if (dup) {   // (uniform over the team: every warp looked at the same list)
                    for (int k = kfirst; k < count; k += kstride) {
                        const double dk = near_d[k];
                        for (int j = 0; j < k; j++)
                            if (near_d[j] == dk) { near_idx[k] = near_idx[j]; break; }
                    }
                    team_sync();
                }
                __syncwarp();
            }
            // ---- choose_parent (:1110-1135): lane per candidate ----
            double mc = INF;
            int bk = 0x7fffffff, bnode = -1;
#pragma unroll 1
            for (int k = kfirst; k < count; k += kstride) {
                const int i = near_idx[k];
                const double2 a = xy[i];
                const double dx = nx - a.x, dy = ny - a.y;
                const double dd = crm_hypot(dx, dy);
                if (near_edge_free(a.x, a.y, nx, ny, dd, band_k, L)) {
                    const double c = cost[i] + dd;
                    if (c < mc) { mc = c; bk = k; bnode = i; }
                }
            }
            warp_argmin(mc, bk);
            if (bk != 0x7fffffff) bnode = __shfl_sync(FULL, bnode, bk & 31);   // (candidate k sits on lane k & 31 of its warp)
            if (TW > 1) {   // the first minimum over the team's warps
                if (lane == 0) { t_mc[warp] = mc; t_bk[warp] = bk; t_bnode[warp] = bnode; }
                __syncthreads();
                mc = t_mc[0]; bk = t_bk[0]; bnode = t_bnode[0];
                for (int w = 1; w < TW; w++)
                    if (t_mc[w] < mc || (t_mc[w] == mc && t_bk[w] < bk)) { mc = t_mc[w]; bk = t_bk[w]; bnode = t_bnode[w]; }
            }
            if (bk != 0x7fffffff) {
                ncost = mc;
                npar = bnode;
            }
            const int newi = n;
            __syncwarp();
            if (lead && lane == 0) { xy[newi] = make_double2(nx, ny); cost[newi] = ncost; parent[newi] = npar; }
            n++;
            __syncwarp();
            // ---- rewire (:1232-1246): entries are independent (no propagation); repeats are idempotent ----
#pragma unroll 1
            for (int b0 = 0; b0 < count; b0 += kstride) {
                const int k = b0 + kfirst;
                if (k < count) {
                    const int i = near_idx[k];
                    const double2 a = xy[i];
                    const double dd = crm_hypot(a.x - nx, a.y - ny);
                    const double sc = ncost + dd;
                    if (cost[i] > sc) {
                        if (near_edge_free(a.x, a.y, nx, ny, dd, band_k, L)) {
                            parent[i] = newi;
                            cost[i] = sc;
                        }
                    }
                }
                __syncwarp();
            }
            // ---- goal bookkeeping (:1094-1103) ----
            if (crm_hypot(nx - gx, ny - gy) < ed) {
                team_sync();   // the walk below reads parents the other warps may just have rewired
                ObsList G = cull_obstacles(obs, n_obs, nx, ny, ed, ws->cull_x, ws->cull_y, ws->cull_r2, lane);
                if (seg_free_warp(nx, ny, gx, gy, G, lane)) {
                    // get_final_course + get_path_len: goal, new node, ..., root, start; serial walk (lane 0)
                    double plen = 0.0;
                    int len = 0;
                    if (lane == 0) {
                        double qx = gx, qy = gy;
                        int k = newi;
                        len = 1;
                        for (int guard = 0; guard <= p.node_cap && parent[k] >= 0; guard++) {
                            double2 a = xy[k];
                            plen += crm_hypot(a.x - qx, a.y - qy);
                            qx = a.x; qy = a.y;
                            k = parent[k];
                            len++;
                        }
                        plen += crm_hypot(sx - qx, sy - qy);
                        len++;
                    }
                    plen = __shfl_sync(FULL, plen, 0);
                    len = __shfl_sync(FULL, len, 0);
                    if (plen < c_best) {
                        c_best = plen;
                        plen_best = len;
                        if (len > p.path_cap) status |= RRTK_Q_PATH_OVERFLOW;
                        if (lead && lane == 0) {
                            int w = 0;
                            if (w < p.path_cap) path[w] = make_double2(gx, gy);
                            w++;
                            for (int k = newi; parent[k] >= 0 && w <= p.node_cap + 1; k = parent[k], w++)
                                if (w < p.path_cap) path[w] = xy[k];
                            if (w < p.path_cap) path[w] = make_double2(sx, sy);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (lead && lane == 0) {
            n_nodes[q] = n;
            path_len[q] = plen_best;
            c_best_out[q] = c_best;
            status_out[q] = status;
        }
        __syncwarp();
    }
}

int launch_informed(const rrtk_informed_params &p, const double *start_goal, const double *rot,
                    const double *obstacles, const int32_t *n_obs, const double *near_rr2, const double *free_s,
                    const double *ball, double *xy, double *cost, int32_t *parent, int32_t *n_nodes, double *path,
                    int32_t *path_len, double *c_best, int32_t *status, int32_t *ws_idx, double *ws_d,
                    unsigned int *counter, cudaStream_t s) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, informed_kernel<INF_WARPS_PER_CTA>, INF_WARPS_PER_CTA * 32, 0);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) per_sm = 1;
    // a CTA per query while every query of the batch is resident at once (p.exec_mode: RRTK_EXEC_WARP / _CTA force one)
    const bool cta = p.exec_mode == RRTK_EXEC_CTA || (p.exec_mode != RRTK_EXEC_WARP && (long long)p.n_queries <= (long long)sms * per_sm);
    const auto kern = cta ? informed_kernel<INF_WARPS_PER_CTA> : informed_kernel<1>;
    long long want = cta ? (long long)p.n_queries : ((long long)p.n_queries + INF_WARPS_PER_CTA - 1) / INF_WARPS_PER_CTA;
    long long grid = (long long)sms * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counter)");
    kern<<<(unsigned)grid, INF_WARPS_PER_CTA * 32, 0, s>>>(
        p, reinterpret_cast<const double4 *>(start_goal), reinterpret_cast<const double4 *>(rot),
        reinterpret_cast<const double4 *>(obstacles), n_obs, reinterpret_cast<const double2 *>(near_rr2),
        reinterpret_cast<const double2 *>(free_s), reinterpret_cast<const double2 *>(ball),
        reinterpret_cast<double2 *>(xy), cost, parent, n_nodes, reinterpret_cast<double2 *>(path), path_len, c_best,
        status, ws_idx, ws_d, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "informed_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
