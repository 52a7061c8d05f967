// rrtk_planner.cuh -- device helpers shared by the planner kernels (RRT*, Informed RRT*).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crmath.h"

namespace rrtk {

constexpr int CULL_CAP = 64;      // culled obstacle list per warp (overflow -> full list)
constexpr unsigned FULL = 0xffffffffu;

struct ObsList {  // SoA view of the obstacles an edge must be tested against
    const double *ox, *oy, *r2;
    int stride;  // element stride (1 for the shared-memory list, 4 for the global AoS rows)
    int m;
};

// warp argmin of (value, index): smaller value wins, ties -> smaller index (list.index(min(..))).  Values are
// non-negative or +inf, so their bit patterns order like the values: three integer REDUX ops find the minimum
// (hi word, lo word, index) instead of five shuffle rounds over 12 bytes (less code in the hot loop).
static __device__ __forceinline__ void warp_argmin(double &v, int &i) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned mhi = __reduce_min_sync(FULL, hi);
    const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
    const bool is_min = hi == mhi && lo == mlo;
    const unsigned mi = __reduce_min_sync(FULL, is_min ? (unsigned)i : 0xffffffffu);
    v = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    i = (int)mi;
}

// Conservative exact cull: keep obstacle o iff |o - c| <= (reach + R_o) * (1 + 1e-9) + 1e-9.
// Returns the list to test edges against (shared-memory survivors, or all obstacles on overflow).
// (out of line they return ONE int -- a 32-byte ObsList comes back through local memory: the survivor count, or -1 when
// the shared-memory list overflowed and all circles must be tested; the inline wrappers below build the ObsList)
static __device__ __forceinline__ int cull_obstacles_impl(const double4 *obs, int n_obs, double cx, double cy,
                                                          double reach, double *sx, double *sy, double *sr2,
                                                          int lane) {
    int count = 0;
    bool overflow = false;
#pragma unroll 2
    for (int base = 0; base < n_obs; base += 32) {
        int j = base + lane;
        bool keep = false;
        double4 o = make_double4(0, 0, 0, 0);
        if (j < n_obs) {
            o = obs[j];
            double dx = o.x - cx, dy = o.y - cy;
            double lim = (reach + o.z) * (1.0 + 1e-9) + 1e-9;
            keep = dx * dx + dy * dy <= lim * lim;
        }
        unsigned mask = __ballot_sync(FULL, keep);
        int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (keep) {
            if (pos < CULL_CAP) { sx[pos] = o.x; sy[pos] = o.y; sr2[pos] = o.w; }
            else overflow = true;
        }
        count += __popc(mask);
    }
    overflow = __any_sync(FULL, overflow);
    __syncwarp();
    return overflow ? -1 : count;
}
static __device__ __noinline__ int cull_obstacles_count(const double4 *obs, int n_obs, double cx, double cy, double reach,
                                                        double *sx, double *sy, double *sr2, int lane) {
    return cull_obstacles_impl(obs, n_obs, cx, cy, reach, sx, sy, sr2, lane);
}
static __device__ __forceinline__ ObsList obs_list_of(int m, const double4 *obs, int n_obs, double *sx, double *sy, double *sr2) {
    ObsList L;
    if (m >= 0) {
        L.ox = sx; L.oy = sy; L.r2 = sr2; L.stride = 1; L.m = m;
    } else {
        const double *g = reinterpret_cast<const double *>(obs);
        L.ox = g; L.oy = g + 1; L.r2 = g + 3; L.stride = 4; L.m = n_obs;
    }
    return L;
}
static __device__ __forceinline__ ObsList cull_obstacles(const double4 *obs, int n_obs, double cx, double cy, double reach,
                                                         double *sx, double *sy, double *sr2, int lane) {
    return obs_list_of(cull_obstacles_count(obs, n_obs, cx, cy, reach, sx, sy, sr2, lane), obs, n_obs, sx, sy, sr2);
}

// ---- obstacle cell grid (rrtk_rrtstar_params.grid_*): per cell the circles that can pass the cull test of ANY point
// of the cell.  cnt[cell] counts them (more than GRID_CELL_CAP = the cell overflowed: callers scan all circles),
// lists[cell][k] holds their indices.  Built once per query by its warp; read with ld.cg (the counts are bumped by
// L2 atomics, which L1 does not see). ----
constexpr int GRID_CELL_CAP = 32;
struct ObsGrid {
    int nx, ny;
    double x0, y0, cell, inv_cell;
    int32_t *cnt;
    uint16_t *lists;
};

static __device__ __noinline__ void build_obstacle_grid(const ObsGrid &g, const double4 *obs, int n_obs, double reach,
                                                    int lane) {
    const int cells = g.nx * g.ny;
    for (int c = lane; c < cells; c += 32) atomicExch(&g.cnt[c], 0);   // (atomics: the counts live in L2 only)
    __syncwarp();
    __threadfence_block();
    for (int j = lane; j < n_obs; j += 32) {
        const double4 o = obs[j];
        const double lim = ((reach + o.z) * (1.0 + 1e-9) + 1e-9) * (1.0 + 1e-9) + 1e-9;  // cull limit + cell-edge rounding
        int ix0 = (int)floor((o.x - lim - g.x0) * g.inv_cell) - 1, ix1 = (int)floor((o.x + lim - g.x0) * g.inv_cell) + 1;
        int iy0 = (int)floor((o.y - lim - g.y0) * g.inv_cell) - 1, iy1 = (int)floor((o.y + lim - g.y0) * g.inv_cell) + 1;
        ix0 = ix0 < 0 ? 0 : ix0; iy0 = iy0 < 0 ? 0 : iy0;
        ix1 = ix1 > g.nx - 1 ? g.nx - 1 : ix1; iy1 = iy1 > g.ny - 1 ? g.ny - 1 : iy1;
        for (int iy = iy0; iy <= iy1; iy++) {
            const double ry0 = g.y0 + iy * g.cell, ry1 = g.y0 + (iy + 1) * g.cell;
            const double dy = o.y < ry0 ? ry0 - o.y : (o.y > ry1 ? o.y - ry1 : 0.0);
            for (int ix = ix0; ix <= ix1; ix++) {
                const double rx0 = g.x0 + ix * g.cell, rx1 = g.x0 + (ix + 1) * g.cell;
                const double dx = o.x < rx0 ? rx0 - o.x : (o.x > rx1 ? o.x - rx1 : 0.0);
                if (dx * dx + dy * dy <= lim * lim) {
                    const int c = iy * g.nx + ix;
                    const int pos = atomicAdd(&g.cnt[c], 1);
                    if (pos < GRID_CELL_CAP) g.lists[c * GRID_CELL_CAP + pos] = (uint16_t)j;
                }
            }
        }
    }
    __syncwarp();
}

// cull through the grid: exact same keep test as cull_obstacles on the cell's candidates.  Falls back to the full scan
// when the point is outside the grid or its cell overflowed.
template <bool MEM>
static __device__ __forceinline__ int cull_obstacles_grid_impl(const ObsGrid g, const double4 *obs, int n_obs, double cx,
                                                               double cy, double reach, double *sx, double *sy, double *sr2,
                                                               int lane) {
    const double fx = floor((cx - g.x0) * g.inv_cell), fy = floor((cy - g.y0) * g.inv_cell);
    int m = GRID_CELL_CAP + 1, c = 0;
    if (g.nx > 0 && fx >= 0.0 && fy >= 0.0 && fx < (double)g.nx && fy < (double)g.ny) {
        c = (int)fy * g.nx + (int)fx;
        m = __ldcg(g.cnt + c);
    }
    if (m > GRID_CELL_CAP) return cull_obstacles_count(obs, n_obs, cx, cy, reach, sx, sy, sr2, lane);   // (rare: one shared copy)
    bool keep = false;
    double4 o = make_double4(0, 0, 0, 0);
    if (lane < m) {
        o = obs[__ldcg(g.lists + c * GRID_CELL_CAP + lane)];
        const double dx = o.x - cx, dy = o.y - cy;
        const double lim = (reach + o.z) * (1.0 + 1e-9) + 1e-9;
        keep = dx * dx + dy * dy <= lim * lim;
    }
    const unsigned mask = __ballot_sync(FULL, keep);
    const int pos = __popc(mask & ((1u << lane) - 1u));
    if (keep) { sx[pos] = o.x; sy[pos] = o.y; sr2[pos] = o.w; }
    __syncwarp();
    return __popc(mask);
}
static __device__ __noinline__ int cull_obstacles_grid_count(const ObsGrid g, const double4 *obs, int n_obs, double cx, double cy,
                                                             double reach, double *sx, double *sy, double *sr2, int lane) {
    return cull_obstacles_grid_impl<false>(g, obs, n_obs, cx, cy, reach, sx, sy, sr2, lane);
}
static __device__ __forceinline__ ObsList cull_obstacles_grid(const ObsGrid g, const double4 *obs, int n_obs, double cx, double cy,
                                                              double reach, double *sx, double *sy, double *sr2, int lane) {
    return obs_list_of(cull_obstacles_grid_count(g, obs, n_obs, cx, cy, reach, sx, sy, sr2, lane), obs, n_obs, sx, sy, sr2);
}

// The same two culls returning the ObsList itself from the out-of-line body (through local memory).  The warp-per-query
// kernel uses these: it sits at the 128-register cap, and a list whose five fields are materialised in registers by the
// inline wrappers above costs it more (measured: 78.8 -> 85.3 ms per 4096-query launch) than reloading them from the stack.
static __device__ __noinline__ ObsList cull_obstacles_mem(const double4 *obs, int n_obs, double cx, double cy, double reach,
                                                          double *sx, double *sy, double *sr2, int lane) {
    return obs_list_of(cull_obstacles_impl(obs, n_obs, cx, cy, reach, sx, sy, sr2, lane), obs, n_obs, sx, sy, sr2);
}
static __device__ __noinline__ ObsList cull_obstacles_grid_mem(const ObsGrid g, const double4 *obs, int n_obs, double cx, double cy,
                                                               double reach, double *sx, double *sy, double *sr2, int lane) {
    return obs_list_of(cull_obstacles_grid_impl<true>(g, obs, n_obs, cx, cy, reach, sx, sy, sr2, lane), obs, n_obs, sx, sy, sr2);
}

// ---- children lists: links[i] = (first child, next sibling, previous sibling, frontier slot i), -1 = none; one
// 16-byte load gives all links of a node; the .w components together are the breadth-first queue of
// propagate_lists.  O(subtree) cost propagation. ----
static __device__ __forceinline__ void link_child(int4 *links, int p, int c) {
    const int f = links[p].x;
    links[c].y = f; links[c].z = -1;
    if (f >= 0) links[f].z = c;
    links[p].x = c;
}
static __device__ __forceinline__ void unlink_child(int4 *links, int p, int c) {
    const int4 lc = links[c];
    const int a = lc.z, b = lc.y;
    if (a >= 0) links[a].y = b; else links[p].x = b;
    if (b >= 0) links[b].z = a;
}

// propagate_cost_to_leaves (rrt_04:1379-1384) over the children lists: breadth-first from `root`, one lane per
// frontier node walking its child list; a node's cost depends only on its parent's, so any top-down order gives
// the reference's values.  The frontier lives in links[.].w, `tail` is a shared-memory int.
static __device__ __noinline__ void propagate_lists(int root, const double2 *xy, double *cost, int4 *links, int *tail,
                                                    int lane) {
    if (links[root].x < 0) return;
    if (lane == 0) { links[0].w = root; *tail = 1; }
    __syncwarp();
    // every queued node already has its final cost, so the queue is simply consumed 32 entries at a time
    for (int head = 0;;) {
        const int end = *tail;
        if (head >= end) break;
        const int k = head + lane;
        __syncwarp();
        if (k < end) {
            const int p = links[k].w;
            const double2 a = xy[p];
            const double cp = cost[p];
            for (int c = links[p].x; c >= 0;) {
                const int4 lc = links[c];
                const double2 b = xy[c];
                cost[c] = cp + crm_hypot(b.x - a.x, b.y - a.y);
                if (lc.x >= 0) links[atomicAdd(tail, 1)].w = c;
                c = lc.y;
            }
        }
        __syncwarp();
        head = end < head + 32 ? end : head + 32;
    }
}

// The same walk with cached edge lengths: elen[c] = hypot(c - parent(c)), exactly what calc_new_cost (rrt_04:1375-1377)
// recomputes at every visit -- positions only change when a re-parented node MOVES, so the value is kept from the
// moment the edge was made (choose_parent / rewire computed it then) and the serial hypot leaves the walk.  NaN = not
// known (edges made by an exact steer that stopped short, resumed trees): computed here and stored.  `root_moved`: the
// root's own position just changed, so the lengths of its child edges are stale.
static __device__ __noinline__ void propagate_lists_elen(int root, bool root_moved, const double2 *xy, double *cost, double *elen,
                                                         int4 *links, int *tail, int lane) {
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
            const double cp = cost[p];
            const bool stale = root_moved && p == root;
            for (int c = links[p].x; c >= 0;) {
                const int4 lc = links[c];
                double e = elen[c];
                if (stale || e != e) {
                    const double2 a = xy[p], b = xy[c];
                    e = crm_hypot(b.x - a.x, b.y - a.y);
                    elen[c] = e;
                }
                cost[c] = cp + e;
                if (lc.x >= 0) links[atomicAdd(tail, 1)].w = c;
                c = lc.y;
            }
        }
        __syncwarp();
        head = end < head + 32 ? end : head + 32;
    }
}

}  // namespace rrtk
