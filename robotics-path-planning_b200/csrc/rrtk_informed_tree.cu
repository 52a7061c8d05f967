// rrtk_informed_tree.cu -- Informed RRT* (rrt_07:1044-1108) on ONE large tree (BASELINE config 3: grow to
// ~10^6 nodes): the whole GPU works on a single query.  The warp-per-query kernel (rrtk_informed.cu) is the
// batched form; here every iteration's two O(n) list scans -- get_nearest_list_index (:1210-1214) and
// find_near_nodes (:1137-1143) -- are spread over all SMs, as are choose_parent (:1110-1135) and rewire
// (:1232-1246) over the near nodes.
//
// Execution model: a persistent grid of G co-resident CTAs (cooperative launch, one per SM) runs the
// iterations in lockstep.  Node i is OWNED by CTA (i / T) % G: only the owner scans it, evaluates it as a
// near candidate and rewires it, so cost[] / parent[] of a node are read and written by one CTA only.
// Per iteration there is ONE grid-wide reduction (a hand-rolled barrier over an L2 counter + one 32-byte
// record per CTA), because the near scan for the new node of iteration `it` and the nearest scan for the
// sample of iteration `it + 1` share one pass over the tree:
//   A  warp 0: new node = nearest + expand_dis * (cos, sin)(atan2(..))  (exact, crmath.h);  warp 1: sample it+1
//   B  CTA-wide: check_collision(nearest, theta, d) and the goal segment against all circles; obstacle cull
//   C  scan of the owned nodes: d^2 to the new node (near hits -> shared-memory list) and to sample it+1 (argmin)
//   D  owned hits: hypot, segment-vs-circle verdict, candidate cost  -> CTA partial (min cost, lowest index)
//   E  grid barrier + reduce of the G partial records (nearest of it+1, best parent of it, flags)
//   F  owner appends the node; every CTA rewires its own hits; goal bookkeeping (c_best, path snapshot)
// The sample of it+1 depends on c_best; when c_best changes in F (rare) the speculative nearest is redone.
//
// Exactness (results equal the sequential reference bit for bit):
//   * the `.index()` quirk of find_near_nodes maps a near node to the FIRST node with an equal d^2, so a node is
//     "shadowed" (never a parent candidate, never rewired) iff a lower-index node has the same d^2 to the new
//     node.  Identical positions are tracked with a per-node flag set at append time; equal d^2 between
//     different positions is detected with an L2 hash set of the hits' d^2 bit patterns and then resolved
//     exactly on a slow path (all-pairs over the global hit list);
//   * choose_parent returns the first minimum of the near list = the lowest unshadowed index among the
//     minimum-cost candidates; rewire entries are independent of each other (no propagation in rrt_07);
//   * segment verdicts of near edges use the reference's end point  a + (cos, sin)(theta) * d  only when the
//     verdict computed with the new node itself as end point lies within a tolerance band of an obstacle
//     boundary (the two end points differ by a few ulp); otherwise the cheap verdict is provably the same.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int TREE_T = 512;         // threads per CTA
constexpr int TREE_HCAP = 2048;     // near hits per CTA kept in shared memory (the rest spills to the workspace)
constexpr int TREE_OBS_CAP = 512;   // circles staged in shared memory
constexpr int TREE_TAB_BITS = 18;   // hash set of d^2 bit patterns: 2 tables x 2^18 x 8 B
constexpr unsigned long long TREE_EMPTY = ~0ull;
constexpr int TREE_PROBES = 64;

constexpr int FLAG_DUP_NEW = 1;     // the new node coincides with an existing node
constexpr int FLAG_EQ_D2 = 2;       // two hits at different positions share d^2 (or the hash set is crowded)
constexpr int HIT_FREE = 1 << 30;   // bits or-ed into a hit's node index
constexpr int HIT_SHADOW = 1 << 29;
constexpr int HIT_MASK = HIT_SHADOW - 1;

struct alignas(32) TreePartial {
    double nn_d2;    // nearest of sample it+1 among the owned nodes
    double cp_cost;  // best parent candidate among the owned hits
    int nn_idx, cp_idx;
    int flags, hits;
};

struct TreeWs {  // carved out of the caller's workspace
    unsigned long long *bar;
    TreePartial *partial;       // [2][G]
    unsigned long long *tab;    // [2][1 << TREE_TAB_BITS]
    uint8_t *dup;               // [node_cap]
    int *sp_idx, *sp_slot;      // [G][seg_cap] spill of the hit lists
    double *sp_d;               // [G][seg_cap]
    int *g_idx;                 // [G][seg_cap] slow path: all hits (index, d^2)
    double *g_d2;
    int *g_cnt;                 // [G]
    int seg_cap;
};

struct TreeArgs {
    rrtk_informed_tree_params p;
    const double4 *obstacles;
    const double2 *near_rr2, *free_s, *ball;
    double2 *xy;
    double *cost;
    int32_t *parent;
    double2 *path;
    rrtk_informed_tree_result *res;
    TreeWs ws;
};

// ---- L2-coherent accessors: the tree is written by other CTAs, so never read it through L1 ----
__device__ __forceinline__ double2 ld_xy(const double2 *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_f64(const double *p) { return __ldcg(p); }
__device__ __forceinline__ int ld_i32(const int *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double tdot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

// distance_squared_point_to_segment(v, w, p) (rrt_07:1249-1261), numpy's fma dot
__device__ __forceinline__ double seg_dd(double x1, double y1, double x2, double y2, double ox, double oy) {
    if (x1 == x2 && y1 == y2) return tdot2(ox - x1, oy - y1, ox - x1, oy - y1);
    double wx = x2 - x1, wy = y2 - y1;
    double l2 = tdot2(wx, wy, wx, wy);
    double t = tdot2(ox - x1, oy - y1, wx, wy) / l2;
    t = t < 1.0 ? t : 1.0;
    t = t > 0.0 ? t : 0.0;
    double px = x1 + t * wx, py = y1 + t * wy;
    return tdot2(ox - px, oy - py, ox - px, oy - py);
}

// the reference's verdict for the edge a -> new node (check_collision(node, theta, d), rrt_07:1271-1276)
__device__ __noinline__ bool edge_free_exact(double ax, double ay, double nx, double ny, double d,
                                             const double4 *s_obs, const int *s_cull, int ncull) {
    double s, c;
    (void)crm_atan2_sincos(ny - ay, nx - ax, &s, &c);
    const double ex = ax + c * d, ey = ay + s * d;
    for (int j = 0; j < ncull; j++) {
        const double4 o = s_obs[s_cull[j]];
        if (seg_dd(ax, ay, ex, ey, o.x, o.y) <= o.w) return false;
    }
    return true;
}

__device__ __forceinline__ bool edge_free(double ax, double ay, double nx, double ny, double d, double band_k,
                                          const double4 *s_obs, const int *s_cull, int ncull) {
    bool hit = false, unsure = !(d > 1e-9);
    for (int j = 0; j < ncull; j++) {
        const double4 o = s_obs[s_cull[j]];
        const double dd = seg_dd(ax, ay, nx, ny, o.x, o.y);
        const double band = band_k + 1e-12 * (dd + o.w);
        hit |= dd <= o.w;
        unsure |= fabs(dd - o.w) <= band;
    }
    if (unsure) return edge_free_exact(ax, ay, nx, ny, d, s_obs, s_cull, ncull);
    return !hit;
}

struct TreeSmem {
    double4 obs[TREE_OBS_CAP];
    int cull[TREE_OBS_CAP];
    int hit_idx[TREE_HCAP];
    int hit_slot[TREE_HCAP];
    double hit_d[TREE_HCAP];
    // block-reduce scratch
    double r_d[2][TREE_T / 32];
    int r_i[2][TREE_T / 32];
    int r_f[TREE_T / 32];
    int r_h[TREE_T / 32];
    // broadcast state
    double nx, ny, ex, ey, fx, fy, rx1, ry1, plen;
    double g_nn_d2, g_cp_cost;
    int g_nn_idx, g_cp_idx, g_flags, g_hits;
    int nhit, ncull, near_goal;
};

__device__ __forceinline__ void lexmin(double &v, int &i, double ov, int oi) {
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// CTA-wide reduction of (nn: d2, idx), (cp: cost, idx), flags (or), hits (sum); result valid in warp 0 lane 0
// and broadcast through shared memory by the caller.
__device__ __forceinline__ void block_reduce(TreeSmem &S, double &nd, int &ni, double &cd, int &ci, int &fl, int &hs) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        lexmin(nd, ni, __shfl_xor_sync(0xffffffffu, nd, off), __shfl_xor_sync(0xffffffffu, ni, off));
        lexmin(cd, ci, __shfl_xor_sync(0xffffffffu, cd, off), __shfl_xor_sync(0xffffffffu, ci, off));
        fl |= __shfl_xor_sync(0xffffffffu, fl, off);
        hs += __shfl_xor_sync(0xffffffffu, hs, off);
    }
    if (lane == 0) { S.r_d[0][w] = nd; S.r_i[0][w] = ni; S.r_d[1][w] = cd; S.r_i[1][w] = ci; S.r_f[w] = fl; S.r_h[w] = hs; }
    __syncthreads();
    if (w == 0) {
        const bool in = lane < TREE_T / 32;
        nd = in ? S.r_d[0][lane] : CUDART_INF; ni = in ? S.r_i[0][lane] : 0x7fffffff;
        cd = in ? S.r_d[1][lane] : CUDART_INF; ci = in ? S.r_i[1][lane] : 0x7fffffff;
        fl = in ? S.r_f[lane] : 0; hs = in ? S.r_h[lane] : 0;
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            lexmin(nd, ni, __shfl_xor_sync(0xffffffffu, nd, off), __shfl_xor_sync(0xffffffffu, ni, off));
            lexmin(cd, ci, __shfl_xor_sync(0xffffffffu, cd, off), __shfl_xor_sync(0xffffffffu, ci, off));
            fl |= __shfl_xor_sync(0xffffffffu, fl, off);
            hs += __shfl_xor_sync(0xffffffffu, hs, off);
        }
    }
}

// Grid-wide reduce: publish this CTA's partial, wait for all G, combine them; the result lands in S.g_*.
// `phase` counts the barriers executed so far (identical in every CTA).
__device__ __forceinline__ void grid_reduce(TreeSmem &S, const TreeWs &ws, unsigned long long &phase, int G,
                                            double nd, int ni, double cd, int ci, int fl, int hs) {
    block_reduce(S, nd, ni, cd, ci, fl, hs);
    TreePartial *slot = ws.partial + (phase & 1ull) * G;
    if (threadIdx.x == 0) {
        double4 a; int4 b;
        a.x = nd; a.y = cd; b.x = ni; b.y = ci; b.z = fl; b.w = hs;
        double2 *dp = reinterpret_cast<double2 *>(slot + blockIdx.x);
        __stcg(dp, make_double2(a.x, a.y));
        __stcg(reinterpret_cast<int4 *>(dp + 1), b);
        __threadfence();
        atomicAdd(ws.bar, 1ull);
        const unsigned long long target = (phase + 1ull) * (unsigned long long)G;
        while (ld_acquire(ws.bar) < target) { }
    }
    __syncthreads();
    nd = CUDART_INF; ni = 0x7fffffff; cd = CUDART_INF; ci = 0x7fffffff; fl = 0; hs = 0;
    if ((int)threadIdx.x < G) {
        const double2 *dp = reinterpret_cast<const double2 *>(slot + threadIdx.x);
        const double2 a = __ldcg(dp);
        const int4 b = __ldcg(reinterpret_cast<const int4 *>(dp + 1));
        nd = a.x; cd = a.y; ni = b.x; ci = b.y; fl = b.z; hs = b.w;
    }
    block_reduce(S, nd, ni, cd, ci, fl, hs);
    if (threadIdx.x == 0) {
        S.g_nn_d2 = nd; S.g_nn_idx = ni; S.g_cp_cost = cd; S.g_cp_idx = ci; S.g_flags = fl; S.g_hits = hs;
    }
    __syncthreads();
    phase++;
}

// plain grid barrier (makes the rewires of this iteration visible before the goal walk)
__device__ __forceinline__ void grid_barrier(const TreeWs &ws, unsigned long long &phase, int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ws.bar, 1ull);
        const unsigned long long target = (phase + 1ull) * (unsigned long long)G;
        while (ld_acquire(ws.bar) < target) { }
    }
    __syncthreads();
    phase++;
}

// informed_sample (rrt_07:1145-1159) for iteration `it` given c_best
__device__ __forceinline__ void draw_sample(const TreeArgs &A, int it, double c_best, double c_min, double xc, double yc,
                                            double &rx, double &ry) {
    if (c_best < CUDART_INF) {
        const double r0 = c_best / 2.0;
        const double r1 = sqrt(c_best * c_best - c_min * c_min) / 2.0;
        const double2 ab = __ldg(A.ball + it);
        double a = ab.x, b = ab.y;
        if (b < a) { double t = a; a = b; b = t; }
        const double ang = 2 * 3.141592653589793 * a / b;
        crm_dd sd, cd;
        crm_sincos_dd(ang, &sd, &cd);
        const double bx = b * cd.hi, by = b * (ang == 0.0 ? ang : sd.hi);
        const double m00 = A.p.rot[0] * r0, m01 = A.p.rot[1] * r1, m10 = A.p.rot[2] * r0, m11 = A.p.rot[3] * r1;
        rx = fma(m00, bx, m01 * by) + xc;
        ry = fma(m10, bx, m11 * by) + yc;
    } else {
        const double2 f = __ldg(A.free_s + it);
        rx = f.x; ry = f.y;
    }
}

__global__ void __launch_bounds__(TREE_T, 1) informed_tree_kernel(const TreeArgs A) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    TreeSmem &S = *reinterpret_cast<TreeSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const TreeWs &ws = A.ws;
    const double INF = CUDART_INF;
    const double ed = A.p.expand_dis;
    const double sx = A.p.start_goal[0], sy = A.p.start_goal[1], gx = A.p.start_goal[2], gy = A.p.start_goal[3];
    const int n_obs = A.p.n_obs;
    const double band_k = 1e-12 * A.p.coord_bound * A.p.coord_bound;
    const long long stride = (long long)G * TREE_T;
    const long long seg = (long long)cta * ws.seg_cap;

    for (int j = tid; j < n_obs; j += TREE_T) S.obs[j] = A.obstacles[j];
    if (cta == 0 && tid == 0) { A.xy[0] = make_double2(sx, sy); A.cost[0] = 0.0; A.parent[0] = -1; }
    __syncthreads();

    unsigned long long phase = 0;
    grid_barrier(ws, phase, G);  // node 0 visible everywhere

    int n = 1, status = RRTK_Q_OK, plen_best = 0, it = 0;
    int ni = 0;                       // nearest node of the current sample
    int last_new = -1;                // node appended in the previous iteration (its xy may not be visible yet)
    double last_x = 0.0, last_y = 0.0;
    double c_best = INF;
    const double c_min = crm_hypot(sx - gx, sy - gy);
    const double xc = (sx + gx) / 2.0, yc = (sy + gy) / 2.0;
    double rx = 0.0, ry = 0.0;
    if (A.p.max_iter > 0) draw_sample(A, 0, c_best, c_min, xc, yc, rx, ry);
    long long total_hits = 0;
    int n_slow = 0, n_goal = 0, n_redo = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};  // CTA 0's clock per phase: A, B, C, D, E (barrier + reduce), F
    long long t0 = clock64();
#define TREE_TICK(k) do { long long t1 = clock64(); cyc[k] += t1 - t0; t0 = t1; } while (0)

    for (; it < A.p.max_iter; it++) {
        const bool have_next = it + 1 < A.p.max_iter;
        // ---- A: new node (warp 0) and the speculative sample of it+1 (warp 1) ----
        if (warp == 0) {
            double2 from;
            if (ni == last_new) from = make_double2(last_x, last_y);
            else from = ld_xy(A.xy + ni);
            double st, ct;
            (void)crm_atan2_sincos(ry - from.y, rx - from.x, &st, &ct);
            const double nx = from.x + ed * ct, ny = from.y + ed * st;   // get_new_node (rrt_07:1216-1224)
            const double d0 = crm_hypot(from.x - nx, from.y - ny);       // line_cost (:1205-1207)
            const double dg = crm_hypot(nx - gx, ny - gy);               // is_near_goal (:1226-1230)
            if (lane == 0) {
                S.nx = nx; S.ny = ny; S.fx = from.x; S.fy = from.y;
                S.ex = from.x + ct * d0; S.ey = from.y + st * d0;        // check_collision end point (:1273-1274)
                S.near_goal = dg < ed;
                S.nhit = 0; S.ncull = 0;
            }
        } else if (warp == 1) {
            double a = 0.0, b = 0.0;
            if (have_next) draw_sample(A, it + 1, c_best, c_min, xc, yc, a, b);
            if (lane == 0) { S.rx1 = a; S.ry1 = b; }
        }
        __syncthreads();
        TREE_TICK(0);
        const double nx = S.nx, ny = S.ny;
        double rx1 = S.rx1, ry1 = S.ry1;
        const bool near_goal = S.near_goal != 0;
        const double2 rr2 = __ldg(A.near_rr2 + n);  // (r, r ** 2) for n_node = n (rrt_07:1138-1139)

        // ---- B: check_collision(nearest, theta, d), goal segment, obstacle cull around the new node ----
        int hit_edge = 0, hit_goal = 0;
        for (int j = tid; j < n_obs; j += TREE_T) {
            const double4 o = S.obs[j];
            hit_edge |= seg_dd(S.fx, S.fy, S.ex, S.ey, o.x, o.y) <= o.w;
            if (near_goal) hit_goal |= seg_dd(nx, ny, gx, gy, o.x, o.y) <= o.w;
            const double dx = o.x - nx, dy = o.y - ny;
            const double lim = (rr2.x + o.z) * (1.0 + 1e-9) + 1e-9;
            if (dx * dx + dy * dy <= lim * lim) S.cull[atomicAdd(&S.ncull, 1)] = j;
        }
        const int blocked = __syncthreads_or(hit_edge | (hit_goal << 1));
        const bool accept = !(blocked & 1);
        const bool goal_event = accept && near_goal && !(blocked & 2);
        const int ncull = S.ncull;
        TREE_TICK(1);
        if (accept && n >= A.p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; break; }

        // ---- C: one pass over the owned nodes: near hits of the new node, nearest of sample it+1 ----
        const double r2 = accept ? rr2.y : -1.0;
        double bd = INF;
        int bi = 0x7fffffff;
        if (accept || have_next) {
            for (long long base = (long long)cta * TREE_T + tid; base < n; base += 4 * stride) {
                double2 a[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const long long i = base + u * stride;
                    ok[u] = i < n;
                    a[u] = ok[u] ? ld_xy(A.xy + i) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (!ok[u]) continue;
                    const int i = (int)(base + u * stride);
                    const double ax = a[u].x - nx, ay = a[u].y - ny;
                    if (ax * ax + ay * ay <= r2) {
                        const int pos = atomicAdd(&S.nhit, 1);
                        if (pos < TREE_HCAP) S.hit_idx[pos] = i;
                        else __stcg(ws.sp_idx + seg + (pos - TREE_HCAP), i);
                    }
                    const double bx = a[u].x - rx1, by = a[u].y - ry1;
                    const double e2 = bx * bx + by * by;
                    if (e2 < bd) { bd = e2; bi = i; }
                }
            }
        }
        __syncthreads();
        const int H = S.nhit;
        TREE_TICK(2);

        // ---- D: choose_parent candidates among the owned hits (rrt_07:1110-1135) ----
        double cc = INF;
        int ci = 0x7fffffff, flags = 0;
        unsigned long long *tab = ws.tab + ((unsigned long long)(it & 1) << TREE_TAB_BITS);
        for (int e = tid; e < H; e += TREE_T) {
            const int i = e < TREE_HCAP ? S.hit_idx[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
            int tag = i, slot = -1;
            double d = 0.0;
            if (__ldcg(ws.dup + i)) {
                tag |= HIT_SHADOW;
            } else {
                const double2 a = ld_xy(A.xy + i);
                const double ax = a.x - nx, ay = a.y - ny;
                const double d2 = ax * ax + ay * ay;
                if (d2 == 0.0) flags |= FLAG_DUP_NEW;
                // hash set of d^2 bit patterns: a repeated key means two different positions at equal d^2
                const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
                unsigned h = (unsigned)(splitmix64(key) >> (64 - TREE_TAB_BITS));
                int probe = 0;
                for (; probe < TREE_PROBES; probe++) {
                    const unsigned long long old = atomicCAS(tab + h, TREE_EMPTY, key);
                    if (old == TREE_EMPTY) { slot = (int)h; break; }
                    if (old == key) { flags |= FLAG_EQ_D2; break; }
                    h = (h + 1) & ((1u << TREE_TAB_BITS) - 1u);
                }
                if (probe == TREE_PROBES) flags |= FLAG_EQ_D2;
                d = crm_hypot(nx - a.x, ny - a.y);
                if (edge_free(a.x, a.y, nx, ny, d, band_k, S.obs, S.cull, ncull)) {
                    tag |= HIT_FREE;
                    lexmin(cc, ci, ld_f64(A.cost + i) + d, i);
                }
            }
            if (e < TREE_HCAP) { S.hit_idx[e] = tag; S.hit_slot[e] = slot; S.hit_d[e] = d; }
            else {
                __stcg(ws.sp_idx + seg + (e - TREE_HCAP), tag);
                __stcg(ws.sp_slot + seg + (e - TREE_HCAP), slot);
                __stcg(ws.sp_d + seg + (e - TREE_HCAP), d);
            }
        }

        // ---- E: grid-wide reduce ----
        __syncthreads();
        TREE_TICK(3);
        grid_reduce(S, ws, phase, G, bd, bi, cc, ci, flags, tid == 0 ? H : 0);
        TREE_TICK(4);
        flags = S.g_flags;
        total_hits += S.g_hits;
        int nn1 = S.g_nn_idx;
        const double nn1_d2 = S.g_nn_d2;

        if (!accept) {  // collision: nothing is added (rrt_07:1080-1082)
            if (have_next) { ni = nn1; rx = rx1; ry = ry1; }
            continue;
        }
        struct Tick { long long *c, *t; __device__ ~Tick() { long long t1 = clock64(); c[5] += t1 - *t; *t = t1; } } tickF{cyc, &t0};

        // ---- slow path: equal d^2 at different positions -> exact shadow flags from the global hit list ----
        if (flags & FLAG_EQ_D2) {
            n_slow++;
            for (int e = tid; e < H; e += TREE_T) {
                const int tag = e < TREE_HCAP ? S.hit_idx[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
                const int i = tag & HIT_MASK;
                const double2 a = ld_xy(A.xy + i);
                const double ax = a.x - nx, ay = a.y - ny;
                __stcg(ws.g_idx + seg + e, i);
                __stcg(ws.g_d2 + seg + e, ax * ax + ay * ay);
            }
            if (tid == 0) __stcg(ws.g_cnt + cta, H);
            grid_barrier(ws, phase, G);
            cc = INF; ci = 0x7fffffff;
            for (int e = tid; e < H; e += TREE_T) {
                int tag = e < TREE_HCAP ? S.hit_idx[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
                const int i = tag & HIT_MASK;
                const double d2 = ld_f64(ws.g_d2 + seg + e);
                bool shadow = false;
                for (int c2 = 0; c2 < G && !shadow; c2++) {
                    const int cnt = ld_i32(ws.g_cnt + c2);
                    const long long s2 = (long long)c2 * ws.seg_cap;
                    for (int k = 0; k < cnt; k++)
                        if (ld_i32(ws.g_idx + s2 + k) < i && ld_f64(ws.g_d2 + s2 + k) == d2) { shadow = true; break; }
                }
                if (shadow) tag |= HIT_SHADOW;
                if (e < TREE_HCAP) S.hit_idx[e] = tag; else __stcg(ws.sp_idx + seg + (e - TREE_HCAP), tag);
                if (!(tag & HIT_SHADOW) && (tag & HIT_FREE)) {
                    const double d = e < TREE_HCAP ? S.hit_d[e] : ld_f64(ws.sp_d + seg + (e - TREE_HCAP));
                    lexmin(cc, ci, ld_f64(A.cost + i) + d, i);
                }
            }
            grid_reduce(S, ws, phase, G, INF, 0x7fffffff, cc, ci, 0, 0);
        }

        // ---- F: parent choice, append, rewire (rrt_07:1232-1246) ----
        double ncost;
        int npar;
        if (S.g_cp_idx != 0x7fffffff) { ncost = S.g_cp_cost; npar = S.g_cp_idx; }
        else { ncost = ld_f64(A.cost + ni) + ed; npar = ni; }
        const int newi = n;
        if (cta == (newi / TREE_T) % G && tid == 0) {
            __stcg(A.xy + newi, make_double2(nx, ny));
            __stcg(A.cost + newi, ncost);
            __stcg(A.parent + newi, npar);
            __stcg(ws.dup + newi, (uint8_t)((flags & FLAG_DUP_NEW) ? 1 : 0));
        }
        for (int e = tid; e < H; e += TREE_T) {
            const bool in_s = e < TREE_HCAP;
            const int tag = in_s ? S.hit_idx[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
            const int slot = in_s ? S.hit_slot[e] : ld_i32(ws.sp_slot + seg + (e - TREE_HCAP));
            if (slot >= 0) tab[slot] = TREE_EMPTY;
            if ((tag & HIT_SHADOW) || !(tag & HIT_FREE)) continue;
            const int i = tag & HIT_MASK;
            const double sc = ncost + (in_s ? S.hit_d[e] : ld_f64(ws.sp_d + seg + (e - TREE_HCAP)));
            if (ld_f64(A.cost + i) > sc) { __stcg(A.parent + i, newi); __stcg(A.cost + i, sc); }
        }
        n++;
        last_new = newi; last_x = nx; last_y = ny;
        if (have_next) {  // the new node joins the nearest candidates of sample it+1 (highest index: loses ties)
            const double bx = nx - rx1, by = ny - ry1;
            if (bx * bx + by * by < nn1_d2) nn1 = newi;
        }

        // ---- goal bookkeeping (rrt_07:1094-1103) ----
        bool c_changed = false;
        if (goal_event) {
            n_goal++;
            grid_barrier(ws, phase, G);  // this iteration's rewires are now visible
            if (tid == 0) {
                double plen = 0.0, qx = gx, qy = gy;
                int k = newi;
                for (int guard = 0; guard <= A.p.node_cap; guard++) {
                    const int pk = k == newi ? npar : ld_i32(A.parent + k);
                    if (pk < 0) break;
                    const double2 a = k == newi ? make_double2(nx, ny) : ld_xy(A.xy + k);
                    plen += crm_hypot(a.x - qx, a.y - qy);
                    qx = a.x; qy = a.y;
                    k = pk;
                }
                plen += crm_hypot(sx - qx, sy - qy);
                S.plen = plen;
            }
            __syncthreads();
            const double plen = S.plen;
            if (plen < c_best) {
                c_best = plen;
                c_changed = true;
                if (cta == 0 && tid == 0) {  // snapshot: goal, new node, ..., first child of the root, start
                    int w = 0;
                    if (w < A.p.path_cap) A.path[w] = make_double2(gx, gy);
                    w++;
                    int k = newi;
                    for (int guard = 0; guard <= A.p.node_cap; guard++, w++) {
                        const int pk = k == newi ? npar : ld_i32(A.parent + k);
                        if (pk < 0) break;
                        if (w < A.p.path_cap) A.path[w] = k == newi ? make_double2(nx, ny) : ld_xy(A.xy + k);
                        k = pk;
                    }
                    if (w < A.p.path_cap) A.path[w] = make_double2(sx, sy);
                    w++;
                    plen_best = w;
                }
            }
            __syncthreads();
        }

        if (!have_next) continue;
        if (c_changed) {  // sample it+1 depends on c_best: redo its nearest search over the whole tree
            n_redo++;
            if (warp == 0) {
                double a, b;
                draw_sample(A, it + 1, c_best, c_min, xc, yc, a, b);
                if (lane == 0) { S.rx1 = a; S.ry1 = b; }
            }
            __syncthreads();
            rx1 = S.rx1; ry1 = S.ry1;
            bd = INF; bi = 0x7fffffff;
            for (long long i = (long long)cta * TREE_T + tid; i < n; i += stride) {
                const double2 a = ld_xy(A.xy + i);
                const double bx = a.x - rx1, by = a.y - ry1;
                const double e2 = bx * bx + by * by;
                if (e2 < bd) { bd = e2; bi = (int)i; }
            }
            grid_reduce(S, ws, phase, G, bd, bi, INF, 0x7fffffff, 0, 0);
            nn1 = S.g_nn_idx;
        }
        ni = nn1; rx = rx1; ry = ry1;
    }

    if (cta == 0 && tid == 0) {
        rrtk_informed_tree_result r;
        r.n_nodes = n; r.path_len = plen_best; r.status = status | (plen_best > A.p.path_cap ? RRTK_Q_PATH_OVERFLOW : 0);
        r.iters_done = it; r.c_best = c_best; r.total_hits = total_hits; r.slow_paths = n_slow;
        r.goal_events = n_goal; r.resamples = n_redo; r.grid = G;
        for (int k = 0; k < 6; k++) r.cycles[k] = cyc[k];
        *A.res = r;
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace layout for a grid of G CTAs
static size_t carve(TreeWs &ws, char *base, int node_cap, int G) {
    const long long chunks = ((long long)node_cap + TREE_T - 1) / TREE_T;
    const int seg_cap = (int)((chunks + G - 1) / G) * TREE_T;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return base ? base + o : nullptr; };
    ws.bar = (unsigned long long *)take(256);
    ws.partial = (TreePartial *)take(sizeof(TreePartial) * 2 * G);
    ws.tab = (unsigned long long *)take(sizeof(unsigned long long) * 2 * (1ull << TREE_TAB_BITS));
    ws.dup = (uint8_t *)take((size_t)node_cap);
    ws.sp_idx = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.sp_slot = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.sp_d = (double *)take(sizeof(double) * (size_t)G * seg_cap);
    ws.g_idx = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.g_d2 = (double *)take(sizeof(double) * (size_t)G * seg_cap);
    ws.g_cnt = (int *)take(sizeof(int) * G);
    ws.seg_cap = seg_cap;
    return off;
}

static int tree_grid(int want, int *grid_out) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDevice");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaFuncSetAttribute(informed_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeSmem));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(smem)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, informed_tree_kernel, TREE_T, sizeof(TreeSmem));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) return set_error(RRTK_ERR_CUDA, "informed_tree_kernel does not fit on an SM");
    int g = sms;  // one CTA per SM
    if (want > 0 && want < g) g = want;
    if (g > TREE_T) g = TREE_T;
    *grid_out = g;
    return RRTK_OK;
}

int informed_tree_workspace_bytes(int node_cap, int grid, size_t *bytes) {
    int G = 0;
    int rc = tree_grid(grid, &G);
    if (rc) return rc;
    TreeWs ws;
    *bytes = carve(ws, nullptr, node_cap, G);
    return RRTK_OK;
}

int launch_informed_tree(const rrtk_informed_tree_params &p, const double *obstacles, const double *near_rr2,
                         const double *free_s, const double *ball, double *xy, double *cost, int32_t *parent,
                         double *path, rrtk_informed_tree_result *res, void *workspace, size_t workspace_bytes,
                         cudaStream_t s) {
    int G = 0;
    int rc = tree_grid(p.grid, &G);
    if (rc) return rc;
    TreeArgs A;
    A.p = p;
    A.obstacles = reinterpret_cast<const double4 *>(obstacles);
    A.near_rr2 = reinterpret_cast<const double2 *>(near_rr2);
    A.free_s = reinterpret_cast<const double2 *>(free_s);
    A.ball = reinterpret_cast<const double2 *>(ball);
    A.xy = reinterpret_cast<double2 *>(xy);
    A.cost = cost; A.parent = parent;
    A.path = reinterpret_cast<double2 *>(path);
    A.res = res;
    const size_t need = carve(A.ws, (char *)workspace, p.node_cap, G);
    if (need > workspace_bytes) return set_error(RRTK_ERR_INVALID, "workspace too small (rrtk_informed_tree_workspace_bytes)");
    cudaError_t e = cudaMemsetAsync(A.ws.bar, 0, 256, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(A.ws.tab, 0xff, sizeof(unsigned long long) * 2 * (1ull << TREE_TAB_BITS), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(A.ws.dup, 0, (size_t)p.node_cap, s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(workspace)");
    void *args[] = {(void *)&A};
    e = cudaLaunchCooperativeKernel((const void *)informed_tree_kernel, dim3(G), dim3(TREE_T), args, sizeof(TreeSmem), s);
    if (e != cudaSuccess) return set_cuda_error(e, "informed_tree_kernel cooperative launch");
    return RRTK_OK;
}

}  // namespace rrtk
