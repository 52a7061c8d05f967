// rrtk_informed_tree.cu -- Informed RRT* (rrt_07:1044-1108) on ONE large tree (BASELINE config 3: grow to
// ~10^6 nodes): the whole GPU works on a single query.  The warp-per-query kernel (rrtk_informed.cu) is the
// batched form; here every iteration's two O(n) list scans -- get_nearest_list_index (:1210-1214) and
// find_near_nodes (:1137-1143) -- are spread over all SMs, as are choose_parent (:1110-1135) and rewire
// (:1232-1246) over the near nodes.
//
// Execution model: a persistent grid of G co-resident CTAs (cooperative launch, one per SM) runs the
// iterations in lockstep.  Node i is OWNED by CTA (i / T) % G: only the owner scans it, evaluates it as a
// near candidate and rewires it, so cost[] / parent[] of a node are read and written by one CTA only.
// One iteration = one pass + one grid-wide exchange:
//   cull   obstacles that can touch an edge into the new node                      (all threads)
//   scan   owned nodes: d^2 to the new node of iteration `it` (near hits -> shared-memory list) and to the
//          sample of iteration `it + 1` (argmin, lowest index)                       (all threads)
//   cand   the CTA's nearest candidate for sample it+1 is extended SPECULATIVELY: exact atan2/cos/sin, new
//          position, check_collision of that edge and of the goal segment          (one warp)
//   hits   choose_parent candidates of the owned hits: hypot, segment verdict, cost (the other warps)
//   xchg   every CTA publishes one 64-byte record of 16-byte pieces (value, index, tag) and
//          polls the G records: nearest of it+1 WITH its already-extended node, best parent of it, flags.
//          Release = fence + one L2 counter increment per CTA; one thread per CTA polls the counter (relaxed) and
//          finishes with an acquire load before the records are read (see exchange()).
//   apply  owner appends the node; every CTA rewires its own hits; goal bookkeeping (c_best, path snapshot)
// so the serial leaf math of iteration it+1 overlaps the candidate evaluation of iteration it.  The sample of
// it+1 depends on c_best and the new node of `it` may itself be the nearest of it+1: both cases are detected
// and redone exactly (a nearest-only pass, or a re-extension from the new node).
//
// Exactness (results equal the sequential reference bit for bit):
//   * the `.index()` quirk of find_near_nodes maps a near node to the FIRST node with an equal d^2, so a node is
//     "shadowed" (never a parent candidate, never rewired) iff a lower-index node has the same d^2 to the new
//     node.  Coincident nodes are flagged at append time (owner-local bitset) and skipped; equal d^2 between
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
constexpr int TREE_NW = TREE_T / 32;
constexpr int TREE_CAND_WARP = TREE_NW - 1;        // extends the speculative next node
constexpr int TREE_HIT_THREADS = TREE_T - 32;      // evaluate the near hits meanwhile
constexpr int TREE_HCAP = 1024;     // near hits per CTA kept in shared memory (the rest spills to the workspace)
constexpr int TREE_OBS_CAP = 512;   // circles staged in shared memory
constexpr int TREE_TAB_BITS = 18;   // hash set of d^2 bit patterns: 2 tables x 2^18 x 8 B
constexpr unsigned long long TREE_EMPTY = ~0ull;
constexpr int TREE_PROBES = 64;
constexpr int TREE_UNROLL = 2;      // node chunks in flight per thread in the scan (code size vs latency hiding)
constexpr int TREE_REC_PIECES = 8;
  // 128 bytes per record, 5 pieces used

constexpr int FLAG_DUP_NEW = 1;     // the new node coincides with an existing node
constexpr int FLAG_EQ_D2 = 2;       // two hits at different positions share d^2 (or the hash set is crowded)
constexpr int CF_BLOCKED = 1;       // candidate flags: edge nearest -> new node hits a circle
constexpr int CF_NEAR_GOAL = 2;     //                  new node within expand_dis of the goal
constexpr int CF_GOAL_BLOCKED = 4;  //                  segment new node -> goal hits a circle
constexpr int HIT_FREE = 1 << 30;   // bits or-ed into a hit's node index
constexpr int HIT_SHADOW = 1 << 29;
constexpr int HIT_MASK = HIT_SHADOW - 1;
constexpr int NO_IDX = 0x7fffffff;

struct TreeWs {  // carved out of the caller's workspace
    unsigned long long *bar;    // counter barrier (rare paths)
    uint4 *rec;                 // [2][G][TREE_REC_PIECES] exchange records
    unsigned long long *tab;    // [2][1 << TREE_TAB_BITS]
    int *sp_idx, *sp_slot;      // [G][seg_cap] spill of the hit lists
    double *sp_d, *sp_c;        // [G][seg_cap]
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
__device__ __forceinline__ uint4 ld_volatile16(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile16(uint4 *p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 piece(double v, int i, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return make_uint4((unsigned)b, (unsigned)(b >> 32), (unsigned)i, tag);
}
__device__ __forceinline__ double piece_f64(uint4 p) {
    return __longlong_as_double((long long)(((unsigned long long)p.y << 32) | p.x));
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
    int hit_tag[TREE_HCAP];
    int hit_slot[TREE_HCAP];
    double hit_x[TREE_HCAP], hit_y[TREE_HCAP], hit_d[TREE_HCAP], hit_c[TREE_HCAP];
    // per-warp partials of the pass
    double w_nd[TREE_NW], w_nx[TREE_NW], w_ny[TREE_NW], w_cc[TREE_NW];
    int w_ni[TREE_NW], w_ci[TREE_NW], w_fl[TREE_NW], w_hs[TREE_NW];
    // the CTA's candidate for the next sample
    double c_d2, c_nx, c_ny;
    int c_idx, c_cf;
    // per-warp partials of the exchange
    double q_nd[TREE_NW], q_nx[TREE_NW], q_ny[TREE_NW], q_cc[TREE_NW];
    int q_ni[TREE_NW], q_cf[TREE_NW], q_ci[TREE_NW], q_fl[TREE_NW], q_hs[TREE_NW];
    double s2x, s2y, s1x, s1y, plen;
    int nhit, ncull;
    unsigned dupbits[1];  // [seg_cap / 32] coincident-node flags of the owned nodes (dynamic tail)
};

struct Winner {   // result of one exchange, identical in every thread of every CTA
    double nn_d2, nx, ny, cp_cost;
    int nn_idx, cf, cp_idx, flags, hits;
};

__device__ __forceinline__ void lexmin(double &v, int &i, double ov, int oi) {
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// counter barrier for the rare paths (goal walk, equal-d^2 resolution)
__device__ __forceinline__ void grid_barrier(const TreeWs &ws, unsigned long long &bphase, int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ws.bar, 1ull);
        const unsigned long long target = (bphase + 1ull) * (unsigned long long)G;
        while (ld_acquire(ws.bar) < target) { }
    }
    __syncthreads();
    bphase++;
}

// warp-wide argmin of (non-negative double, int) with lowest-index ties, via integer redux: 3 REDUX + 2 compares
// instead of 5 shuffle rounds (code size matters here).  Every lane gets the winning lane id.
__device__ __forceinline__ int warp_argmin_lane(double v, int i) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);  // v >= 0 or +inf: bits are order-preserving
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool is_min = hi == mhi && lo == mlo;
    const unsigned mi = __reduce_min_sync(0xffffffffu, is_min ? (unsigned)i : 0xffffffffu);
    return __ffs(__ballot_sync(0xffffffffu, is_min && (unsigned)i == mi)) - 1;
}

// Grid-wide exchange.  In: the per-warp partials S.w_cc / w_ci / w_fl / w_hs and the candidate S.c_* (all written
// before the caller's last __syncthreads).  Out: the combined result in every thread.
// Each CTA stores one 64-byte record, then a release (MEMBAR + L2 counter increment); one thread per CTA polls the
// counter; G threads read the records back and reduce.  Measured on B200 (tools/micro/sync_costs.cu): this counter
// form costs ~2200 cycles at G = 148, an all-to-all poll of self-tagged records ~2000-4900 depending on the piece
// count, and the latter collapses under the G^2 polling traffic once real work runs beside it.
// (rec / bar / seq by value: a reference to the workspace struct or to the caller's counter is a round trip through local memory)
__device__ __noinline__ Winner exchange(TreeSmem &S, uint4 *rec, unsigned long long *bar, const unsigned seq, int G) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint4 *recs = rec + (size_t)(seq & 1u) * G * TREE_REC_PIECES;
    if (warp == 0) {
        const bool in = lane < TREE_NW;
        const double cc0 = in ? S.w_cc[lane] : CUDART_INF;
        const int ci0 = in ? S.w_ci[lane] : NO_IDX;
        const int src = warp_argmin_lane(cc0, ci0);
        const double cc = __shfl_sync(0xffffffffu, cc0, src);
        const int ci = __shfl_sync(0xffffffffu, ci0, src);
        const int fl = (int)__reduce_or_sync(0xffffffffu, in ? (unsigned)S.w_fl[lane] : 0u);
        const int hs = (int)__reduce_add_sync(0xffffffffu, in ? (unsigned)S.w_hs[lane] : 0u);
        // lanes 0..3 store one 16-byte piece each (selected without branches)
        const double pv = lane == 0 ? S.c_d2 : lane == 1 ? cc : lane == 2 ? S.c_nx : S.c_ny;
        const int pi = lane == 0 ? S.c_idx : lane == 1 ? ci : lane == 2 ? S.c_cf : 0;
        const int pt = lane == 0 ? fl : lane == 1 ? hs : 0;
        if (lane < 4) __stcg(recs + (size_t)blockIdx.x * TREE_REC_PIECES + lane, piece(pv, pi, (unsigned)pt));
        __syncwarp();
        if (lane == 0) {
            // release: the record and this CTA's tree writes (ordered before by __syncthreads) precede the count
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            atomicAdd(bar + 8, 1ull);
            const unsigned long long target = (unsigned long long)(seq + 1u) * (unsigned long long)G;
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar + 8) : "memory"); } while (v < target);
            (void)ld_acquire(bar + 8);   // acquire: pairs with the releasing fence + atomicAdd of every CTA, so the record / tree reads below are ordered after them
        }
    }
    __syncthreads();
    double nd = CUDART_INF, nx = 0.0, ny = 0.0, cc = CUDART_INF;
    int ni = NO_IDX, cf = 0, ci = NO_IDX, fl = 0, hs = 0;
    if (tid < G) {   // L2 loads: the records were written before the counter reached its target
        const uint4 *r = recs + (size_t)tid * TREE_REC_PIECES;
        const uint4 p0 = __ldcg(r), p1 = __ldcg(r + 1), p2 = __ldcg(r + 2), p3 = __ldcg(r + 3);
        nd = piece_f64(p0); ni = (int)p0.z; fl = (int)p0.w;
        cc = piece_f64(p1); ci = (int)p1.z; hs = (int)p1.w;
        nx = piece_f64(p2); cf = (int)p2.z;
        ny = piece_f64(p3);
    }
    const int nwq = (G + 31) >> 5;
    if (warp < nwq) {
        const int sn = warp_argmin_lane(nd, ni), sc = warp_argmin_lane(cc, ci);
        nd = __shfl_sync(0xffffffffu, nd, sn); ni = __shfl_sync(0xffffffffu, ni, sn);
        nx = __shfl_sync(0xffffffffu, nx, sn); ny = __shfl_sync(0xffffffffu, ny, sn);
        cf = __shfl_sync(0xffffffffu, cf, sn);
        cc = __shfl_sync(0xffffffffu, cc, sc); ci = __shfl_sync(0xffffffffu, ci, sc);
        fl = (int)__reduce_or_sync(0xffffffffu, (unsigned)fl);
        hs = (int)__reduce_add_sync(0xffffffffu, (unsigned)hs);
        if (lane == 0) {
            S.q_nd[warp] = nd; S.q_ni[warp] = ni; S.q_nx[warp] = nx; S.q_ny[warp] = ny; S.q_cf[warp] = cf;
            S.q_cc[warp] = cc; S.q_ci[warp] = ci; S.q_fl[warp] = fl; S.q_hs[warp] = hs;
        }
    }
    __syncthreads();
    Winner W;
    W.nn_d2 = CUDART_INF; W.nn_idx = NO_IDX; W.nx = 0.0; W.ny = 0.0; W.cf = 0;
    W.cp_cost = CUDART_INF; W.cp_idx = NO_IDX; W.flags = 0; W.hits = 0;
    for (int w = 0; w < nwq; w++) {
        const double ond = S.q_nd[w];
        const int oni = S.q_ni[w];
        if (ond < W.nn_d2 || (ond == W.nn_d2 && oni < W.nn_idx)) {
            W.nn_d2 = ond; W.nn_idx = oni; W.nx = S.q_nx[w]; W.ny = S.q_ny[w]; W.cf = S.q_cf[w];
        }
        lexmin(W.cp_cost, W.cp_idx, S.q_cc[w], S.q_ci[w]);
        W.flags |= S.q_fl[w];
        W.hits += S.q_hs[w];
    }
    return W;
}

// informed_sample (rrt_07:1145-1159) for iteration `it` given c_best
__device__ __noinline__ double2 draw_sample_v(const double2 *free_s, const double2 *ball, double r00, double r01, double r10,
                                              double r11, int it, double c_best, double c_min, double xc, double yc) {
    double rx, ry;
    if (c_best < CUDART_INF) {
        const double r0 = c_best / 2.0;
        const double r1 = sqrt(c_best * c_best - c_min * c_min) / 2.0;
        const double2 ab = __ldg(ball + it);
        double a = ab.x, b = ab.y;
        if (b < a) { double t = a; a = b; b = t; }
        const double ang = 2 * 3.141592653589793 * a / b;
        double sn, cs;
        crm_sincos(ang, &sn, &cs);
        const double bx = b * cs, by = b * (ang == 0.0 ? ang : sn);
        const double m00 = r00 * r0, m01 = r01 * r1, m10 = r10 * r0, m11 = r11 * r1;
        rx = fma(m00, bx, m01 * by) + xc;
        ry = fma(m10, bx, m11 * by) + yc;
    } else {
        const double2 f = __ldg(free_s + it);
        rx = f.x; ry = f.y;
    }
    return make_double2(rx, ry);
}
// (by value across the call; the kernel argument block stays in the constant bank)
#define draw_sample(A, it, c_best, c_min, xc, yc, rx, ry) do { \
        const double2 ds_ = draw_sample_v((A).free_s, (A).ball, (A).p.rot[0], (A).p.rot[1], (A).p.rot[2], (A).p.rot[3], (it), (c_best), (c_min), (xc), (yc)); \
        (rx) = ds_.x; (ry) = ds_.y; } while (0)

// One warp extends `from` towards the sample: get_new_node (rrt_07:1216-1224), line_cost, check_collision(nearest,
// theta, d) (:1271-1276), is_near_goal (:1226-1230) and the goal segment test (:1096); lanes split the circles.
struct ExtResult { double nx, ny; int cf; };
__device__ __noinline__ ExtResult extend_candidate_v(const double4 *s_obs, int n_obs, double fx, double fy, double tx, double ty,
                                                    double ed, double gx, double gy) {
    double nx, ny;
    int cf;
    const int lane = threadIdx.x & 31;
    double st, ct;
    (void)crm_atan2_sincos(ty - fy, tx - fx, &st, &ct);
    nx = fx + ed * ct; ny = fy + ed * st;
    const double d0 = crm_hypot(fx - nx, fy - ny);
    const double ex = fx + ct * d0, ey = fy + st * d0;
    const bool near_goal = crm_hypot(nx - gx, ny - gy) < ed;
    bool he = false, hg = false;
    for (int j = lane; j < n_obs; j += 32) {
        const double4 o = s_obs[j];
        he |= seg_dd(fx, fy, ex, ey, o.x, o.y) <= o.w;
        if (near_goal) hg |= seg_dd(nx, ny, gx, gy, o.x, o.y) <= o.w;
    }
    he = __any_sync(0xffffffffu, he);
    hg = __any_sync(0xffffffffu, hg);
    cf = (he ? CF_BLOCKED : 0) | (near_goal ? CF_NEAR_GOAL : 0) | (hg ? CF_GOAL_BLOCKED : 0);
    ExtResult r;
    r.nx = nx; r.ny = ny; r.cf = cf;
    return r;
}
// (result by value across the call: reference outputs would go through local memory)
#define extend_candidate(obs_, n_, fx_, fy_, tx_, ty_, ed_, gx_, gy_, nx_, ny_, cf_) do { \
        const ExtResult er_ = extend_candidate_v((obs_), (n_), (fx_), (fy_), (tx_), (ty_), (ed_), (gx_), (gy_)); \
        (nx_) = er_.nx; (ny_) = er_.ny; (cf_) = er_.cf; } while (0)


// Exact `.index()` resolution when two hits at different positions have a bit-equal d^2 (rare): every CTA publishes
// its hit list, each hit looks for a lower-index hit with the same d^2 (-> shadowed), the best-parent partials are
// recomputed and exchanged again.
__device__ __noinline__ void resolve_equal_d2(TreeSmem &S, const TreeArgs &A, const TreeWs &ws, unsigned long long &bphase,
                                             unsigned &seq, int G, int H, long long seg, double nx, double ny, Winner &W) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const double INF = CUDART_INF;
    {
            for (int e = tid; e < H; e += TREE_T) {
                const bool in_s = e < TREE_HCAP;
                const int i = (in_s ? S.hit_tag[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP))) & HIT_MASK;
                double ax, ay;
                if (in_s) { ax = S.hit_x[e]; ay = S.hit_y[e]; } else { const double2 a = ld_xy(A.xy + i); ax = a.x; ay = a.y; }
                const double qx = ax - nx, qy = ay - ny;
                __stcg(ws.g_idx + seg + e, i);
                __stcg(ws.g_d2 + seg + e, qx * qx + qy * qy);
            }
            if (tid == 0) __stcg(ws.g_cnt + cta, H);
            grid_barrier(ws, bphase, G);
            double cc = INF;
            int ci = NO_IDX;
            for (int e = tid; e < H; e += TREE_T) {
                const bool in_s = e < TREE_HCAP;
                int tag = in_s ? S.hit_tag[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
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
                if (in_s) S.hit_tag[e] = tag; else __stcg(ws.sp_idx + seg + (e - TREE_HCAP), tag);
                if (!(tag & HIT_SHADOW) && (tag & HIT_FREE)) {
                    const double d = in_s ? S.hit_d[e] : ld_f64(ws.sp_d + seg + (e - TREE_HCAP));
                    const double c_i = in_s ? S.hit_c[e] : ld_f64(ws.sp_c + seg + (e - TREE_HCAP));
                    lexmin(cc, ci, c_i + d, i);
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
                lexmin(cc, ci, __shfl_xor_sync(0xffffffffu, cc, off), __shfl_xor_sync(0xffffffffu, ci, off));
            if (lane == 0) { S.w_cc[warp] = cc; S.w_ci[warp] = ci; S.w_fl[warp] = 0; S.w_hs[warp] = 0; }
            if (tid == 0) { S.c_d2 = INF; S.c_idx = NO_IDX; S.c_nx = 0.0; S.c_ny = 0.0; S.c_cf = 0; }
            __syncthreads();
            const Winner W2 = exchange(S, ws.rec, ws.bar, seq, G);
        seq++;
            W.cp_cost = W2.cp_cost; W.cp_idx = W2.cp_idx;
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
    const int stride = G * TREE_T;   // nodes per ownership round
    const long long seg = (long long)cta * ws.seg_cap;

    for (int j = tid; j < n_obs; j += TREE_T) S.obs[j] = A.obstacles[j];
    for (int j = tid; j < (ws.seg_cap >> 5); j += TREE_T) S.dupbits[j] = 0u;
    if (cta == 0 && tid == 0) { A.xy[0] = make_double2(sx, sy); A.cost[0] = 0.0; A.parent[0] = -1; }
    __syncthreads();

    unsigned long long bphase = 0;
    unsigned seq = 0;
    grid_barrier(ws, bphase, G);  // node 0 visible everywhere

    int n = 1, status = RRTK_Q_OK, plen_best = 0, it = 0;
    double c_best = INF;
    const double c_min = crm_hypot(sx - gx, sy - gy);
    const double xc = (sx + gx) / 2.0, yc = (sy + gy) / 2.0;
    long long total_hits = 0;
    int n_slow = 0, n_goal = 0, n_redo = 0, n_reext = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};  // CTA 0's clock per phase: cull, scan, cand + hits, exchange, apply, goal / redo
    long long t0 = clock64();
#define TREE_TICK(k) do { long long t1 = clock64(); cyc[k] += t1 - t0; t0 = t1; } while (0)

    // current iteration: nearest index `ni`, its extended node (cur_x, cur_y) and flags; next sample (s1x, s1y)
    int ni = 0, cur_cf = 0;
    double cur_x = 0.0, cur_y = 0.0, s1x = 0.0, s1y = 0.0;
    if (A.p.max_iter > 0) {
        if (warp == 0) {
            double rx, ry, nx, ny;
            int cf;
            draw_sample(A, 0, c_best, c_min, xc, yc, rx, ry);
            extend_candidate(S.obs, n_obs, sx, sy, rx, ry, ed, gx, gy, nx, ny, cf);
            if (lane == 0) { S.c_nx = nx; S.c_ny = ny; S.c_cf = cf; }
        } else if (warp == 1 && A.p.max_iter > 1) {
            double a, b;
            draw_sample(A, 1, c_best, c_min, xc, yc, a, b);
            if (lane == 0) { S.s1x = a; S.s1y = b; }
        }
        __syncthreads();
        cur_x = S.c_nx; cur_y = S.c_ny; cur_cf = S.c_cf; s1x = S.s1x; s1y = S.s1y;
    }

    // One pass over the owned nodes + exchange.  r2 < 0: nearest only.  Leaves the hit list in shared memory.
    auto pass = [&](double nx, double ny, double r, double r2, bool want_nn, double tx, double ty, int draw_it, int &H,
                    int &ncull_out) -> Winner {
        if (tid == 0) { S.nhit = 0; S.ncull = 0; }
        __syncthreads();
        if (r2 >= 0.0) {
            for (int j = tid; j < n_obs; j += TREE_T) {
                const double4 o = S.obs[j];
                const double dx = o.x - nx, dy = o.y - ny;
                const double lim = (r + o.z) * (1.0 + 1e-9) + 1e-9;
                if (dx * dx + dy * dy <= lim * lim) S.cull[atomicAdd(&S.ncull, 1)] = j;
            }
        }
        if (warp == 1 && draw_it >= 0) {  // the sample after next, while the others scan
            double a, b;
            draw_sample(A, draw_it, c_best, c_min, xc, yc, a, b);
            if (lane == 0) { S.s2x = a; S.s2y = b; }
        }
        TREE_TICK(0);
        // ---- scan ----
        double bd = INF, bx = 0.0, by = 0.0;
        int bi = NO_IDX, myhits = 0;
        if (r2 >= 0.0 || want_nn) {
            int chunk = 0;
            for (long long base = (long long)cta * TREE_T + tid; base < n; base += (long long)TREE_UNROLL * stride, chunk += TREE_UNROLL) {
                double2 a[TREE_UNROLL];
                bool ok[TREE_UNROLL];
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    const long long i = base + (long long)u * stride;
                    ok[u] = i < n;
                    a[u] = ok[u] ? ld_xy(A.xy + i) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    if (!ok[u]) continue;
                    const int i = (int)(base + (long long)u * stride);
                    const double ax = a[u].x - nx, ay = a[u].y - ny;
                    if (ax * ax + ay * ay <= r2) {
                        myhits++;
                        const int li = (chunk + u) * TREE_T + tid;
                        if (!((S.dupbits[li >> 5] >> (li & 31)) & 1u)) {   // coincident twins are shadowed: skip
                            const int pos = atomicAdd(&S.nhit, 1);
                            if (pos < TREE_HCAP) { S.hit_tag[pos] = i; S.hit_x[pos] = a[u].x; S.hit_y[pos] = a[u].y; }
                            else __stcg(ws.sp_idx + seg + (pos - TREE_HCAP), i);
                        }
                    }
                    const double ex = a[u].x - tx, ey = a[u].y - ty;
                    const double e2 = ex * ex + ey * ey;
                    if (e2 < bd) { bd = e2; bi = i; bx = a[u].x; by = a[u].y; }
                }
            }
        }
        {
            const int src = warp_argmin_lane(bd, bi);
            bd = __shfl_sync(0xffffffffu, bd, src); bi = __shfl_sync(0xffffffffu, bi, src);
            bx = __shfl_sync(0xffffffffu, bx, src); by = __shfl_sync(0xffffffffu, by, src);
        }
        if (lane == 0) { S.w_nd[warp] = bd; S.w_ni[warp] = bi; S.w_nx[warp] = bx; S.w_ny[warp] = by; }
        __syncthreads();
        TREE_TICK(1);
        H = S.nhit;
        const int ncull = S.ncull;
        ncull_out = ncull;
        double cc = INF;
        int ci = NO_IDX, flags = 0;
        if (warp == TREE_CAND_WARP) {
            // ---- cand: extend the CTA's nearest candidate for the next sample ----
            bd = lane < TREE_NW ? S.w_nd[lane] : INF; bi = lane < TREE_NW ? S.w_ni[lane] : NO_IDX;
            bx = lane < TREE_NW ? S.w_nx[lane] : 0.0; by = lane < TREE_NW ? S.w_ny[lane] : 0.0;
            {
                const int src = warp_argmin_lane(bd, bi);
                bd = __shfl_sync(0xffffffffu, bd, src); bi = __shfl_sync(0xffffffffu, bi, src);
                bx = __shfl_sync(0xffffffffu, bx, src); by = __shfl_sync(0xffffffffu, by, src);
            }
            double cx = 0.0, cy = 0.0;
            int cf = 0;
            if (want_nn && bi != NO_IDX) extend_candidate(S.obs, n_obs, bx, by, tx, ty, ed, gx, gy, cx, cy, cf);
            if (lane == 0) { S.c_d2 = want_nn ? bd : INF; S.c_idx = want_nn ? bi : NO_IDX; S.c_nx = cx; S.c_ny = cy; S.c_cf = cf; }
        } else {
            // ---- hits: choose_parent candidates among the owned hits (rrt_07:1110-1135) ----
            unsigned long long *tab = ws.tab + ((unsigned long long)(it & 1) << TREE_TAB_BITS);
            for (int e = tid; e < H; e += TREE_HIT_THREADS) {
                const bool in_s = e < TREE_HCAP;
                int i;
                double ax, ay;
                if (in_s) { i = S.hit_tag[e]; ax = S.hit_x[e]; ay = S.hit_y[e]; }
                else { i = ld_i32(ws.sp_idx + seg + (e - TREE_HCAP)); const double2 a = ld_xy(A.xy + i); ax = a.x; ay = a.y; }
                const double c_i = ld_f64(A.cost + i);
                const double qx = ax - nx, qy = ay - ny;
                const double d2 = qx * qx + qy * qy;
                // hash set of d^2 bit patterns: a repeated key means two different positions at equal d^2
                const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
                unsigned h = (unsigned)(splitmix64(key) >> (64 - TREE_TAB_BITS));
                unsigned long long old = atomicCAS(tab + h, TREE_EMPTY, key);
                const double d = crm_hypot(nx - ax, ny - ay);
                const bool free_e = edge_free(ax, ay, nx, ny, d, band_k, S.obs, S.cull, ncull);
                int slot = -1, tag = i;
                if (d2 == 0.0) flags |= FLAG_DUP_NEW;
                for (int probe = 0;; probe++) {
                    if (old == TREE_EMPTY) { slot = (int)h; break; }
                    if (old == key || probe == TREE_PROBES) { flags |= FLAG_EQ_D2; break; }
                    h = (h + 1) & ((1u << TREE_TAB_BITS) - 1u);
                    old = atomicCAS(tab + h, TREE_EMPTY, key);
                }
                if (free_e) { tag |= HIT_FREE; lexmin(cc, ci, c_i + d, i); }
                if (in_s) { S.hit_tag[e] = tag; S.hit_slot[e] = slot; S.hit_d[e] = d; S.hit_c[e] = c_i; }
                else {
                    __stcg(ws.sp_idx + seg + (e - TREE_HCAP), tag);
                    __stcg(ws.sp_slot + seg + (e - TREE_HCAP), slot);
                    __stcg(ws.sp_d + seg + (e - TREE_HCAP), d);
                    __stcg(ws.sp_c + seg + (e - TREE_HCAP), c_i);
                }
            }
        }
        {
            const int src = warp_argmin_lane(cc, ci);
            cc = __shfl_sync(0xffffffffu, cc, src); ci = __shfl_sync(0xffffffffu, ci, src);
            flags = (int)__reduce_or_sync(0xffffffffu, (unsigned)flags);
            myhits = (int)__reduce_add_sync(0xffffffffu, (unsigned)myhits);
        }
        if (lane == 0) { S.w_cc[warp] = cc; S.w_ci[warp] = ci; S.w_fl[warp] = flags; S.w_hs[warp] = myhits; }
        __syncthreads();
        TREE_TICK(2);
        const Winner W = exchange(S, ws.rec, ws.bar, seq, G);
        seq++;
        TREE_TICK(3);
        return W;
    };

    // `redo`: the pass is a nearest-only repeat for iteration it+1 after c_best changed (one call site keeps the
    // hot code small: the kernel is instruction-fetch sensitive, L1.5 I-cache = 32 KB)
    bool redo = false;
    while (redo || it < A.p.max_iter) {
        const bool have1 = it + 1 < A.p.max_iter, have2 = it + 2 < A.p.max_iter;
        const bool accept = !redo && !(cur_cf & CF_BLOCKED);
        const bool goal_event = accept && (cur_cf & CF_NEAR_GOAL) && !(cur_cf & CF_GOAL_BLOCKED);
        if (accept && n >= A.p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; break; }
        const double nx = cur_x, ny = cur_y;
        const double2 rr2 = __ldg(A.near_rr2 + n);  // (r, r ** 2) for n_node = n (rrt_07:1138-1139)
        int H = 0, ncull = 0;
        Winner W = pass(nx, ny, rr2.x, accept ? rr2.y : -1.0, redo || have1, s1x, s1y, have2 ? it + 2 : -1, H, ncull);
        const double s2x = S.s2x, s2y = S.s2y;
        total_hits += W.hits;

        if (!accept) {  // collision: nothing is added (rrt_07:1080-1082); or the nearest-only repeat
            ni = W.nn_idx; cur_x = W.nx; cur_y = W.ny; cur_cf = W.cf;
            s1x = s2x; s1y = s2y;
            redo = false;
            it++;
            continue;
        }
        const int flags = W.flags;

        // ---- slow path: equal d^2 at different positions -> exact shadow flags from the global hit list ----
        if (flags & FLAG_EQ_D2) {
            n_slow++;
            resolve_equal_d2(S, A, ws, bphase, seq, G, H, seg, nx, ny, W);
        }

        // ---- apply: parent choice, append, rewire (rrt_07:1232-1246) ----
        double ncost;
        int npar;
        if (W.cp_idx != NO_IDX) { ncost = W.cp_cost; npar = W.cp_idx; }
        else { ncost = ld_f64(A.cost + ni) + ed; npar = ni; }
        const int newi = n;
        if (cta == (newi / TREE_T) % G && tid == 0) {
            __stcg(A.xy + newi, make_double2(nx, ny));
            __stcg(A.cost + newi, ncost);
            __stcg(A.parent + newi, npar);
            if (flags & FLAG_DUP_NEW) {
                const int li = (int)(newi / stride) * TREE_T + newi % TREE_T;
                S.dupbits[li >> 5] |= 1u << (li & 31);
            }
        }
        {
            unsigned long long *tab = ws.tab + ((unsigned long long)(it & 1) << TREE_TAB_BITS);
            for (int e = tid; e < H; e += TREE_T) {
                const bool in_s = e < TREE_HCAP;
                const int tag = in_s ? S.hit_tag[e] : ld_i32(ws.sp_idx + seg + (e - TREE_HCAP));
                const int slot = in_s ? S.hit_slot[e] : ld_i32(ws.sp_slot + seg + (e - TREE_HCAP));
                if (slot >= 0) tab[slot] = TREE_EMPTY;
                if ((tag & HIT_SHADOW) || !(tag & HIT_FREE)) continue;
                const int i = tag & HIT_MASK;
                const double sc = ncost + (in_s ? S.hit_d[e] : ld_f64(ws.sp_d + seg + (e - TREE_HCAP)));
                const double c_i = in_s ? S.hit_c[e] : ld_f64(ws.sp_c + seg + (e - TREE_HCAP));
                if (c_i > sc) { __stcg(A.parent + i, newi); __stcg(A.cost + i, sc); }
            }
        }
        n++;
        // the new node joins the nearest candidates of sample it+1 (highest index: it loses ties)
        bool reextend = false;
        if (have1) {
            const double bx = nx - s1x, by = ny - s1y;
            reextend = bx * bx + by * by < W.nn_d2;
        }
        TREE_TICK(4);

        // ---- goal bookkeeping (rrt_07:1094-1103) ----
        bool c_changed = false;
        if (goal_event) {
            n_goal++;
            grid_barrier(ws, bphase, G);  // this iteration's rewires are now visible
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

        if (have1 && c_changed) {  // samples it+1, it+2 depend on c_best: redraw, redo the nearest search (incl. new node)
            n_redo++;
            if (warp == 0) {
                double a, b;
                draw_sample(A, it + 1, c_best, c_min, xc, yc, a, b);
                if (lane == 0) { S.s1x = a; S.s1y = b; }
            }
            __syncthreads();
            s1x = S.s1x; s1y = S.s1y;
            redo = true;
            TREE_TICK(5);
            continue;
        }
        if (have1) {
            if (reextend) {  // the node just appended is the nearest of sample it+1: extend from it
                n_reext++;
                if (warp == 0) {
                    double cx, cy;
                    int cf;
                    extend_candidate(S.obs, n_obs, nx, ny, s1x, s1y, ed, gx, gy, cx, cy, cf);
                    if (lane == 0) { S.c_nx = cx; S.c_ny = cy; S.c_cf = cf; }
                }
                __syncthreads();
                ni = newi; cur_x = S.c_nx; cur_y = S.c_ny; cur_cf = S.c_cf;
            } else {
                ni = W.nn_idx; cur_x = W.nx; cur_y = W.ny; cur_cf = W.cf;
            }
            s1x = s2x; s1y = s2y;
        }
        it++;
        TREE_TICK(5);
    }

    if (tid == 0) {   // per-phase clocks: CTA 0's own, plus max / min over the CTAs (result struct is zeroed by the launcher)
        for (int k = 0; k < 6; k++) {
            atomicMax((long long *)&A.res->cycles_max[k], cyc[k]);
            atomicMax((long long *)&A.res->cycles_negmin[k], -cyc[k]);
        }
    }
    if (cta == 0 && tid == 0) {
        rrtk_informed_tree_result r;
        r.n_nodes = n; r.path_len = plen_best; r.status = status | (plen_best > A.p.path_cap ? RRTK_Q_PATH_OVERFLOW : 0);
        r.iters_done = it; r.c_best = c_best; r.total_hits = total_hits; r.slow_paths = n_slow;
        r.goal_events = n_goal; r.resamples = n_redo; r.grid = G; r.reextends = n_reext; r.pad_ = 0;
        A.res->n_nodes = r.n_nodes; A.res->path_len = r.path_len; A.res->status = r.status; A.res->iters_done = r.iters_done;
        A.res->c_best = r.c_best; A.res->total_hits = r.total_hits; A.res->slow_paths = r.slow_paths;
        A.res->goal_events = r.goal_events; A.res->resamples = r.resamples; A.res->grid = r.grid;
        A.res->reextends = r.reextends; A.res->pad_ = 0;
        for (int k = 0; k < 6; k++) A.res->cycles[k] = cyc[k];
    }
}


// ================================================================================================================
// Batched form: B samples per pass ("batch-parallel sampling", BASELINE config 3).  The sequential semantics are kept
// exactly by only batching samples that provably do not interact, and cutting the batch short where they might:
//   pass A   nearest of the B samples among the nodes that exist at the batch start (one scan, B argmins);
//   exchange winners; warp k extends winner k (exact leaf math, B warps in parallel on the same code);
//   cut #1   sample j leaves the batch (and everything after it) if an accepted new node i < j is nearer to s_j than
//            its winner (its nearest would be an in-batch node), or lies within the near radius of new node j (it would be
//            in j's near list); the batch also ends right after the first sample that connects to the goal (c_best, and
//            with it every later sample, may change);
//   pass B   near hits of all accepted new nodes in one scan; a node hit by two samples marks the later one;
//   exchange best parent per sample + the masks (overlap, equal-d^2, coincident);
//   cut #2   the batch ends before the first sample whose near list overlaps an earlier one (its costs could have been
//            rewired) or that needs the exact equal-d^2 resolution (then it runs alone, first in the next batch);
//   apply    appends and rewires of the surviving samples (disjoint near lists: any order), goal bookkeeping.
// Samples cut from a batch are simply redone by the next one, against the tree as it then is.
// ================================================================================================================
constexpr int TB_MAX = 8;
constexpr int TB_REC = 16;           // 16-byte pieces per CTA record
constexpr int TB_SUB_BITS = TREE_TAB_BITS - 3;   // one hash sub-table per sample
constexpr int TB_HCAP = 4096;        // near hits of one batch owned by one CTA (shared memory); a fuller batch is retried smaller

struct BatchSmem {
    double4 obs[TREE_OBS_CAP];
    short cull[TB_MAX][TREE_OBS_CAP];
    int hit_tag[TB_HCAP], hit_slot[TB_HCAP];
    unsigned char hit_k[TB_HCAP];
    double hit_x[TB_HCAP], hit_y[TB_HCAP], hit_d[TB_HCAP], hit_c[TB_HCAP];
    double w_d[TB_MAX][TREE_NW];      // per-warp partial (value, index) per sample
    int w_i[TB_MAX][TREE_NW];
    int w_m[TREE_NW], w_h[TB_MAX][TREE_NW];
    double q_d[TB_MAX][TREE_NW];      // per-warp partial of the cross-CTA reduce
    int q_i[TB_MAX][TREE_NW], q_h[TB_MAX][TREE_NW], q_m[TREE_NW];
    double c_r2[TB_MAX], c_r[TB_MAX];        // cut #1 results per sample: near radius^2, radius (idx in c_idx)
    int c_idx[TB_MAX];
    double sx[2 * TB_MAX], sy[2 * TB_MAX];   // sample window: [k] = sample of iteration it + k; [0, B) is the batch, the
                                             // rest is drawn ahead by the idle warps of the extend phase
    double fx[TB_MAX], fy[TB_MAX];    // nearest node position
    double nx[TB_MAX], ny[TB_MAX];    // extended node
    double nn_d2[TB_MAX], cp_cost[TB_MAX], plen;
    int nn_idx[TB_MAX], cf[TB_MAX], cp_idx[TB_MAX], hits[TB_MAX], ncull[TB_MAX];
    int nhit, g_mask, gcount;
    unsigned dupbits[1];
};

// counter barrier of the batch kernel (records were stored before by this CTA's warps)
__device__ __forceinline__ void batch_sync(unsigned long long *bar, const unsigned seq, int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        atomicAdd(bar + 8, 1ull);
        const unsigned long long target = (unsigned long long)(seq + 1u) * (unsigned long long)G;
        unsigned long long v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar + 8) : "memory"); } while (v < target);
            (void)ld_acquire(bar + 8);   // acquire: pairs with the releasing fence + atomicAdd of every CTA, so the record / tree reads below are ordered after them
    }
    __syncthreads();
}

// Cross-CTA argmin per sample.  In: S.w_d / w_i [k][warp] per-warp partials of this CTA (and S.w_h hit counts, S.w_m
// masks when `with_masks`).  Out: S.q_d[k][0], S.q_i[k][0] (+ S.hits[k], S.g_mask), valid for every thread after return.
__device__ __noinline__ void batch_exchange(BatchSmem &S, uint4 *rec, unsigned long long *bar, const unsigned seq, int G, int B,
                                            bool with_masks) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint4 *recs = rec + (size_t)(seq & 1u) * G * TB_REC;
    if (warp < B) {   // warp k reduces the CTA's partials of sample k and stores its piece
        const bool in = lane < TREE_NW;
        const double v = in ? S.w_d[warp][lane] : CUDART_INF;
        const int i = in ? S.w_i[warp][lane] : NO_IDX;
        const int src = warp_argmin_lane(v, i);
        const double bv = __shfl_sync(0xffffffffu, v, src);
        const int bi = __shfl_sync(0xffffffffu, i, src);
        const int hs = with_masks ? (int)__reduce_add_sync(0xffffffffu, in ? (unsigned)S.w_h[warp][lane] : 0u) : 0;
        if (lane == 0) __stcg(recs + (size_t)blockIdx.x * TB_REC + warp, piece(bv, bi, (unsigned)hs));
    } else if (warp == TB_MAX && with_masks) {
        const int m = (int)__reduce_or_sync(0xffffffffu, lane < TREE_NW ? (unsigned)S.w_m[lane] : 0u);
        if (lane == 0) __stcg(recs + (size_t)blockIdx.x * TB_REC + TB_MAX, make_uint4((unsigned)m, 0u, 0u, 0u));
    }
    batch_sync(bar, seq, G);
    double v[TB_MAX];
    int i[TB_MAX], h[TB_MAX], m = 0;
#pragma unroll
    for (int k = 0; k < TB_MAX; k++) { v[k] = CUDART_INF; i[k] = NO_IDX; h[k] = 0; }
    if (tid < G) {
        const uint4 *r = recs + (size_t)tid * TB_REC;
#pragma unroll
        for (int k = 0; k < TB_MAX; k++)
            if (k < B) { const uint4 pc = __ldcg(r + k); v[k] = piece_f64(pc); i[k] = (int)pc.z; h[k] = (int)pc.w; }
        if (with_masks) m = (int)__ldcg(r + TB_MAX).x;
    }
    const int nwq = (G + 31) >> 5;
    if (warp < nwq) {
#pragma unroll
        for (int k = 0; k < TB_MAX; k++) {
            if (k >= B) break;
            const int src = warp_argmin_lane(v[k], i[k]);
            const double bv = __shfl_sync(0xffffffffu, v[k], src);
            const int bi = __shfl_sync(0xffffffffu, i[k], src);
            const int hs = with_masks ? (int)__reduce_add_sync(0xffffffffu, (unsigned)h[k]) : 0;
            if (lane == 0) { S.q_d[k][warp] = bv; S.q_i[k][warp] = bi; S.q_h[k][warp] = hs; }
        }
        if (with_masks) { m = (int)__reduce_or_sync(0xffffffffu, (unsigned)m); if (lane == 0) S.q_m[warp] = m; }
    }
    __syncthreads();
    if (tid < B) {   // thread k combines sample k
        double bv = CUDART_INF;
        int bi = NO_IDX, hs = 0;
        for (int w = 0; w < nwq; w++) { lexmin(bv, bi, S.q_d[tid][w], S.q_i[tid][w]); hs += S.q_h[tid][w]; }
        S.q_d[tid][0] = bv; S.q_i[tid][0] = bi; S.hits[tid] = hs;
    } else if (tid == TB_MAX && with_masks) {
        int mm = 0;
        for (int w = 0; w < nwq; w++) mm |= S.q_m[w];
        S.g_mask = mm;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TREE_T, 1) informed_tree_batch_kernel(const TreeArgs A) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    BatchSmem &S = *reinterpret_cast<BatchSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const TreeWs &ws = A.ws;
    const double INF = CUDART_INF;
    const double ed = A.p.expand_dis;
    const double sx0 = A.p.start_goal[0], sy0 = A.p.start_goal[1], gx = A.p.start_goal[2], gy = A.p.start_goal[3];
    const int n_obs = A.p.n_obs;
    const int BMAX = A.p.batch < 1 ? 1 : (A.p.batch > TB_MAX ? TB_MAX : A.p.batch);
    const double band_k = 1e-12 * A.p.coord_bound * A.p.coord_bound;
    const int stride = G * TREE_T;
    const long long seg = (long long)cta * ws.seg_cap;

    for (int j = tid; j < n_obs; j += TREE_T) S.obs[j] = A.obstacles[j];
    for (int j = tid; j < (ws.seg_cap >> 5); j += TREE_T) S.dupbits[j] = 0u;
    if (cta == 0 && tid == 0) { A.xy[0] = make_double2(sx0, sy0); A.cost[0] = 0.0; A.parent[0] = -1; }
    __syncthreads();
    unsigned long long bphase = 0;
    unsigned seq = 0;
    grid_barrier(ws, bphase, G);

    int n = 1, status = RRTK_Q_OK, plen_best = 0, it = 0;
    double c_best = INF;
    const double c_min = crm_hypot(sx0 - gx, sy0 - gy);
    const double xc = (sx0 + gx) / 2.0, yc = (sy0 + gy) / 2.0;
    long long total_hits = 0, n_batches = 0;
    int n_slow = 0, n_goal = 0, n_redo = 0, n_cut = 0;
    bool stop = false;
    int blimit = BMAX;   // halved after a batch whose hit list did not fit, restored after a good one
    int win_n = 0;       // valid entries of the sample window S.sx / S.sy (iterations it .. it + win_n - 1) ...
    double win_cb = 0.0; // ... drawn with this c_best
    // per-phase clocks of CTA 0 (reported in res->cycles[1..5]): pass A | exchange A | extend + cut + pass B + candidates |
    // exchange B + cut #2 | apply + goal
    long long tc[6] = {0, 0, 0, 0, 0, 0}, t_prev = clock64();
#define TB_TICK(k) do { const long long t_ = clock64(); tc[k] += t_ - t_prev; t_prev = t_; } while (0)
    // finer split (res->cycles_max[0..5]): samples | scan A | (argmin A = cycles[1] - these two) | extend | cut #1 + cull +
    // scan B | (candidates = cycles[3] - these two); slot 2 and 5 are left 0
    long long ts[6] = {0, 0, 0, 0, 0, 0}, t_sub = t_prev;
#define TB_SUB(k) do { const long long t_ = clock64(); ts[k] += t_ - t_sub; t_sub = t_; } while (0)
#define TB_SUB_RESET() do { t_sub = clock64(); } while (0)

    while (it < A.p.max_iter && !stop) {
        const int B = min(blimit, A.p.max_iter - it);
        n_batches++;
        // ---- samples (informed_sample, rrt_07:1145-1159) ----
        if (!(win_n >= B && win_cb == c_best)) {   // window empty or drawn for another ellipse: draw the batch now
            if (warp < B) {
                double a, b;
                draw_sample(A, it + warp, c_best, c_min, xc, yc, a, b);
                if (lane == 0) { S.sx[warp] = a; S.sy[warp] = b; }
            }
            win_n = B; win_cb = c_best;
        }
        if (tid == 0) { S.nhit = 0; S.gcount = 0; }
        if (tid < TB_MAX) S.ncull[tid] = 0;
        __syncthreads();
        TB_SUB(0);
        // ---- pass A: nearest of every sample among the owned nodes ----
        {
            double bd[TB_MAX], qx[TB_MAX], qy[TB_MAX];
            int bi[TB_MAX];
#pragma unroll
            for (int k = 0; k < TB_MAX; k++) { bd[k] = INF; bi[k] = NO_IDX; qx[k] = k < B ? S.sx[k] : 0.0; qy[k] = k < B ? S.sy[k] : 0.0; }
            for (long long base = (long long)cta * TREE_T + tid; base < n; base += (long long)TREE_UNROLL * stride) {
                double2 a[TREE_UNROLL];
                bool ok[TREE_UNROLL];
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    const long long i = base + (long long)u * stride;
                    ok[u] = i < n;
                    a[u] = ok[u] ? ld_xy(A.xy + i) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    if (!ok[u]) continue;
                    const int i = (int)(base + (long long)u * stride);
#pragma unroll
                    for (int k = 0; k < TB_MAX; k++) {
                        if (k >= B) break;
                        const double ex = a[u].x - qx[k], ey = a[u].y - qy[k];
                        const double e2 = ex * ex + ey * ey;
                        if (e2 < bd[k]) { bd[k] = e2; bi[k] = i; }
                    }
                }
            }
            TB_SUB(1);
#pragma unroll
            for (int k = 0; k < TB_MAX; k++) {
                if (k >= B) break;
                const int src = warp_argmin_lane(bd[k], bi[k]);
                const double v = __shfl_sync(0xffffffffu, bd[k], src);
                const int i = __shfl_sync(0xffffffffu, bi[k], src);
                if (lane == 0) { S.w_d[k][warp] = v; S.w_i[k][warp] = i; }
            }
        }
        __syncthreads();
        TB_TICK(1);
        batch_exchange(S, ws.rec, ws.bar, seq, G, B, false);
        seq++;
        TB_TICK(2);
        TB_SUB_RESET();
        // ---- warp k extends winner k (get_new_node + check_collision + goal tests, exact leaf math) ----
        if (warp < B) {
            const int wi = S.q_i[warp][0];
            const double2 f = ld_xy(A.xy + wi);
            double cx, cy;
            int cf;
            extend_candidate(S.obs, n_obs, f.x, f.y, S.sx[warp], S.sy[warp], ed, gx, gy, cx, cy, cf);
            if (lane == 0) {
                S.nn_idx[warp] = wi; S.nn_d2[warp] = S.q_d[warp][0]; S.fx[warp] = f.x; S.fy[warp] = f.y;
                S.nx[warp] = cx; S.ny[warp] = cy; S.cf[warp] = cf;
            }
        } else if (warp >= TB_MAX) {
            // look-ahead sampling by the warps that have no winner to extend: warp j draws the sample of iteration it + j,
            // so informed_sample's load + sqrt / sin / cos chain is off the critical path of the next batch
            if (win_n >= TB_MAX && warp >= win_n && it + warp < A.p.max_iter) {
                double a, b;
                draw_sample(A, it + warp, c_best, c_min, xc, yc, a, b);
                if (lane == 0) { S.sx[warp] = a; S.sy[warp] = b; }
            }
        }
        if (win_n >= TB_MAX) win_n = max(win_n, min(2 * TB_MAX, A.p.max_iter - it));
        __syncthreads();
        TB_SUB(3);
        // ---- cut #1 (uniform): in-batch dependencies, goal connection, capacity ----
        // A rolled loop over shared memory that every thread runs identically (its 28 unrolled pair tests were 29 KB of
        // straight-line code per batch); the per-sample results go through shared memory into the register arrays.
        int B1 = B, acc = 0, idx_of[TB_MAX], acc_mask = 0;
        double r2_of[TB_MAX], r_of[TB_MAX];
        bool goal_last = false;
        if (tid < TB_MAX) { S.c_idx[tid] = -1; S.c_r2[tid] = -1.0; S.c_r[tid] = 0.0; }
        __syncthreads();
#pragma unroll 1
        for (int j = 0; j < B; j++) {
            const bool acc_j = !(S.cf[j] & CF_BLOCKED);
            const double2 rr2 = __ldg(A.near_rr2 + (n + acc));
            const double sxj = S.sx[j], syj = S.sy[j], nxj = S.nx[j], nyj = S.ny[j], nnd = S.nn_d2[j];
            bool dep = false;
#pragma unroll 1
            for (int i = 0; i < j; i++) {
                if (!((acc_mask >> i) & 1)) continue;
                const double nxi = S.nx[i], nyi = S.ny[i];
                const double ax = nxi - sxj, ay = nyi - syj;
                dep |= ax * ax + ay * ay < nnd;                              // node i would be the nearest of sample j
                if (acc_j) {
                    const double bx = nxi - nxj, by = nyi - nyj;
                    dep |= bx * bx + by * by <= rr2.y;                       // node i would be in the near list of node j
                }
            }
            if (dep) { B1 = j; n_cut++; break; }
            if (acc_j) {
                if (n + acc >= A.p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; stop = true; B1 = j; break; }
                if (tid == 0) { S.c_idx[j] = n + acc; S.c_r2[j] = rr2.y; S.c_r[j] = rr2.x; }
                acc_mask |= 1 << j; acc++;
                if ((S.cf[j] & CF_NEAR_GOAL) && !(S.cf[j] & CF_GOAL_BLOCKED)) { B1 = j + 1; goal_last = true; break; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TB_MAX; j++) { idx_of[j] = S.c_idx[j]; r2_of[j] = S.c_r2[j]; r_of[j] = S.c_r[j]; }
        if (B1 == 0) break;   // only when the tree is full
        // ---- obstacle cull per accepted sample ----
        for (int j = tid; j < n_obs; j += TREE_T) {
            const double4 o = S.obs[j];
#pragma unroll
            for (int k = 0; k < TB_MAX; k++) {
                if (k >= B1) break;
                if (!((acc_mask >> k) & 1)) continue;
                const double dx = o.x - S.nx[k], dy = o.y - S.ny[k];
                const double lim = (r_of[k] + o.z) * (1.0 + 1e-9) + 1e-9;
                if (dx * dx + dy * dy <= lim * lim) S.cull[k][atomicAdd(&S.ncull[k], 1)] = (short)j;
            }
        }
        // ---- pass B: near hits of every accepted new node; a node hit twice marks the later sample ----
        int ovl = 0, myhits[TB_MAX];
#pragma unroll
        for (int k = 0; k < TB_MAX; k++) myhits[k] = 0;
        if (acc_mask) {
            double px[TB_MAX], py[TB_MAX];
#pragma unroll
            for (int k = 0; k < TB_MAX; k++) { px[k] = k < B1 ? S.nx[k] : 0.0; py[k] = k < B1 ? S.ny[k] : 0.0; }
            int chunk = 0;
            for (long long base = (long long)cta * TREE_T + tid; base < n; base += (long long)TREE_UNROLL * stride, chunk += TREE_UNROLL) {
                double2 a[TREE_UNROLL];
                bool ok[TREE_UNROLL];
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    const long long i = base + (long long)u * stride;
                    ok[u] = i < n;
                    a[u] = ok[u] ? ld_xy(A.xy + i) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < TREE_UNROLL; u++) {
                    if (!ok[u]) continue;
                    const int i = (int)(base + (long long)u * stride);
                    int m = 0;
#pragma unroll
                    for (int k = 0; k < TB_MAX; k++) {
                        if (k >= B1) break;
                        const double ax = a[u].x - px[k], ay = a[u].y - py[k];
                        if (ax * ax + ay * ay <= r2_of[k]) { m |= 1 << k; myhits[k]++; }
                    }
                    if (m) {
                        ovl |= m & (m - 1);
                        const int li = (chunk + u) * TREE_T + tid;
                        if (!((S.dupbits[li >> 5] >> (li & 31)) & 1u)) {
                            for (int mm = m; mm; mm &= mm - 1) {
                                const int k = __ffs(mm) - 1;
                                const int pos = atomicAdd(&S.nhit, 1);
                                if (pos < TB_HCAP) { S.hit_tag[pos] = i; S.hit_k[pos] = (unsigned char)k; S.hit_x[pos] = a[u].x; S.hit_y[pos] = a[u].y; }
                                else ovl |= 1 << 24;   // hit list of this CTA is full: reported through the mask
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        TB_SUB(4);
        const int H = min(S.nhit, TB_HCAP);
        // ---- choose_parent candidates of the owned hits (rrt_07:1110-1135), per sample ----
        int eq = 0, dupn = 0;
        unsigned long long *tab = ws.tab + ((unsigned long long)(n_batches & 1) << TREE_TAB_BITS);
#pragma unroll
        for (int k = 0; k < TB_MAX; k++) if (lane == 0) { S.w_d[k][warp] = INF; S.w_i[k][warp] = NO_IDX; }
        __syncwarp();
        for (int e0 = 0; e0 < H; e0 += TREE_T) {
            const int e = e0 + tid;
            int k = -1, i = NO_IDX;
            double c = INF;
            if (e < H) {
                i = S.hit_tag[e]; k = S.hit_k[e];
                const double ax = S.hit_x[e], ay = S.hit_y[e], nx = S.nx[k], ny = S.ny[k];
                const double c_i = ld_f64(A.cost + i);
                const double qx = ax - nx, qy = ay - ny;
                const double d2 = qx * qx + qy * qy;
                const unsigned long long key = (unsigned long long)__double_as_longlong(d2);
                unsigned long long *sub = tab + ((unsigned long long)k << TB_SUB_BITS);
                unsigned h = (unsigned)(splitmix64(key) >> (64 - TB_SUB_BITS));
                unsigned long long old = atomicCAS(sub + h, TREE_EMPTY, key);
                const double d = crm_hypot(nx - ax, ny - ay);
                bool unsure = !(d > 1e-9), hitc = false;
                for (int j = 0; j < S.ncull[k]; j++) {
                    const double4 o = S.obs[S.cull[k][j]];
                    const double dd = seg_dd(ax, ay, nx, ny, o.x, o.y);
                    hitc |= dd <= o.w;
                    unsure |= fabs(dd - o.w) <= band_k + 1e-12 * (dd + o.w);
                }
                bool free_e = !hitc;
                if (unsure) {   // exact end point (rare)
                    double s, cth;
                    (void)crm_atan2_sincos(ny - ay, nx - ax, &s, &cth);
                    const double ex = ax + cth * d, ey = ay + s * d;
                    free_e = true;
                    for (int j = 0; j < S.ncull[k] && free_e; j++) {
                        const double4 o = S.obs[S.cull[k][j]];
                        free_e = !(seg_dd(ax, ay, ex, ey, o.x, o.y) <= o.w);
                    }
                }
                int slot = -1, tag = i;
                if (d2 == 0.0) dupn |= 1 << k;
                for (int probe = 0;; probe++) {
                    if (old == TREE_EMPTY) { slot = (int)(((unsigned)k << TB_SUB_BITS) + h); break; }
                    if (old == key || probe == TREE_PROBES) { eq |= 1 << k; break; }
                    h = (h + 1) & ((1u << TB_SUB_BITS) - 1u);
                    old = atomicCAS(sub + h, TREE_EMPTY, key);
                }
                if (free_e) { tag |= HIT_FREE; c = c_i + d; }
                S.hit_tag[e] = tag; S.hit_slot[e] = slot; S.hit_d[e] = d; S.hit_c[e] = c_i;
            }
            // per-sample partial minimum of this round, folded into the warp's running partial
#pragma unroll
            for (int kk = 0; kk < TB_MAX; kk++) {
                if (kk >= B1) break;
                const double v = k == kk ? c : INF;
                const int src = warp_argmin_lane(v, k == kk ? i : NO_IDX);
                const double bv = __shfl_sync(0xffffffffu, v, src);
                const int bi = __shfl_sync(0xffffffffu, k == kk ? i : NO_IDX, src);
                if (lane == 0) lexmin(S.w_d[kk][warp], S.w_i[kk][warp], bv, bv < INF ? bi : NO_IDX);
            }
        }
        {
            const int mk = (int)__reduce_or_sync(0xffffffffu, (unsigned)(ovl | (eq << 8) | (dupn << 16)));
#pragma unroll
            for (int k = 0; k < TB_MAX; k++) {
                const int hs = (int)__reduce_add_sync(0xffffffffu, (unsigned)myhits[k]);
                if (lane == 0) S.w_h[k][warp] = hs;
            }
            if (lane == 0) S.w_m[warp] = mk;
        }
        __syncthreads();
        TB_TICK(3);
        batch_exchange(S, ws.rec, ws.bar, seq, G, B1, true);
        seq++;
        const int gm = S.g_mask;
        if ((gm >> 24) & 1) {   // some CTA's hit list overflowed: nothing was applied yet -- undo the hash inserts and retry smaller
            for (int e = tid; e < H; e += TREE_T) { const int slot = S.hit_slot[e]; if (slot >= 0) tab[slot] = TREE_EMPTY; }
            if (B == 1) { status |= RRTK_Q_NEAR_OVERFLOW; break; }
            blimit = B / 2;
            __syncthreads();
            continue;
        }
        blimit = BMAX;
        // ---- cut #2: overlapping near lists / equal-d^2 resolution ----
        int Be = B1;
        {
            const int bad = ((gm & 0xff) | ((gm >> 8) & 0xff)) & ~1;   // sample 0 cannot overlap; its equal-d^2 case is below
            if (bad) { Be = __ffs(bad) - 1; if (Be < B1) n_cut++; }
        }
        if (tid < TB_MAX) { S.cp_cost[tid] = S.q_d[tid][0]; S.cp_idx[tid] = S.q_i[tid][0]; }
        __syncthreads();
        if ((gm >> 8) & 1) {   // sample 0 has two hits at different positions with equal d^2: exact `.index()` shadowing
            n_slow++;
            Be = 1;
            for (int e = tid; e < H; e += TREE_T)
                if (S.hit_k[e] == 0) {
                    const int i = S.hit_tag[e] & HIT_MASK;
                    const double qx = S.hit_x[e] - S.nx[0], qy = S.hit_y[e] - S.ny[0];
                    const int pos = atomicAdd(&S.gcount, 1);
                    __stcg(ws.g_idx + seg + pos, i);
                    __stcg(ws.g_d2 + seg + pos, qx * qx + qy * qy);
                }
            __syncthreads();
            if (tid == 0) __stcg(ws.g_cnt + cta, S.gcount);
            grid_barrier(ws, bphase, G);
            double cc = INF;
            int ci = NO_IDX;
            for (int e = tid; e < H; e += TREE_T) {
                if (S.hit_k[e] != 0) continue;
                int tag = S.hit_tag[e];
                const int i = tag & HIT_MASK;
                const double qx = S.hit_x[e] - S.nx[0], qy = S.hit_y[e] - S.ny[0];
                const double d2 = qx * qx + qy * qy;
                bool shadow = false;
                for (int c2 = 0; c2 < G && !shadow; c2++) {
                    const int cnt = ld_i32(ws.g_cnt + c2);
                    const long long s2 = (long long)c2 * ws.seg_cap;
                    for (int q = 0; q < cnt; q++)
                        if (ld_i32(ws.g_idx + s2 + q) < i && ld_f64(ws.g_d2 + s2 + q) == d2) { shadow = true; break; }
                }
                if (shadow) { tag |= HIT_SHADOW; S.hit_tag[e] = tag; }
                if (!(tag & HIT_SHADOW) && (tag & HIT_FREE)) lexmin(cc, ci, S.hit_c[e] + S.hit_d[e], i);
            }
            {
                const int src = warp_argmin_lane(cc, ci);
                cc = __shfl_sync(0xffffffffu, cc, src); ci = __shfl_sync(0xffffffffu, ci, src);
                if (lane == 0) { S.w_d[0][warp] = cc; S.w_i[0][warp] = cc < INF ? ci : NO_IDX; S.w_h[0][warp] = 0; S.w_m[warp] = 0; }
            }
            __syncthreads();
            batch_exchange(S, ws.rec, ws.bar, seq, G, 1, true);
        seq++;
            if (tid == 0) { S.cp_cost[0] = S.q_d[0][0]; S.cp_idx[0] = S.q_i[0][0]; }
            __syncthreads();
        }
        TB_TICK(4);
        // ---- apply: append + rewire (rrt_07:1232-1246) for the surviving samples ----
        double ncost_of[TB_MAX];
        int npar_of[TB_MAX], n_acc = 0;
#pragma unroll
        for (int k = 0; k < TB_MAX; k++) {
            ncost_of[k] = INF; npar_of[k] = -1;
            if (k >= Be) continue;
            total_hits += S.hits[k];
            if (!((acc_mask >> k) & 1)) continue;
            if (S.cp_idx[k] != NO_IDX && S.cp_cost[k] < INF) { ncost_of[k] = S.cp_cost[k]; npar_of[k] = S.cp_idx[k]; }
            else { ncost_of[k] = ld_f64(A.cost + S.nn_idx[k]) + ed; npar_of[k] = S.nn_idx[k]; }
            const int newi = idx_of[k];
            if (cta == (newi / TREE_T) % G && tid == 0) {
                __stcg(A.xy + newi, make_double2(S.nx[k], S.ny[k]));
                __stcg(A.cost + newi, ncost_of[k]);
                __stcg(A.parent + newi, npar_of[k]);
                if ((gm >> (16 + k)) & 1) {
                    const int li = (newi / stride) * TREE_T + newi % TREE_T;
                    S.dupbits[li >> 5] |= 1u << (li & 31);
                }
            }
            n_acc++;
        }
        for (int e = tid; e < H; e += TREE_T) {
            const int tag = S.hit_tag[e], slot = S.hit_slot[e], k = S.hit_k[e];
            if (slot >= 0) tab[slot] = TREE_EMPTY;
            if (k >= Be || (tag & HIT_SHADOW) || !(tag & HIT_FREE)) continue;
            const int i = tag & HIT_MASK;
            double nc = INF;
            int ni = -1;
#pragma unroll
            for (int kk = 0; kk < TB_MAX; kk++) if (kk == k) { nc = ncost_of[kk]; ni = idx_of[kk]; }
            const double sc = nc + S.hit_d[e];
            if (S.hit_c[e] > sc) { __stcg(A.parent + i, ni); __stcg(A.cost + i, sc); }
        }
        // ---- goal bookkeeping (rrt_07:1094-1103): only the last surviving sample can have connected ----
        if (goal_last && Be == B1) {
            n_goal++;
            const int k = B1 - 1;
            int newi = -1, npar = -1;
#pragma unroll
            for (int kk = 0; kk < TB_MAX; kk++) if (kk == k) { newi = idx_of[kk]; npar = npar_of[kk]; }
            const double nx = S.nx[k], ny = S.ny[k];
            grid_barrier(ws, bphase, G);
            if (tid == 0) {
                double plen = 0.0, qx = gx, qy = gy;
                int c = newi;
                for (int guard = 0; guard <= A.p.node_cap; guard++) {
                    const int pk = c == newi ? npar : ld_i32(A.parent + c);
                    if (pk < 0) break;
                    const double2 a = c == newi ? make_double2(nx, ny) : ld_xy(A.xy + c);
                    plen += crm_hypot(a.x - qx, a.y - qy);
                    qx = a.x; qy = a.y;
                    c = pk;
                }
                plen += crm_hypot(sx0 - qx, sy0 - qy);
                S.plen = plen;
            }
            __syncthreads();
            const double plen = S.plen;
            if (plen < c_best) {
                c_best = plen;
                n_redo++;
                if (cta == 0 && tid == 0) {
                    int w = 0;
                    if (w < A.p.path_cap) A.path[w] = make_double2(gx, gy);
                    w++;
                    int c = newi;
                    for (int guard = 0; guard <= A.p.node_cap; guard++, w++) {
                        const int pk = c == newi ? npar : ld_i32(A.parent + c);
                        if (pk < 0) break;
                        if (w < A.p.path_cap) A.path[w] = c == newi ? make_double2(nx, ny) : ld_xy(A.xy + c);
                        c = pk;
                    }
                    if (w < A.p.path_cap) A.path[w] = make_double2(sx0, sy0);
                    w++;
                    plen_best = w;
                }
            }
        }
        n += n_acc;
        it += Be;
        __syncthreads();
        {   // slide the sample window by the Be iterations just consumed
            double vx = 0.0, vy = 0.0;
            const bool mv = tid + Be < win_n;
            if (mv) { vx = S.sx[tid + Be]; vy = S.sy[tid + Be]; }
            __syncthreads();
            if (mv) { S.sx[tid] = vx; S.sy[tid] = vy; }
            win_n = win_n > Be ? win_n - Be : 0;
            __syncthreads();
        }
        TB_TICK(5);
        TB_SUB_RESET();
    }
#undef TB_TICK
#undef TB_SUB
#undef TB_SUB_RESET

    if (cta == 0 && tid == 0) {
        A.res->n_nodes = n; A.res->path_len = plen_best;
        A.res->status = status | (plen_best > A.p.path_cap ? RRTK_Q_PATH_OVERFLOW : 0);
        A.res->iters_done = it; A.res->c_best = c_best; A.res->total_hits = total_hits; A.res->slow_paths = n_slow;
        A.res->goal_events = n_goal; A.res->resamples = n_redo; A.res->grid = G; A.res->reextends = n_cut; A.res->pad_ = 0;
        A.res->cycles[0] = n_batches;
        for (int k = 1; k < 6; k++) A.res->cycles[k] = tc[k];
        for (int k = 0; k < 6; k++) A.res->cycles_max[k] = ts[k];
    }
}


// Diagnostic: the grid-wide exchange alone, `iters` times on a co-resident grid; out[cta] = cycles per exchange.
__global__ void __launch_bounds__(TREE_T, 1) tree_exchange_probe_kernel(TreeWs ws, int iters, long long *out) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    TreeSmem &S = *reinterpret_cast<TreeSmem *>(smem_raw);
    const int tid = threadIdx.x, G = gridDim.x;
    if (tid < TREE_NW) { S.w_cc[tid] = 1.0 + tid; S.w_ci[tid] = tid; S.w_fl[tid] = 0; S.w_hs[tid] = 1; }
    if (tid == 0) { S.c_d2 = 1.0 + blockIdx.x; S.c_idx = blockIdx.x; S.c_nx = 0.5; S.c_ny = 0.25; S.c_cf = 0; }
    __syncthreads();
    unsigned seq = 0;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int k = 0; k < iters; k++) {
        const Winner W = exchange(S, ws.rec, ws.bar, seq, G);
        seq++;
        acc += W.nn_d2 + W.hits;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = (t1 - t0) / (iters > 0 ? iters : 1) + (acc < 0.0 ? 1 : 0);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace layout for a grid of G CTAs
static size_t carve(TreeWs &ws, char *base, int node_cap, int G) {
    const long long chunks = ((long long)node_cap + TREE_T - 1) / TREE_T;
    const int seg_cap = (int)((chunks + G - 1) / G) * TREE_T;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return base ? base + o : nullptr; };
    ws.bar = (unsigned long long *)take(256);
    ws.rec = (uint4 *)take(sizeof(uint4) * 2 * (TB_REC > TREE_REC_PIECES ? TB_REC : TREE_REC_PIECES) * G);
    ws.tab = (unsigned long long *)take(sizeof(unsigned long long) * 2 * (1ull << TREE_TAB_BITS));
    ws.sp_idx = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.sp_slot = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.sp_d = (double *)take(sizeof(double) * (size_t)G * seg_cap);
    ws.sp_c = (double *)take(sizeof(double) * (size_t)G * seg_cap);
    ws.g_idx = (int *)take(sizeof(int) * (size_t)G * seg_cap);
    ws.g_d2 = (double *)take(sizeof(double) * (size_t)G * seg_cap);
    ws.g_cnt = (int *)take(sizeof(int) * G);
    ws.seg_cap = seg_cap;
    return off;
}

static size_t tree_smem_bytes(int seg_cap, bool batch = false) {
    return (batch ? sizeof(BatchSmem) : sizeof(TreeSmem)) + sizeof(unsigned) * (size_t)(seg_cap / 32 + 1);
}

static int tree_grid(int want, int *grid_out) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDevice");
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceGetAttribute");
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
    const bool batched = p.batch > 1;
    const void *kern = batched ? (const void *)informed_tree_batch_kernel : (const void *)informed_tree_kernel;
    const size_t smem = tree_smem_bytes(A.ws.seg_cap, batched);
    int per_sm = 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(smem): node_cap too large for the owner-local flags");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TREE_T, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) return set_error(RRTK_ERR_CUDA, "informed tree kernel does not fit on an SM");
    e = cudaMemsetAsync(res, 0, sizeof(*res), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(A.ws.bar, 0, 256, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(A.ws.rec, 0, sizeof(uint4) * 2 * TB_REC * G, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(A.ws.tab, 0xff, sizeof(unsigned long long) * 2 * (1ull << TREE_TAB_BITS), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(workspace)");
    void *args[] = {(void *)&A};
    e = cudaLaunchCooperativeKernel(kern, dim3(G), dim3(TREE_T), args, smem, s);
    if (e != cudaSuccess) return set_cuda_error(e, "informed_tree_kernel cooperative launch");
    return RRTK_OK;
}

int launch_tree_exchange_probe(int grid, int iters, long long *out_dev, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    int G = 0;
    int rc = tree_grid(grid, &G);
    if (rc) return rc;
    TreeWs ws;
    const size_t need = carve(ws, (char *)workspace, TREE_T, G);
    if (need > workspace_bytes) return set_error(RRTK_ERR_INVALID, "workspace too small");
    const size_t smem = tree_smem_bytes(ws.seg_cap);
    cudaError_t e = cudaFuncSetAttribute(tree_exchange_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.bar, 0, 256, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.rec, 0, sizeof(uint4) * 2 * TREE_REC_PIECES * G, s);
    if (e != cudaSuccess) return set_cuda_error(e, "tree_exchange_probe setup");
    void *args[] = {(void *)&ws, (void *)&iters, (void *)&out_dev};
    e = cudaLaunchCooperativeKernel((const void *)tree_exchange_probe_kernel, dim3(G), dim3(TREE_T), args, smem, s);
    if (e != cudaSuccess) return set_cuda_error(e, "tree_exchange_probe_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
