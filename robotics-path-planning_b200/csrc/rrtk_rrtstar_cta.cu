// rrtk_rrtstar_cta.cu -- batched RRT / RRT*, ONE CTA (4 warps) PER QUERY, the tree in shared memory.
//
// Same loop and the same results, bit for bit, as the warp-per-query kernel (rrtk_rrtstar.cu): `planning()` of
// rrt_04:1036-1084 (and rrt_01:71-101 with RRT_ONLY).  What changes is where the time goes.  One warp per query is a
// serial chain of L2 round trips (node scan, children lists, candidate loads): 2 400 resident queries hide each other's
// latency, 512 do not -- and every multi-GPU configuration of BASELINE.json is a small batch per GPU.  Here a query
// owns a CTA and its working set lives in that CTA's shared memory:
//   s_xy[node_cap]   double2   node positions (FP64, the scan reads them with one LDS.128 per node)
//   s_link[node_cap] ushort2   children lists: first child, next sibling (propagate_cost_to_leaves, rrt_04:1379-1384)
//   s_par[node_cap]  ushort    parent (read when a node is re-parented)
// 22 B / node = 44 KB at config 2's 2001 nodes, four queries per SM.  Global xy / parent are written through (they are
// the outputs); node costs stay in global memory only: every read of a cost is either issued long before its use
// (candidates) or replaced by a value carried in shared memory (the propagate frontier carries the parent's cost).
// An iteration is two stages.  Stage A (nothing in it reads a cost or a parent):
//   scan        nearest (rrt_04:1196-1202) + speculative near set around the SAMPLE, strided over the scanning threads;
//               hits are appended unordered (shared-memory atomic)                                                B1
//   one warp    first edge: steer + play area + collision (edge_verdict_fast / exact steer)           } concurrently
//   one warp    the NEXT sample and the obstacle cull around it (its L2 latency never shows)          }
//   the rest    near list: rank by index (ascending order) + the `.index()` first-equal-d2 mapping    }          B2
// Stage B:
//   all warps   choose_parent (rrt_04:1242-1282): 4 lanes per candidate split the culled circles                B3
//   all warps   rewire edges (rrt_04:1340-1373), same split, only entries that pass node.cost > new.cost + d;
//               the few entries that can apply go to a compact list                                               B4
//   warp 0      ordered apply + propagate over the shared-memory children lists, append
// SOFTWARE PIPELINE: the apply phase of iteration i (warp 0 alone, ~a quarter of the iteration) and stage A of iteration
// i + 1 (warps 1-3: the new node's position is known since B3, and stage A reads positions only) run side by side whenever
// no node can MOVE in that apply phase and the goal is not searched every iteration; one barrier (B2) joins them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "crmath.h"
#include "rrtk_device.cuh"
#include "rrtk_planner.cuh"
// (draw_sample stays inline: out of line it takes the parameter block and the Sobol state by reference, i.e. through local memory)
#include "rrtk_rrtstar_common.cuh"

namespace rrtk {

#define CTA_RARE(c) __builtin_expect(!!(c), 0)   // rare branches leave the hot straight-line path (the kernel is fetch-bound)

constexpr int CTA_W = 4;              // warps per query
constexpr int CTA_T = CTA_W * 32;
constexpr int CTA_NC_SMALL = 256;     // layout capacities of the near list (p.near_cap <= NC): batches / single queries
constexpr int CTA_NC_LARGE = 1024;
constexpr int CTA_FQ = 64;            // shared-memory slots of propagate's frontier (more spill to the workspace)
constexpr int CTA_AL = 32;            // compact apply list (more: the apply phase reads the near-list arrays)
constexpr unsigned short NONE16 = 0xffffu;
constexpr int BIG = 0x7fffffff;

// -DRRTK_CTA_PROFILE (tools/probe_cta_phases.py, `make profile`): the TRACE instantiation records thread 0's SM clock per
// phase of every iteration instead of the decision trace
#ifdef RRTK_CTA_PROFILE
#define CTA_CLK(k) do { if (TRACE && tid == 0) clk[k] = clock64(); } while (0)
#else
#define CTA_CLK(k) do { } while (0)
#endif

// fixed part of the CTA's shared memory (every address an immediate); the tree arrays follow it
template <int NC>
struct CtaSmemT {
    double nd[NC];               // d2 to the new node, then hypot(new - node)
    double s_nc[NC];             // unordered d2 while the list is built, then the candidates' node costs
    double cull[2][3][CULL_CAP];     // culled circles (x, y, R2) of this iteration [it & 1] and the next
    double fq_cost[CTA_FQ];
    double red_d[CTA_W], red_ex[CTA_W], red_ey[CTA_W];
    double nx, ny;                   // the new node (warp 0 -> all)
    double smp[2][2];                // the sample of this iteration [it & 1] and the next (drawn one iteration ahead)
    double al_ec[CTA_AL], al_c0[CTA_AL], al_dk[CTA_AL];   // compact apply list: new cost, node cost seen, edge length
    double p_cx, p_cy, p_ccost, p_elen;   // the outstanding apply phase (CtaPending), kept here rather than in every thread's registers
    int near_idx[NC];
    int near_ok[NC];             // unordered indices while the list is built, then the rewire flags
    int red_i[CTA_W];
    int cull_m[2], cull_glob[2];
    int al_w[CTA_AL];                // compact apply list: (flag << 30) | (list position << 16) | node
    int count2[2];                   // near hits of this iteration's scan [it & 1] (the other one is being reset)
    int al_n2[2], any_moves2[2];     // compact-list length / "a node would move" of this iteration's stage B [it & 1]
    int cpok, tail;
    int p_n, p_best, p_count, p_cb, p_flags;
    long long sob_n;                 // Sobol state of the query (only the sampling warp touches it)
    // per-query pointers / constants used by one role only: read here at the point of use instead of being carried in
    // every thread's registers for the whole loop (the kernel sits at the 128-register cap; what it spills goes to L2)
    double2 *q_xy; int32_t *q_parent; double *q_elen; int *q_gidx; double *q_gcost; const double2 *q_stream;
    int32_t *q_gcnt; uint16_t *q_glists; double q_ginv;
    unsigned int sob_q0, sob_q1;
    int accept, near_valid, t_status;
    int done, gi, status_or, ni, leader;
    unsigned int q;
    unsigned short fq_idx[CTA_FQ];
};

template <int NC>
__host__ __device__ inline size_t cta_smem_bytes(int node_cap) {
    size_t b = (sizeof(CtaSmemT<NC>) + 15) & ~(size_t)15;
    b += (size_t)node_cap * 16;                    // s_xy
    b += (size_t)node_cap * 4;                     // s_link
    b += ((size_t)node_cap * 2 + 15) & ~(size_t)15;  // s_par
    return b;
}

static __device__ __forceinline__ void cta_link(ushort2 *s_link, int p, int c) {
    s_link[c].y = s_link[p].x;
    s_link[p].x = (unsigned short)c;
}
static __device__ __forceinline__ void cta_unlink(ushort2 *s_link, int p, int c) {
    const unsigned short nxt = s_link[c].y;
    unsigned short x = s_link[p].x;
    if (x == (unsigned short)c) { s_link[p].x = nxt; return; }
    for (;;) {
        const unsigned short y = s_link[x].y;
        if (y == (unsigned short)c || y == NONE16) break;
        x = y;
    }
    s_link[x].y = nxt;
}

// propagate_cost_to_leaves (rrt_04:1379-1384) by one warp over the shared-memory children lists; the frontier carries each
// node's new cost, so no cost is read back from global memory.  Slots >= CTA_FQ spill to the query's workspace.
// elen[c] = hypot(c - parent(c)) is cached from the moment the edge was made (see propagate_lists_elen, rrtk_planner.cuh):
// a level of the walk is a child-list hop, one load and one add.
template <class SM>
static __device__ __noinline__ void cta_propagate(int root, double root_cost, bool root_moved, const double2 *s_xy,
                                                  const ushort2 *s_link, double *cost, double *elen, SM &S, int *g_idx,
                                                  double *g_cost, int lane) {
    if (s_link[root].x == NONE16) return;
    volatile int *tail = &S.tail;
    if (lane == 0) { S.fq_idx[0] = (unsigned short)root; S.fq_cost[0] = root_cost; *tail = 1; }
    __syncwarp();
    for (int head = 0;;) {
        const int end = *tail;
        if (head >= end) break;
        const int k = head + lane;
        __syncwarp();
        if (k < end) {
            const int pn = k < CTA_FQ ? (int)S.fq_idx[k] : g_idx[k];
            const double cp = k < CTA_FQ ? S.fq_cost[k] : g_cost[k];
            const bool stale = root_moved && pn == root;
            for (unsigned short c = s_link[pn].x; c != NONE16;) {
                const ushort2 lc = s_link[c];
                double e = elen[c];
                if (stale || e != e) {
                    const double2 a = s_xy[pn], b = s_xy[c];
                    e = crm_hypot(b.x - a.x, b.y - a.y);
                    elen[c] = e;
                }
                const double cc = cp + e;
                cost[c] = cc;
                if (lc.x != NONE16) {
                    const int slot = atomicAdd(&S.tail, 1);
                    if (slot < CTA_FQ) { S.fq_idx[slot] = c; S.fq_cost[slot] = cc; }
                    else { g_idx[slot] = (int)c; g_cost[slot] = cc; }
                }
                c = lc.y;
            }
        }
        __syncwarp();
        head = end < head + 32 ? end : head + 32;
    }
}

// view of every G-th circle of L starting at `sub` (the share of one lane of a candidate's group of G lanes)
static __device__ __forceinline__ ObsList sub_list(const ObsList &L, int sub, int G) {
    ObsList s;
    s.ox = L.ox + (size_t)sub * L.stride; s.oy = L.oy + (size_t)sub * L.stride; s.r2 = L.r2 + (size_t)sub * L.stride;
    s.stride = L.stride * G;
    s.m = L.m > sub ? (L.m - sub + G - 1) / G : 0;
    return s;
}

// rewire entries [from, count) one at a time against the current tree, by warp 0 (a re-parented node MOVED,
// rrt_04:1365-1371: positions and costs seen by the parallel pass are stale).  Rare.
template <class SM>
static __device__ __noinline__ void cta_rewire_serial(const PlanConsts p, int from, int count, SM &S, double2 *s_xy,
                                                      ushort2 *s_link, unsigned short *s_par, double2 *xy, double *cost,
                                                      int32_t *parent, double *elen, int n, double cx, double cy, double ccost,
                                                      const ObsList &L, int *g_idx, double *g_cost, int lane, int &t_rwok,
                                                      int &t_rwap) {
    const double res = p.res;
    for (int k = from; k < count; k++) {
        const int i = S.near_idx[k];
        const double2 a = s_xy[i];
        Steer st = steer(cx, cy, a.x, a.y, CUDART_INF, res);
        const bool ok = edge_free_warp(cx, cy, st, a.x, a.y, L, lane) && inside_play(p, st.ex, st.ey);
        const double ec = ccost + st.d;
        t_rwok += ok ? 1 : 0;
        if (ok && cost[i] > ec) {
            __syncwarp();
            if (lane == 0) {
                cta_unlink(s_link, (int)s_par[i], i);
                cta_link(s_link, n, i);
                s_xy[i] = make_double2(st.ex, st.ey);
                xy[i] = make_double2(st.ex, st.ey);
                cost[i] = ec;
                parent[i] = n;
                s_par[i] = (unsigned short)n;
                elen[i] = __longlong_as_double(0x7ff8000000000000ll);   // (the node may have moved: recomputed on demand)
            }
            __syncwarp();
            t_rwap++;
            cta_propagate(i, ec, true, s_xy, s_link, cost, elen, S, g_idx, g_cost, lane);
            __syncwarp();
        }
    }
}

// near list of the CTA: the unordered hits (S.near_ok = index, S.s_nc = d2) -> S.near_idx / S.nd in ascending index
// order, every hit replaced by the FIRST hit (lowest index) with the same d2 -- `dist_list.index(i)` of rrt_04:1336-1337.
// Threads t0, t0 + nt, ... of the caller take the entries.
// (out of line and not unrolled: the kernel is bound by instruction fetch -- every byte of the per-iteration path counts)
template <class SM>
static __device__ __noinline__ void cta_rank_near(SM &S, int count, int t0, int nt) {
#pragma unroll 1
    for (int k = t0; k < count; k += nt) {
        const int ik = S.near_ok[k];
        const double dk = S.s_nc[k];
        int rank = 0, first = ik;
#pragma unroll 1
        for (int j = 0; j < count; j++) {
            const int ij = S.near_ok[j];
            rank += ij < ik ? 1 : 0;
            if (S.s_nc[j] == dk && ij < first) first = ij;
        }
        S.near_idx[rank] = first;
        S.nd[rank] = dk;
    }
}

// State of an iteration whose edges are done (after B4) and whose apply phase is outstanding.
struct CtaPending {
    double cx, cy, ccost, new_elen;   // the new node (after choose_parent), its cost, hypot(new - parent) or NaN
    int n, best, count, al_n, cb;     // its index, its parent, the near-list length, compact-list length, cull buffer of L
    bool compact, c_is_new;
};

// rewire, ordered apply (rrt_04:1361-1371) + propagate_cost_to_leaves (:1379-1384) + the append (:1064), by ONE warp.
// The entries that can apply (flag != 0 and node.cost > new.cost + d when the edges were made) come from the compact list
// S.al_* (P.compact: <= 32 entries in any order) or from the near-list arrays 32 at a time; they are taken in list order.
template <bool TRACE, class SM>
static __device__ __forceinline__ void cta_apply(const rrtk_rrtstar_params &p, const CtaPending &P, SM &S, double2 *s_xy,
                                              ushort2 *s_link, unsigned short *s_par, double2 *xy, double *cost, int32_t *parent,
                                              double *elen, const double4 *obs, int n_obs, int *g_idx, double *g_cost, int lane,
                                              int &t_rwok, int &t_rwap) {
    const double INF = CUDART_INF, NaN = __longlong_as_double(0x7ff8000000000000ll);
    const double res = p.path_resolution, cx = P.cx, cy = P.cy, ccost = P.ccost;
    const int n = P.n;
    bool dirty = false;      // a propagate ran: node costs must be re-read
    int fallback_from = -1;  // >= 0: a node moved; redo entries from here serially
    const int nb = P.compact ? (P.al_n > 0 ? 1 : 0) : P.count;
#pragma unroll 1
    for (int b0 = 0; b0 < nb && fallback_from < 0; b0 += 32) {
        int k = BIG, i = -1, fl = 0;
        double ecost = 0.0, snc = 0.0, dk = 0.0;
        if (P.compact) {
            if (lane < P.al_n) {
                const int w = S.al_w[lane];
                i = w & 0xffff; k = (w >> 16) & 0x3fff; fl = (w >> 30) & 3;
                ecost = S.al_ec[lane]; snc = S.al_c0[lane]; dk = S.al_dk[lane];
            }
        } else if (b0 + lane < P.count) {
            k = b0 + lane;
            i = S.near_idx[k];
            fl = S.near_ok[k];
            snc = S.s_nc[k];
            dk = S.nd[k];
            ecost = ccost + dk;
        }
        const unsigned okmask = __ballot_sync(FULL, fl != 0);
        if (TRACE) t_rwok += __popc(okmask);
        unsigned m = __ballot_sync(FULL, fl != 0 && (snc > ecost));
        while (m) {  // apply in list order
            const int kk = (int)__reduce_min_sync(FULL, ((m >> lane) & 1u) ? (unsigned)k : 0xffffffffu);
            const int b = __ffs(__ballot_sync(FULL, k == kk && ((m >> lane) & 1u))) - 1;
            m &= ~(1u << b);
            const int ii = __shfl_sync(FULL, i, b);
            const double ec = __shfl_sync(FULL, ecost, b);
            const double c0 = __shfl_sync(FULL, snc, b);
            const double dkb = __shfl_sync(FULL, dk, b);
            const int flb = __shfl_sync(FULL, fl, b);
            const double ci = dirty ? cost[ii] : c0;
            if (ci > ec) {
                const double2 a = s_xy[ii];
                double ex = a.x, ey = a.y;
                if (CTA_RARE(flb == 2)) {   // the exact edge stops short of the node: it moves there
                    const Steer st = steer(cx, cy, a.x, a.y, INF, res);
                    ex = st.ex; ey = st.ey;
                }
                const bool moved = (a.x != ex) || (a.y != ey);
                __syncwarp();
                if (lane == 0) {
                    cta_unlink(s_link, (int)s_par[ii], ii);
                    cta_link(s_link, n, ii);
                    if (moved) { s_xy[ii] = make_double2(ex, ey); xy[ii] = make_double2(ex, ey); }
                    cost[ii] = ec;
                    parent[ii] = n;
                    s_par[ii] = (unsigned short)n;
                    elen[ii] = moved ? NaN : dkb;   // hypot(node - new node), what propagate would compute
                }
                __syncwarp();
                if (TRACE) t_rwap++;
                cta_propagate(ii, ec, moved, s_xy, s_link, cost, elen, S, g_idx, g_cost, lane);
                __syncwarp();
                dirty = true;
                if (CTA_RARE(moved)) {
                    // the node no longer sits where the parallel pass saw it, and the costs of its descendants may have
                    // gone UP: every later entry is re-evaluated from the current tree, one at a time
                    fallback_from = kk + 1;
                    if (TRACE) t_rwok -= __popc(okmask >> b >> 1);  // recounted below (TRACE never runs compact)
                    break;
                }
            }
        }
    }
    if (CTA_RARE(fallback_from >= 0)) {
        ObsList L;
        if (S.cull_glob[P.cb]) {
            const double *g = reinterpret_cast<const double *>(obs);
            L.ox = g; L.oy = g + 1; L.r2 = g + 3; L.stride = 4; L.m = n_obs;
        } else {
            L.ox = S.cull[P.cb][0]; L.oy = S.cull[P.cb][1]; L.r2 = S.cull[P.cb][2]; L.stride = 1; L.m = S.cull_m[P.cb];
        }
        cta_rewire_serial(plan_consts(p), fallback_from, P.count, S, s_xy, s_link, s_par, xy, cost, parent, elen, n, cx, cy, ccost, L,
                          g_idx, g_cost, lane, t_rwok, t_rwap);
    }
    if (lane == 0) {
        s_xy[n] = make_double2(cx, cy); s_par[n] = (unsigned short)P.best;
        xy[n] = make_double2(cx, cy); cost[n] = ccost; parent[n] = P.best;
        elen[n] = P.new_elen;
        cta_link(s_link, P.best, n);
    }
    __syncwarp();
}

template <bool RRT_ONLY, bool TRACE, bool RESUME, int NC>
__global__ void __launch_bounds__(CTA_T, 4)
rrtstar_cta_kernel(rrtk_rrtstar_params p, const double4 *__restrict__ start_goal,
                   const double4 *__restrict__ obstacles, const int32_t *__restrict__ n_obs_arr,
                   const double *__restrict__ near_r2, const double2 *__restrict__ sample_stream,
                   const int64_t *__restrict__ sobol_offset, double2 *xy_all, double *cost_all,
                   int32_t *parent_all, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
                   int32_t *status_out, int32_t *trace_all, int32_t *workspace, unsigned int *counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef CtaSmemT<NC> CtaSmem;
    CtaSmem &S = *reinterpret_cast<CtaSmem *>(smem_raw);
    double2 *s_xy = reinterpret_cast<double2 *>(smem_raw + ((sizeof(CtaSmem) + 15) & ~(size_t)15));
    ushort2 *s_link = reinterpret_cast<ushort2 *>(s_xy + p.node_cap);
    unsigned short *s_par = reinterpret_cast<unsigned short *>(s_link + p.node_cap);
    // Roles go by VIRTUAL warp number: warp 0 of every CTA lands on the same scheduler (SM sub-partition) of its SM, and
    // virtual warp 0 carries most of the serial work, so the (up to four) resident CTAs rotate it: CTA slot c of the SM
    // makes its hardware warp c the leader.  %warpid = the warp's slot in the SM, four consecutive ones per CTA.
    if (threadIdx.x == 0) {
        unsigned hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        S.leader = (int)((hw / CTA_W) % CTA_W);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = ((int)(threadIdx.x >> 5) + CTA_W - S.leader) % CTA_W;
    const int tid = warp * 32 + lane;
    const int near_cap = p.near_cap;
    const double res = p.path_resolution;
    const double INF = CUDART_INF;
    const double reach = (p.near_r_max > p.expand_dis ? p.near_r_max : p.expand_dis) + res;
    const double goal_reach = p.expand_dis > res ? p.expand_dis : res;

    for (;;) {
        __syncthreads();   // the previous query's shared memory is no longer read
        if (tid == 0) S.q = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int q = S.q;
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

        // workspace of the query: [4 * node_cap ints: spill of the propagate frontier (idx, then cost)] [cached edge lengths
        // elen[i] = hypot(node i - its parent), NaN = not known] [obstacle cells]
        const int grid_cells = p.grid_nx * p.grid_ny;
        int32_t *wsq = workspace + (size_t)q * RRTK_RRTSTAR_WS_INTS(p.node_cap, p.grid_nx, p.grid_ny);
        int *g_idx = wsq;
        double *g_cost = reinterpret_cast<double *>(wsq + 2 * (size_t)((p.node_cap + 1) / 2));
        ObsGrid grid;
        grid.nx = p.grid_nx; grid.ny = p.grid_ny; grid.x0 = p.grid_x0; grid.y0 = p.grid_y0;
        grid.cell = p.grid_cell; grid.inv_cell = grid_cells > 0 ? 1.0 / p.grid_cell : 0.0;
        double *elen = reinterpret_cast<double *>(wsq + 4 * (size_t)p.node_cap);
        const double NaN = __longlong_as_double(0x7ff8000000000000ll);
        grid.cnt = wsq + 4 * (size_t)p.node_cap + 4 * (size_t)((p.node_cap + 1) / 2);
        grid.lists = reinterpret_cast<uint16_t *>(grid.cnt + grid_cells);
        if (tid == 0) {
            S.q_xy = xy; S.q_parent = parent; S.q_elen = elen; S.q_gidx = g_idx; S.q_gcost = g_cost; S.q_stream = stream;
            S.q_gcnt = grid.cnt; S.q_glists = grid.lists; S.q_ginv = grid.inv_cell;
        }

        int n = 1, status = RRTK_Q_OK, gi = -1, it = 0, it_prev = 0;
        SobolState sob;
        sob.n = sobol_base < 0 ? 0 : sobol_base;
        bool skip = false;
        if (!RESUME) {
            if (tid == 0) {
                s_xy[0] = make_double2(sg.x, sg.y);
                s_link[0] = make_ushort2(NONE16, NONE16);
                s_par[0] = NONE16;
                xy[0] = make_double2(sg.x, sg.y);
                cost[0] = 0.0;
                parent[0] = -1;
                elen[0] = 0.0;
            }
        } else {
            // continue the tree a previous call left in xy / cost / parent: reload it, rebuild the children lists (their
            // order only fixes the traversal order of propagate, not its values), skip the Sobol points already consumed.
            // A query that had finished (goal found in early-exit mode, or an overflow) stays as it is.
            n = n_nodes[q];
            const int st_prev = status_out[q];
            it_prev = iters_done[q];
            if ((st_prev & (RRTK_Q_NEAR_OVERFLOW | RRTK_Q_NODE_OVERFLOW)) ||
                ((RRT_ONLY || !p.search_until_max_iter) && goal_index[q] >= 0))
                skip = true;
            if (n < 1 || n > p.node_cap) { n = 1; status |= RRTK_Q_NODE_OVERFLOW; }
            for (int i = tid; i < n; i += CTA_T) {
                s_xy[i] = xy[i];
                s_link[i] = make_ushort2(NONE16, NONE16);
                const int pp = parent[i];
                s_par[i] = pp < 0 ? NONE16 : (unsigned short)pp;
                elen[i] = NaN;
            }
            __syncthreads();
            if (tid == 0)
                for (int i = 1; i < n; i++) { const unsigned short pp = s_par[i]; if (pp != NONE16 && (int)pp < n) cta_link(s_link, (int)pp, i); }
            if (p.sampler == RRTK_SAMPLER_SOBOL) {
                int used = 0;
                for (int k = lane; k < p.iter_offset; k += 32)
                    used += (int)(splitmix64(rng_key(p.seed, (uint64_t)(q + p.query_base), (uint64_t)k)) % 101ull) > p.goal_sample_rate;
                sob.n += (int64_t)__reduce_add_sync(FULL, (unsigned)used);
            }
        }
        if (skip) continue;   // uniform: every thread read the same outputs
        if (warp == 0 && grid_cells > 0) build_obstacle_grid(grid, obs, n_obs, reach, lane);
        if (tid == 0) {
            S.count2[0] = 0; S.count2[1] = 0; S.cpok = 0; S.done = 0; S.gi = -1; S.status_or = 0;
            S.al_n2[0] = S.al_n2[1] = 0; S.any_moves2[0] = S.any_moves2[1] = 0;
        }
        sobol2(sob.n, sob.q0, sob.q1);
        if (tid == 0) { S.sob_n = sob.n; S.sob_q0 = sob.q0; S.sob_q1 = sob.q1; }
        bool done = false;
        // the apply phase of an iteration may run beside stage A of the next one only when nothing it produces is read before
        // the next stage B: not with a goal search per iteration, not with a trace (its counters close an iteration)
        const bool can_overlap = !RRT_ONLY && !TRACE && p.search_until_max_iter != 0;
        const bool goal_mode = RRT_ONLY || p.search_until_max_iter == 0;   // a goal test closes every iteration
        const double inv_res = 1.0 / res, q_expand = floor(p.expand_dis / res);   // steer's n_expand at full extension
        __syncthreads();   // the obstacle cells are built
        // the sampling warp draws the samples (one iteration ahead) and gathers the circles around them
        if (warp == 1 && p.max_iter > 0) {
            SobolState sb;
            sb.n = S.sob_n; sb.q0 = S.sob_q0; sb.q1 = S.sob_q1;
            const Sample s0 = draw_sample(p, (int)q, 0, RESUME ? p.iter_offset : 0, gx, gy, stream, sb);
            if (lane == 0) { S.smp[0][0] = s0.x; S.smp[0][1] = s0.y; S.sob_n = sb.n; S.sob_q0 = sb.q0; S.sob_q1 = sb.q1; }
            if (inside_play(p, s0.x, s0.y)) {
                const ObsList L0 = cull_obstacles_grid(grid, obs, n_obs, s0.x, s0.y, reach, S.cull[0][0], S.cull[0][1],
                                                       S.cull[0][2], lane);
                if (lane == 0) { S.cull_m[0] = L0.m; S.cull_glob[0] = L0.stride == 4 ? 1 : 0; }
            }
        }
        __syncthreads();

        // the obstacle cell grid, rebuilt from the parameter block + shared memory where it is used
        auto mk_grid = [&]() {
            ObsGrid g;
            g.nx = p.grid_nx; g.ny = p.grid_ny; g.x0 = p.grid_x0; g.y0 = p.grid_y0; g.cell = p.grid_cell;
            g.inv_cell = S.q_ginv; g.cnt = S.q_gcnt; g.lists = S.q_glists;
            return g;
        };
        bool pending = false, overlap = false, appended = false;
#ifdef RRTK_CTA_STATS
        int st_ov = 0, st_pend = 0, st_big = 0, st_moves = 0;
        long long st_c[6] = {0, 0, 0, 0, 0, 0};   // [0] role work of this warp in overlapped rounds, [1] its B2 wait, [2] scan, [3] stage B choose, [4] rewire, [5] rounds
        long long st_t0 = 0, st_t1 = 0, st_r[4] = {0, 0, 0, 0}; int st_rn[4] = {0, 0, 0, 0}, st_cat = -1;   // round time by category
#endif
        // trace row of the iteration being closed
        int t_ni = 0, t_status = 0, t_near = 0, t_par = -1, t_rwok = 0, t_rwap = 0;

        for (it = 0;; it++) {
            const int cb = it & 1;
#ifdef RRTK_CTA_PROFILE
            long long clk[10];
#endif
#ifdef RRTK_CTA_STATS
            { const long long now = clock64(); if (st_cat >= 0) { st_r[st_cat] += now - st_t0; st_rn[st_cat]++; } st_t0 = now; }
#endif
            CTA_CLK(0);
            // ======== close iteration it - 1: its apply phase (warp 0), its trace row, the goal tests ========
            const bool ov = pending && overlap;           // ... while the other warps already run stage A of iteration it
            if (it > 0) {
                if (pending && warp == 0) {
                    CtaPending P;
                    P.cx = S.p_cx; P.cy = S.p_cy; P.ccost = S.p_ccost; P.new_elen = S.p_elen;
                    P.n = S.p_n; P.best = S.p_best; P.count = S.p_count; P.cb = S.p_cb;
                    P.al_n = TRACE ? 0 : S.al_n2[P.cb];
                    P.compact = !TRACE && P.al_n <= CTA_AL; P.c_is_new = (S.p_flags & 2) != 0;
                    cta_apply<TRACE>(p, P, S, s_xy, s_link, s_par, S.q_xy, cost, S.q_parent, S.q_elen, obs, n_obs, S.q_gidx, S.q_gcost, lane, t_rwok, t_rwap);
                }
                if (TRACE && tid == 0) {
#ifndef RRTK_CTA_PROFILE
                    int32_t *tr = trace + (size_t)(it - 1) * 8;
                    tr[0] = t_ni; tr[1] = t_status; tr[2] = t_near; tr[3] = t_par; tr[4] = S.cpok;
                    tr[5] = t_rwok; tr[6] = t_rwap; tr[7] = n;
#endif
                    S.cpok = 0;
                }
                // ---- goal tests, by warp 0 (it wrote the last node itself); scratch = the closed iteration's circle list ----
                if (goal_mode && warp == 0 && !done) {
                    __syncwarp();
                    if (RRT_ONLY) {
                        // goal test on the last node (rrt_01:90-96)
                        const double2 last = s_xy[n - 1];
                        if (crm_hypot(last.x - gx, last.y - gy) <= p.expand_dis) {
                            ObsList G = cull_obstacles(obs, n_obs, gx, gy, goal_reach, S.cull[cb ^ 1][0], S.cull[cb ^ 1][1], S.cull[cb ^ 1][2], lane);
                            Steer st = steer(last.x, last.y, gx, gy, p.expand_dis, res);
                            if (edge_free_warp(last.x, last.y, st, gx, gy, G, lane) && lane == 0) { S.gi = n - 1; S.done = 1; }
                        }
                    } else {
                        bool ovf = false;
                        // obstacles that can touch an edge into the goal (search_best_goal_node steers end there)
                        ObsList G = cull_obstacles(obs, n_obs, gx, gy, goal_reach, S.cull[cb ^ 1][0], S.cull[cb ^ 1][1], S.cull[cb ^ 1][2], lane);
                        const int g = best_goal(plan_consts(p), n, s_xy, cost, gx, gy, G, S.near_idx, S.nd, near_cap, lane, ovf);
                        if (lane == 0) {
                            if (ovf) S.status_or |= RRTK_Q_NEAR_OVERFLOW;
                            if (g >= 0) { S.gi = g; S.done = 1; }
                        }
                    }
                }
                if (!ov && (pending || appended || goal_mode)) {
                    __syncthreads();   // ---- BA: stage A below reads what the apply phase / the append wrote
                    pending = false;
                }
                if (goal_mode && S.done) { gi = S.gi; done = true; }
            }
            CTA_CLK(9);
            if (it >= p.max_iter || done) break;
            appended = false;
            t_rwok = 0; t_rwap = 0; t_near = 0; t_par = -1;

            // ======== stage A of iteration it ========
            // roles.  Alone: 128 threads scan, warp 0 takes the first edge, warp 1 the next sample, warps 2-3 rank.  Beside an
            // apply phase (ov): warp 1 the next sample at once, warps 2-3 scan, then warp 2 the first edge and warp 3 ranks.
            const bool scanning = !ov || warp >= 2;
            const int scan_t0 = ov ? tid - 64 : tid, scan_nt = ov ? 64 : CTA_T;
            const int nslots = ov ? 2 : CTA_W;
            const bool r_first = ov ? warp == 2 : warp == 0, r_samp = warp == 1;
            const double rx = S.smp[cb][0], ry = S.smp[cb][1];
            const double r2 = RRT_ONLY ? -1.0 : near_r2[n + 1];   // near radius^2 at the current tree size (rrt_04:1329-1335)
            const int n_scan = ov ? n - 1 : n;   // (the node of the outstanding apply phase is not in shared memory yet)
            // ---- get_nearest_node_index (rrt_04:1196-1202), merged with a SPECULATIVE find_near_nodes around the sample
            // (when the steered node snaps onto the sample -- the common case once the tree is dense -- the near scan
            // would compute exactly these d2 again) ----
            double bd = INF;
            int bi = BIG;
            if (scanning) {
#pragma unroll 2
                for (int i = scan_t0; i < n_scan; i += scan_nt) {
                    const double2 a = s_xy[i];
                    const double ddx = a.x - rx, ddy = a.y - ry;
                    const double d = ddx * ddx + ddy * ddy;
                    if (d < bd) { bd = d; bi = i; }
                    if (d <= r2) {
                        const int slot = atomicAdd(&S.count2[cb], 1);
                        if (slot < near_cap) { S.near_ok[slot] = i; S.s_nc[slot] = d; }
                    }
                }
                if (ov && scan_t0 == 0) {   // the node the outstanding apply phase appends: index n - 1 at (S.p_cx, S.p_cy)
                    const double ddx = S.p_cx - rx, ddy = S.p_cy - ry;
                    const double d = ddx * ddx + ddy * ddy;
                    if (d < bd) { bd = d; bi = n - 1; }
                    if (d <= r2) {
                        const int slot = atomicAdd(&S.count2[cb], 1);
                        if (slot < near_cap) { S.near_ok[slot] = n - 1; S.s_nc[slot] = d; }
                    }
                }
#ifdef RRTK_CTA_STATS
                if (ov) st_c[2] += clock64() - st_t0;
#endif
                warp_argmin(bd, bi);
                if (lane == 0) { S.red_d[ov ? warp - 2 : warp] = bd; S.red_i[ov ? warp - 2 : warp] = bi; }
                CTA_CLK(1);
                // ---- B1: the scanning warps
                if (ov) asm volatile("bar.sync 1, %0;" ::"n"(64) : "memory");
                else __syncthreads();
                bd = S.red_d[0]; bi = S.red_i[0];
#pragma unroll 1
                for (int w = 1; w < nslots; w++) {
                    const double dw = S.red_d[w];
                    const int iw = S.red_i[w];
                    if (dw < bd || (dw == bd && iw < bi)) { bd = dw; bi = iw; }
                }
            }
            int ni = bi;
            int count = 0;
            CTA_CLK(2);
            if (r_first) {
                // ---- steer towards the sample (rrt_04:1051-1052).  Fast form: if the edge certainly snaps onto the sample the
                // new node IS the sample and only the collision verdict is needed; otherwise the exact steer runs ----
                const double2 from = (ov && ni == n - 1) ? make_double2(S.p_cx, S.p_cy) : s_xy[ni];
                int f_status = 0;
                bool accept = false, near_valid = false;
                double nx = rx, ny = ry;
                // (the verdict keeps a 1e-9 margin on every use of the edge length: a plain sqrt is as good as the correctly
                // rounded hypot here; the exact steer computes its own)
                const double ddx0 = rx - from.x, ddy0 = ry - from.y;
                const double d0 = sqrt(ddx0 * ddx0 + ddy0 * ddy0);
                int v = -1;
                ObsList L;
                L.ox = S.cull[cb][0]; L.oy = S.cull[cb][1]; L.r2 = S.cull[cb][2]; L.stride = 1; L.m = 0;
                {
                    if (snap_certain(d0, false, p.expand_dis, q_expand, res, inv_res)) {
                        if (inside_play(p, nx, ny)) {
                            if (S.cull_glob[cb]) {
                                const double *g = reinterpret_cast<const double *>(obs);
                                L.ox = g; L.oy = g + 1; L.r2 = g + 3; L.stride = 4; L.m = n_obs;
                            } else {
                                L.m = S.cull_m[cb];
                            }
                            // (the <true> instantiation: the body the candidate passes run, already in the instruction cache)
                            const int vl = edge_verdict_fast<true>(from.x, from.y, rx, ry, d0, false, p.expand_dis, q_expand, res, inv_res, L,
                                                                   lane, 32, ~0ull).v;
                            const unsigned blocked = __ballot_sync(FULL, vl == 0), unsure = __ballot_sync(FULL, vl < 0);
                            v = blocked ? 0 : (unsure ? -1 : 1);
                            if (v >= 0) { f_status = 1; accept = v == 1; near_valid = true; }
                        } else {
                            v = 0;  // outside the play area: rejected before the collision check (rrt_04:1054)
                        }
                    }
                }
                if (v < 0) {
                    Steer e0 = steer(from.x, from.y, rx, ry, p.expand_dis, res);
                    nx = e0.ex; ny = e0.ey;
                    if (inside_play(p, nx, ny)) {
                        f_status = 1;
                        // (the list prefetched for the sample only holds if the node landed on it)
                        L = cull_obstacles_grid(mk_grid(), obs, n_obs, nx, ny, reach, S.cull[cb][0], S.cull[cb][1], S.cull[cb][2], lane);
                        if (lane == 0) { S.cull_m[cb] = L.m; S.cull_glob[cb] = L.stride == 4 ? 1 : 0; }
                        accept = edge_free_warp(from.x, from.y, e0, rx, ry, L, lane);
                    }
                }
                if (lane == 0) {
                    S.nx = nx; S.ny = ny; S.accept = accept ? 1 : 0; S.near_valid = near_valid ? 1 : 0; S.t_status = f_status;
                    S.ni = ni;
                    if (accept && n < p.node_cap) s_link[n] = make_ushort2(NONE16, NONE16);  // children arrive through rewire
                    // counters of the NEXT scan (last read right after B2 of iteration it - 1, which every thread left before
                    // this iteration's B1 / the previous B3) and of THIS iteration's stage B (last read after B4 of iteration
                    // it - 2; B2 of iteration it - 1 lies in between)
                    S.count2[cb ^ 1] = 0; S.al_n2[cb] = 0; S.any_moves2[cb] = 0;
                }
            } else if (r_samp) {
                // ---- the NEXT sample and the circles around it (consumed by iteration it + 1 when its node lands on the sample) ----
                if (it + 1 < p.max_iter) {
                    SobolState sb;
                    sb.n = S.sob_n; sb.q0 = S.sob_q0; sb.q1 = S.sob_q1;
                    const Sample sn = draw_sample(p, (int)q, it + 1, RESUME ? it + 1 + p.iter_offset : it + 1, gx, gy, S.q_stream, sb);
                    if (lane == 0) { S.smp[cb ^ 1][0] = sn.x; S.smp[cb ^ 1][1] = sn.y; S.sob_n = sb.n; S.sob_q0 = sb.q0; S.sob_q1 = sb.q1; }
                    if (inside_play(p, sn.x, sn.y)) {
                        const ObsList Ln = cull_obstacles_grid(mk_grid(), obs, n_obs, sn.x, sn.y, reach, S.cull[cb ^ 1][0],
                                                               S.cull[cb ^ 1][1], S.cull[cb ^ 1][2], lane);
                        if (lane == 0) { S.cull_m[cb ^ 1] = Ln.m; S.cull_glob[cb ^ 1] = Ln.stride == 4 ? 1 : 0; }
                    }
                }
            } else if (!RRT_ONLY && scanning) {
                const int cnt = S.count2[cb];
                if (cnt <= near_cap) {
                    if (ov) cta_rank_near(S, cnt, lane, 32);
                    else cta_rank_near(S, cnt, tid - 64, 64);
                }
            }
#ifdef RRTK_CTA_STATS
            st_t1 = clock64();
            if (ov) { st_c[0] += st_t1 - st_t0; st_c[5]++; }
#endif
            CTA_CLK(3);
            __syncthreads();   // ---- B2 (also the end of an apply phase that ran beside this stage A)
#ifdef RRTK_CTA_STATS
            if (ov) st_c[1] += clock64() - st_t1;
            st_t1 = clock64();
#endif
            pending = false;
            CTA_CLK(4); CTA_CLK(5); CTA_CLK(6); CTA_CLK(7); CTA_CLK(8);

            // ======== stage B of iteration it ========
            ni = S.ni; count = S.count2[cb];
            bool accept = S.accept != 0;
            const double nx = S.nx, ny = S.ny;
            t_status = S.t_status; t_ni = ni;
            if (accept && n >= p.node_cap) { status |= RRTK_Q_NODE_OVERFLOW; accept = false; done = true; }
            if (accept && RRT_ONLY) {
                if (tid == 0) {
                    s_xy[n] = make_double2(nx, ny); s_par[n] = (unsigned short)ni;
                    S.q_xy[n] = make_double2(nx, ny); cost[n] = 0.0; S.q_parent[n] = ni;
                }
                t_status = 2; t_par = ni;
                n++;
                appended = true;
            } else if (accept) {
                if (CTA_RARE(!S.near_valid)) {
                    // ---- find_near_nodes (rrt_04:1314-1338) around the new node (it is not the sample) ----
                    __syncthreads();   // (every thread has read the speculative count)
                    if (tid == 0) S.count2[cb] = 0;
                    __syncthreads();
                    for (int i = tid; i < n; i += CTA_T) {
                        const double2 a = s_xy[i];
                        const double ddx = a.x - nx, ddy = a.y - ny;
                        const double d = ddx * ddx + ddy * ddy;
                        if (d <= r2) {
                            const int slot = atomicAdd(&S.count2[cb], 1);
                            if (slot < near_cap) { S.near_ok[slot] = i; S.s_nc[slot] = d; }
                        }
                    }
                    __syncthreads();
                    count = S.count2[cb];
                    if (count <= near_cap) cta_rank_near(S, count, tid, CTA_T);
                    __syncthreads();
                }
                if (CTA_RARE(count > near_cap)) {
                    status |= RRTK_Q_NEAR_OVERFLOW;
                    done = true;
                } else {
                    t_near = count;
                    ObsList L;
                    if (S.cull_glob[cb]) {
                        const double *g = reinterpret_cast<const double *>(obs);
                        L.ox = g; L.oy = g + 1; L.r2 = g + 3; L.stride = 4; L.m = n_obs;
                    } else {
                        L.ox = S.cull[cb][0]; L.oy = S.cull[cb][1]; L.r2 = S.cull[cb][2]; L.stride = 1; L.m = S.cull_m[cb];
                    }
                    // lanes per candidate: as many as keep the whole near list in ONE round of the 128 threads (a second round
                    // costs a warp the fixed part of an edge again -- hypot, the quotient, the error band -- which is more than
                    // the longer share of the circles; and with 4 lanes each, warp 0 alone took the second round of a list of
                    // 33 .. 40 entries while the others waited at the barrier)
                    const int G = count <= CTA_T / 4 ? 4 : (count <= CTA_T / 2 ? 2 : 1), CPR = CTA_T / G;
                    const int sub = tid & (G - 1), grp = lane & ~(G - 1);
                    // ---- choose_parent (rrt_04:1242-1282): G lanes per candidate, each tests its share of the circles ----
                    double bc = INF, bex = 0.0, bey = 0.0;
                    int bk = BIG;
                    // this lane's circles near the segment (its first-round candidate) - (new node), for the reverse edge
                    unsigned long long seg_near = ~0ull;
#pragma unroll 1
                    for (int k0 = 0; k0 < count; k0 += CPR) {
                        const int k = k0 + tid / G;
                        const bool valid = k < count;
                        double dk = 0.0, ci = 0.0, ex = nx, ey = ny;
                        bool blocked = false, exact = false;
                        if (valid) {
                            const int i = S.near_idx[k];
                            const double2 a = s_xy[i];
                            ci = cost[i];
                            dk = crm_hypot(nx - a.x, ny - a.y);   // what steer's calc_distance_and_angle returns
                            const EdgeVerdict ev = edge_verdict_fast<true>(a.x, a.y, nx, ny, dk, true, INF, INF, res, inv_res, L, sub, G, ~0ull);
                            const int vv = ev.v;
                            if (k0 == 0) seg_near = vv == 1 ? ev.near : ~0ull;
                            blocked = vv == 0;                    // (the new node is inside the play area)
                            if (CTA_RARE(vv < 0)) {
                                Steer st = steer(a.x, a.y, nx, ny, INF, res);
                                blocked = !(edge_free_lane(a.x, a.y, st, nx, ny, sub_list(L, sub, G)) && inside_play(p, st.ex, st.ey));
                                ex = st.ex; ey = st.ey;
                                exact = true;
                            }
                        }
                        const unsigned bm = (__ballot_sync(FULL, blocked) >> grp) & ((1u << G) - 1u);
                        const unsigned em = (__ballot_sync(FULL, exact) >> grp) & ((1u << G) - 1u);
                        const int src = grp + (em ? __ffs(em) - 1 : 0);
                        ex = __shfl_sync(FULL, ex, src);          // an exactly steered edge ends where ITS steer ends
                        ey = __shfl_sync(FULL, ey, src);
                        if (valid && sub == 0) {
                            S.nd[k] = dk;      // = hypot(new - node), calc_new_cost's distance (rrt_04:1375-1377)
                            S.s_nc[k] = ci;
                            if (!bm) {
                                if (TRACE) atomicAdd(&S.cpok, 1);
                                const double c = ci + dk;
                                if (c < bc) { bc = c; bk = k; bex = ex; bey = ey; }
                            }
                        }
                    }
                    {
                        const double mine = bc;
                        const int mk = bk;
                        warp_argmin(bc, bk);   // first minimum of the cost list (this warp's candidates)
                        const unsigned wm = __ballot_sync(FULL, mk == bk && mine == bc && bk != BIG);
                        if (wm) {
                            const int src = __ffs(wm) - 1;
                            bex = __shfl_sync(FULL, bex, src);
                            bey = __shfl_sync(FULL, bey, src);
                        }
                        if (lane == 0) { S.red_d[warp] = bc; S.red_i[warp] = bk; S.red_ex[warp] = bex; S.red_ey[warp] = bey; }
                    }
#ifdef RRTK_CTA_STATS
                    st_c[3] += clock64() - st_t1; st_t1 = clock64();
#endif
                    CTA_CLK(5);
                    __syncthreads();   // ---- B3
                    CTA_CLK(6); CTA_CLK(7); CTA_CLK(8);
                    bc = S.red_d[0]; bk = S.red_i[0]; bex = S.red_ex[0]; bey = S.red_ey[0];
#pragma unroll 1
                    for (int w = 1; w < CTA_W; w++) {
                        const double dw = S.red_d[w];
                        const int kw = S.red_i[w];
                        if (dw < bc || (dw == bc && kw < bk)) { bc = dw; bk = kw; bex = S.red_ex[w]; bey = S.red_ey[w]; }
                    }
                    if (bk != BIG) {
                        const int best = S.near_idx[bk];
                        // the node is re-steered from the winner (rrt_04:1279): same edge as above
                        const double cx = bex, cy = bey, ccost = bc;
                        const bool c_is_new = (cx == nx) && (cy == ny);  // the winner's edge snapped
                        const double new_elen = c_is_new ? S.nd[bk] : __longlong_as_double(0x7ff8000000000000ll);
                        if (tid == 0) {   // the outstanding apply phase (read after B4: by warp 0, and by the next stage A for the node)
                            S.p_cx = cx; S.p_cy = cy; S.p_ccost = ccost; S.p_elen = new_elen;
                            S.p_n = n; S.p_best = best; S.p_count = count; S.p_cb = cb; S.p_flags = c_is_new ? 2 : 0;
                        }
                        // ---- rewire (rrt_04:1340-1373), edges.  An entry can only be re-parented if node.cost > new.cost + d
                        // (:1362); costs never increase while the apply loop runs (unless a node MOVES, handled there), so the
                        // steer + collision of an entry that fails the test now is dead work.  With a trace every edge is
                        // evaluated (the trace counts collision-free rewire edges).  Flags: 0 = not applicable, 1 = free and
                        // ends on the node, 2 = free but the exact steer stops short of it (the node would MOVE).  The
                        // entries with a flag also go to the compact list the apply phase reads. ----
#pragma unroll 1
                        for (int k0 = 0; k0 < count; k0 += CPR) {
                            const int k = k0 + tid / G;
                            const bool valid = k < count;
                            double dk = 0.0;
                            bool want = false, blocked = false, moves = false;
                            int i = 0;
                            if (valid) {
                                i = S.near_idx[k];
                                const double2 a = s_xy[i];
                                // hypot(node - c) == the forward edge's d when c is the sample point itself
                                dk = c_is_new ? S.nd[k] : crm_hypot(a.x - cx, a.y - cy);
                                want = TRACE || (S.s_nc[k] > ccost + dk);
                                // the reverse edge runs along the segment choose_parent tested (when the new node is the sample):
                                // only the circles found near it then can matter, usually none
                                const unsigned long long only = (c_is_new && k0 == 0) ? seg_near : ~0ull;
                                if (want && only == 0ull) {
                                    blocked = !inside_play(p, a.x, a.y);
                                } else if (want) {
                                    const int vv = edge_verdict_fast<true>(cx, cy, a.x, a.y, dk, true, INF, INF, res, inv_res, L, sub, G, only).v;
                                    if (CTA_RARE(vv < 0)) {
                                        Steer st = steer(cx, cy, a.x, a.y, INF, res);
                                        blocked = !(edge_free_lane(cx, cy, st, a.x, a.y, sub_list(L, sub, G)) && inside_play(p, st.ex, st.ey));
                                        moves = (st.ex != a.x) || (st.ey != a.y);
                                    } else {   // snapped: the edge ends on the node itself (it does not move)
                                        blocked = !(vv == 1 && inside_play(p, a.x, a.y));
                                    }
                                }
                            }
                            const unsigned bm = (__ballot_sync(FULL, blocked) >> grp) & ((1u << G) - 1u);
                            const unsigned mm = (__ballot_sync(FULL, moves) >> grp) & ((1u << G) - 1u);
                            if (valid && sub == 0) {
                                if (!c_is_new) S.nd[k] = dk;
                                const int fl = (want && !bm) ? (mm ? 2 : 1) : 0;
                                S.near_ok[k] = fl;
                                if (!TRACE && fl) {
                                    const int slot = atomicAdd(&S.al_n2[cb], 1);
                                    if (slot < CTA_AL) {
                                        S.al_w[slot] = (fl << 30) | (k << 16) | i;
                                        S.al_ec[slot] = ccost + dk; S.al_c0[slot] = S.s_nc[k]; S.al_dk[slot] = dk;
                                    }
                                    if (fl == 2) S.any_moves2[cb] = 1;
                                }
                            }
                        }
#ifdef RRTK_CTA_STATS
                        st_c[4] += clock64() - st_t1;
#endif
                        CTA_CLK(7);
                        __syncthreads();   // ---- B4
                        CTA_CLK(8);
                        // the apply phase is now outstanding: warp 0 runs it at the top of the next round -- beside that
                        // iteration's stage A when no node can move (a moved node changes what the scan reads)
                        const int al_n = TRACE ? 0 : S.al_n2[cb];
                        const bool compact = !TRACE && al_n <= CTA_AL;
                        pending = true;
                        overlap = can_overlap && compact && S.any_moves2[cb] == 0 && it + 1 < p.max_iter;
                        t_status = 3; t_par = best;
#ifdef RRTK_CTA_STATS
                        st_pend++; st_ov += overlap ? 1 : 0; st_big += compact ? 0 : 1; st_moves += S.any_moves2[cb] ? 1 : 0;
#endif
                    } else {
                        if (tid == 0) {
                            const double2 f0 = s_xy[ni];
                            const double nlen = crm_hypot(nx - f0.x, ny - f0.y);
                            const double ncost = cost[ni] + nlen;
                            s_xy[n] = make_double2(nx, ny); s_par[n] = (unsigned short)ni;
                            S.q_xy[n] = make_double2(nx, ny); cost[n] = ncost; S.q_parent[n] = ni; S.q_elen[n] = nlen;
                            cta_link(s_link, ni, n);
                        }
                        t_status = 2; t_par = ni;
                        appended = true;
                    }
                    n++;
                }
            }
#ifdef RRTK_CTA_STATS
            st_cat = (ov ? 0 : 2) + (pending || appended ? 0 : 1);
#endif
#ifdef RRTK_CTA_PROFILE
            if (TRACE && tid == 0) {   // apply of the previous iteration | scan | B1 | first edge | B2 | choose_parent | B3 | rewire edges + B4
                int32_t *tr = trace + (size_t)it * 8;
                tr[0] = (int)(clk[9] - clk[0]); tr[1] = (int)(clk[1] - clk[9]); tr[2] = (int)(clk[2] - clk[1]);
                tr[3] = (int)(clk[3] - clk[2]); tr[4] = (int)(clk[4] - clk[3]); tr[5] = (int)(clk[5] - clk[4]);
                tr[6] = (int)(clk[6] - clk[5]); tr[7] = (int)(clk[8] - clk[6]);
            }
#endif
        }
#ifdef RRTK_CTA_STATS
        if (lane == 0) {
            int32_t *o = trace_all + 32 * (size_t)q;
            if (warp == 0) { o[0] = st_pend; o[1] = st_ov; o[2] = st_big; o[3] = st_moves; }
            for (int k = 0; k < 6; k++) o[4 + 6 * warp + k] = (int)(st_c[k] >> 4);
            if (warp == 0) for (int k = 0; k < 4; k++) { o[28 + k] = st_rn[k] ? (int)(st_r[k] / st_rn[k]) : 0; }
            if (warp == 1) for (int k = 0; k < 4; k++) o[28 + k] += 0;
        }
#endif
        status |= S.status_or;
        if (!done && !RRT_ONLY && warp == 0) {
            bool ovf = false;
            ObsList G = cull_obstacles(obs, n_obs, gx, gy, goal_reach, S.cull[0][0], S.cull[0][1], S.cull[0][2], lane);
            gi = best_goal(plan_consts(p), n, s_xy, cost, gx, gy, G, S.near_idx, S.nd, near_cap, lane, ovf);
            if (ovf) status |= RRTK_Q_NEAR_OVERFLOW;
        }
        if (tid == 0) {
            n_nodes[q] = n;
            iters_done[q] = it_prev + it;
            goal_index[q] = gi;
            status_out[q] = status;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// near-list layout the launch uses: 256 entries (batches: four queries per SM at config 2) or 1024 (what the planner
// classes ask for, the reference having no cap at all); 0 = does not fit
static int cta_layout(const rrtk_rrtstar_params &p) {
    if (p.node_cap > 65535) return 0;
    if (p.near_cap <= CTA_NC_SMALL && cta_smem_bytes<CTA_NC_SMALL>(p.node_cap) <= 227 * 1024) return CTA_NC_SMALL;
    if (!p.rrt_only && p.near_cap <= CTA_NC_LARGE && cta_smem_bytes<CTA_NC_LARGE>(p.node_cap) <= 227 * 1024) return CTA_NC_LARGE;
    return 0;
}
bool rrtstar_cta_fits(const rrtk_rrtstar_params &p) { return cta_layout(p) != 0; }

typedef void (*cta_kernel_t)(rrtk_rrtstar_params, const double4 *, const double4 *, const int32_t *, const double *,
                             const double2 *, const int64_t *, double2 *, double *, int32_t *, int32_t *, int32_t *,
                             int32_t *, int32_t *, int32_t *, int32_t *, unsigned int *);

static cta_kernel_t cta_pick(const rrtk_rrtstar_params &p, int nc, bool trace) {
    if (p.rrt_only)
        return p.resume ? rrtstar_cta_kernel<true, false, true, CTA_NC_SMALL>
                        : (trace ? rrtstar_cta_kernel<true, true, false, CTA_NC_SMALL> : rrtstar_cta_kernel<true, false, false, CTA_NC_SMALL>);
    if (nc == CTA_NC_SMALL)
        return p.resume ? rrtstar_cta_kernel<false, false, true, CTA_NC_SMALL>
                        : (trace ? rrtstar_cta_kernel<false, true, false, CTA_NC_SMALL> : rrtstar_cta_kernel<false, false, false, CTA_NC_SMALL>);
    return p.resume ? rrtstar_cta_kernel<false, false, true, CTA_NC_LARGE>
                    : (trace ? rrtstar_cta_kernel<false, true, false, CTA_NC_LARGE> : rrtstar_cta_kernel<false, false, false, CTA_NC_LARGE>);
}

// persistent grid of the launch = CTAs resident on the device at once (a multiple of the SM count); 0 on error
static int cta_grid(const rrtk_rrtstar_params &p, int nc, cta_kernel_t kern, size_t smem) {
    (void)nc;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CTA_T, smem) != cudaSuccess) return 0;
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

int rrtstar_cta_resident(const rrtk_rrtstar_params &p) {
    const int nc = cta_layout(p);
    if (!nc) return 0;
    const size_t smem = nc == CTA_NC_SMALL ? cta_smem_bytes<CTA_NC_SMALL>(p.node_cap) : cta_smem_bytes<CTA_NC_LARGE>(p.node_cap);
    const int g = cta_grid(p, nc, cta_pick(p, nc, false), smem);
    cudaGetLastError();
    return g;
}

int launch_rrtstar_cta(const rrtk_rrtstar_params &p, const double *start_goal, const double *obstacles,
                       const int32_t *n_obs, const double *near_r2, const double *sample_stream,
                       const int64_t *sobol_offset, double *xy, double *cost, int32_t *parent,
                       int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                       int32_t *trace, int32_t *workspace, unsigned int *counter, cudaStream_t s) {
    const int nc = cta_layout(p);
    if (!nc)
        return set_error(RRTK_ERR_INVALID, "exec_mode = CTA needs node_cap <= 65535, near_cap <= 1024 and 22 B / node + the near list "
                                           "in 227 KB of shared memory");
    const size_t smem = nc == CTA_NC_SMALL ? cta_smem_bytes<CTA_NC_SMALL>(p.node_cap) : cta_smem_bytes<CTA_NC_LARGE>(p.node_cap);
#ifdef RRTK_CTA_STATS
    const cta_kernel_t kern = cta_pick(p, nc, false);   // (the plain kernel writes its pipeline counters into the trace buffer)
#else
    const cta_kernel_t kern = cta_pick(p, nc, trace != nullptr);
#endif
    long long grid = cta_grid(p, nc, kern, smem);
    if (grid < 1) return set_cuda_error(cudaGetLastError(), "cudaFuncSetAttribute / occupancy (rrtstar_cta_kernel)");
    if (grid > p.n_queries) grid = p.n_queries;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(counter)");
    kern<<<(unsigned)grid, CTA_T, smem, s>>>(
        p, reinterpret_cast<const double4 *>(start_goal), reinterpret_cast<const double4 *>(obstacles),
        n_obs, near_r2, reinterpret_cast<const double2 *>(sample_stream), sobol_offset,
        reinterpret_cast<double2 *>(xy), cost, parent, n_nodes, iters_done, goal_index, status, trace,
        workspace, counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "rrtstar_cta_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
