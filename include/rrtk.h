/* rrtk.h -- C-ABI of librrtk.so: B200-native (sm_100a) kernels for the RRT-family hot path of
 * gouldberg/robotics-path-planning.
 *
 * The reference is pure Python and has no FFI; the boundary it exposes is its class API
 * (`RRT(...).planning()`).  This header is what a binding for that path binds:
 * each entry point names the reference function(s) it replaces (alias:line into
 * /root/reference/src_path_planning, aliases in SURVEY.md section 0).  The Python classes in
 * robotics-path-planning_b200/rrtk/ (same names, kwargs and return values as the reference's)
 * call these through ctypes; INTEGRATION.md shows the stub a maintainer of the reference adds.
 *
 * Conventions
 *   - every function returns an int status: 0 = OK, negative = error (rrtk_last_error() gives the
 *     message for the calling thread); nothing throws;
 *   - `*_dev` entry points take CALLER-OWNED DEVICE pointers (e.g. torch `tensor.data_ptr()`),
 *     explicit element counts, and a `cudaStream_t` passed as `void*` (NULL = default stream);
 *     they enqueue work and return without synchronising;
 *   - `*_host` entry points take HOST pointers, allocate temporaries, copy in, run, copy out and
 *     synchronise before returning (the end-to-end path);
 *   - no global mutable state; the current CUDA device is the caller's;
 *   - all floating-point tree state is FP64 (see DESIGN.md "Arithmetic"): the reference computes in
 *     Python floats and its trees are only reproducible bit-for-bit in FP64 with correctly rounded
 *     hypot/atan2/cos/sin (csrc/crmath.h).
 */
#ifndef RRTK_H
#define RRTK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRTK_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define RRTK_API __attribute__((visibility("default")))
#else
#define RRTK_API
#endif

/* status codes */
#define RRTK_OK 0
#define RRTK_ERR_INVALID (-1) /* bad argument */
#define RRTK_ERR_CUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define RRTK_ERR_NO_DEVICE (-3)

/* per-query status bits written by the planners */
#define RRTK_Q_OK 0
#define RRTK_Q_NEAR_OVERFLOW 1 /* |near| exceeded near_cap; the query stopped at that iteration */
#define RRTK_Q_NODE_OVERFLOW 2 /* tree reached node_cap */
#define RRTK_Q_PATH_OVERFLOW 4 /* best path longer than path_cap (length is still reported) */
#define RRTK_Q_DIV_ZERO 8      /* path smoothing: the reference would raise ZeroDivisionError at this iteration */
#define RRTK_Q_NONE_STEER 16   /* RRT-Dubins with a play area: steer returned None and the reference evaluates `node.x` of it
                                  (AttributeError, rrt_03:1626); the query stopped at that iteration */

/* sampler kinds (get_random_node / get_random_node_sobol, rrt_04:1132-1153) */
#define RRTK_SAMPLER_STREAM 0  /* read (x, y) per iteration from `sample_stream` */
#define RRTK_SAMPLER_SOBOL 1   /* in-kernel: goal coin + Sobol point `sobol_offset[q] + #non-goal` */
#define RRTK_SAMPLER_UNIFORM 2 /* in-kernel: goal coin + two counter-based uniforms */

RRTK_API int rrtk_version(void);
RRTK_API const char *rrtk_last_error(void);
/* number of CUDA devices visible, or a negative status */
RRTK_API int rrtk_device_count(void);
/* sizeof of the parameter / result structs as this library was compiled (bindings check their own layout against it):
 * 0 rrtk_rrtstar_params, 1 rrtk_informed_params, 2 rrtk_informed_tree_params, 3 rrtk_informed_tree_result,
 * 4 rrtk_dubins_params, 5 rrtk_closed_loop_params, 6 rrtk_bitstar_params; -1 for an unknown index */
RRTK_API int rrtk_sizeof(int which);

/* ---------------------------------------------------------------------------------------------
 * Sobol generator: i4_sobol (rrt_04:230-503) in closed form (Gray-code order, 30 bits, <= 40 dims,
 * Bratley-Fox direction numbers rrt_04:320-364 + recurrence :393-436).
 *   out[i * dim + d] = coordinate d of point (first_index + i), as FP64 in [0, 1).
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_sobol_fill_dev(int dim, int64_t first_index, int64_t count, double *out_dev, void *stream);
RRTK_API int rrtk_sobol_fill_host(int dim, int64_t first_index, int64_t count, double *out_host);
/* the scaled direction integers V[dim][30] (host memory), for tests */
RRTK_API int rrtk_sobol_table(int dim, uint32_t *v_host);

/* ---------------------------------------------------------------------------------------------
 * Batched RRT* -- the whole `planning()` loop of rrt_04:1036-1084 for Q independent queries in one
 * persistent kernel (one warp per query): get_random_node[_sobol] (:1132-1153),
 * get_nearest_node_index (:1196-1202), steer (:1086-1115), check_if_outside_play_area (:1204-1214),
 * check_collision (:1216-1230), find_near_nodes (:1314-1338, including the `.index()` quirk),
 * choose_parent (:1242-1282), rewire (:1340-1373) + propagate_cost_to_leaves (:1379-1384),
 * search_best_goal_node (:1284-1312).
 * ------------------------------------------------------------------------------------------- */
typedef struct rrtk_rrtstar_params {
    int32_t n_queries;
    int32_t max_iter;
    int32_t node_cap;               /* capacity of each tree, >= max_iter + 1 */
    int32_t obs_stride;             /* obstacles per query in the `obstacles` array (max count) */
    int32_t near_cap;               /* capacity of the near list per query (multiple of 32) */
    int32_t search_until_max_iter;  /* 0: stop at the first goal connection (script default) */
    int32_t sampler;                /* RRTK_SAMPLER_* */
    int32_t goal_sample_rate;       /* percent, `random.randint(0, 100) > rate` -> non-goal */
    int32_t has_play_area;
    int32_t rrt_only;               /* 1: basic RRT loop of rrt_01:71-101 (no near/choose/rewire) */
    double expand_dis;
    double path_resolution;
    double min_rand, max_rand;
    double play_area[4];            /* xmin, xmax, ymin, ymax */
    uint64_t seed;                  /* counter-based RNG key for the in-kernel samplers */
    /* optional uniform cell grid over the obstacles of each query (0 x 0 = off): the per-iteration obstacle cull
     * then reads one cell's list instead of scanning all circles.  The grid must cover every point a node can
     * take (samples, starts, goals); nodes outside it fall back to the full scan, so results never depend on it. */
    int32_t grid_nx, grid_ny;       /* cells, each <= 64 */
    double grid_x0, grid_y0;        /* lower-left corner */
    double grid_cell;               /* cell edge */
    /* incremental planning: resume = 1 continues the trees a previous call left in xy / cost / parent / n_nodes for
     * max_iter MORE iterations; iter_offset = iterations already done (keeps the in-kernel coin / Sobol streams going;
     * with RRTK_SAMPLER_STREAM the caller simply passes the next max_iter samples).  k calls of m iterations build the
     * tree of one call of k * m iterations, bit for bit. */
    int32_t resume, iter_offset;
    /* upper bound of the near radius sqrt(near_r2[k]) over the table.  0 = expand_dis, which is what the reference's
     * clip guarantees (rrt_04:1333-1335).  MANDATORY (> expand_dis) when the caller passes an unclipped table: the
     * per-iteration obstacle cull and the cell grid keep the circles within max(expand_dis, near_r_max) +
     * path_resolution + R of the new node, and choose_parent / rewire edges are tested against that list only. */
    double near_r_max;
    /* how a query is executed (results are bit-identical): RRTK_EXEC_WARP = one warp per query, tree in L2;
     * RRTK_EXEC_CTA = one CTA of 4 warps per query, tree (positions, children lists, parents) in shared memory --
     * needs node_cap <= 65535 and 22 B / node + the near list in <= 227 KB; RRTK_EXEC_AUTO picks CTA when it fits
     * and the whole batch is resident at once (n_queries <= SMs x CTAs per SM: a latency-bound launch), WARP otherwise. */
    int32_t exec_mode;
    /* global index of this launch's query 0: the in-kernel samplers key their counter-based RNG by (seed, query_base + q,
     * iteration), so a shard of a larger batch draws exactly what the unsharded batch draws for the same queries */
    int32_t query_base;
} rrtk_rrtstar_params;

#define RRTK_EXEC_AUTO 0
#define RRTK_EXEC_WARP 1
#define RRTK_EXEC_CTA 2

/* every planner workspace ends with this many extra int32 (the work-queue counter of the persistent grid lives there:
 * the library allocates nothing per call) */
#define RRTK_WS_TAIL_INTS 4

/* ints of workspace per query for rrtk_rrtstar_run_dev */
#define RRTK_RRTSTAR_WS_INTS(node_cap, grid_nx, grid_ny) \
    (4 * (size_t)(node_cap) + 4 * (((size_t)(node_cap) + 1) / 2) + \
     4 * ((17 * (size_t)(grid_nx) * (size_t)(grid_ny) + 3) / 4)) /* three parts, each a multiple of 4: 16-byte rows */

/* Device-pointer entry point.
 *   start_goal   [Q][4]                 sx, sy, gx, gy
 *   obstacles    [Q][obs_stride][4]     x, y, R = size + robot_radius, R2 = (size + robot_radius)**2
 *                                        (R2 is computed by the host exactly as rrt_04:1227 does)
 *   n_obs        [Q]
 *   near_r2      [node_cap + 2]         near_r2[k] = min(ccd*sqrt(log(k)/k), expand_dis)**2, k = #nodes+1
 *                                        (rrt_04:1329-1335; host-evaluated, same libm as the reference)
 *   sample_stream[Q][max_iter][2] or NULL   (RRTK_SAMPLER_STREAM)
 *   sobol_offset [Q] or NULL                first Sobol index of each query (RRTK_SAMPLER_SOBOL)
 * outputs (caller allocated):
 *   xy [Q][node_cap][2], cost [Q][node_cap], parent [Q][node_cap] (-1 = root),
 *   n_nodes [Q], iters_done [Q], goal_index [Q] (-1 = no path), status [Q],
 *   trace [Q][max_iter][8] or NULL: nearest, status, n_near, parent, cp_ok, rw_ok, rw_applied, n_after
 * scratch (caller allocated, contents undefined afterwards):
 *   workspace [Q][RRTK_RRTSTAR_WS_INTS(node_cap, grid_nx, grid_ny)] + [RRTK_WS_TAIL_INTS] int32: children lists (first
 *             child, next / previous sibling) and the breadth-first frontier of propagate_cost_to_leaves, the cached edge
 *             lengths hypot(node - parent) (doubles), the obstacle cell lists (per cell a count + 32 uint16 indices),
 *             then the work-queue counter; 16-byte aligned
 */
RRTK_API int rrtk_rrtstar_run_dev(const rrtk_rrtstar_params *p, const double *start_goal,
                         const double *obstacles, const int32_t *n_obs, const double *near_r2,
                         const double *sample_stream, const int64_t *sobol_offset, double *xy,
                         double *cost, int32_t *parent, int32_t *n_nodes, int32_t *iters_done,
                         int32_t *goal_index, int32_t *status, int32_t *trace, int32_t *workspace,
                         void *stream);

/* Same with HOST pointers (allocates, copies in, runs, copies out, synchronises). */
RRTK_API int rrtk_rrtstar_run_host(const rrtk_rrtstar_params *p, const double *start_goal,
                          const double *obstacles, const int32_t *n_obs, const double *near_r2,
                          const double *sample_stream, const int64_t *sobol_offset, double *xy,
                          double *cost, int32_t *parent, int32_t *n_nodes, int32_t *iters_done,
                          int32_t *goal_index, int32_t *status, int32_t *trace);

/* The per-step primitive: steer (rrt_04:1086-1115, with calc_distance_and_angle :1232-1238) + check_collision (:1216-1230)
 * + check_if_outside_play_area (:1204-1214) for N independent edges -- what an overridden `steer` / `check_collision`
 * pair of a subclass calls; the planner kernels run the same device functions inside their iterations.
 *   from_xy, to_xy [N][2]; extend_length (+inf = the reference's default); obstacles: rows x, y, R, R**2 in sets of
 *   obs_stride rows, request r tests set obs_set[r] (NULL = set 0) with n_obs[set] rows; play_area = xmin, xmax, ymin,
 *   ymax on the device, or NULL.
 * outputs: new_xy [N][2] (the new node), dist [N] = hypot(to - from), n_points [N] = len(path_x), free_flag [N] (1 = no
 *   path point inside a circle), inside_flag [N] (1 = the new node is inside the play area, or there is none). */
RRTK_API int rrtk_steer_collide_dev(int64_t n_req, const double *from_xy, const double *to_xy, double extend_length,
                                    double path_resolution, const int32_t *obs_set, const double *obstacles,
                                    int32_t obs_stride, const int32_t *n_obs, const double *play_area, double *new_xy,
                                    double *dist, int32_t *n_points, uint8_t *free_flag, uint8_t *inside_flag, void *stream);

/* The other per-step methods of the reference's planner classes as stand-alone calls (rrtk.RRT / rrtk.RRTStar route
 * their `steer`, `check_collision`, `get_nearest_node_index` and `find_near_nodes` methods here; the planning loop
 * itself is rrtk_rrtstar_run_dev).  FP64, the reference's operation order.
 *   rrtk_steer_points_dev: the points steer (rrt_04:1086-1115) appends to path_x / path_y for N edges.  extend_length
 *     [N] per edge, or NULL = extend_all for every edge (+inf = the reference's default); points [N][pt_cap][2],
 *     n_points [N] = len(path_x) (points beyond pt_cap are not written: size pt_cap >= floor(extend / res) + 2).
 *   rrtk_points_collide_dev: check_collision (rrt_04:1216-1230) of N point lists (a node's path_x / path_y):
 *     free_flag [N] = 1 when no point lies within (size + robot_radius) of a circle.  Obstacle rows / sets as in
 *     rrtk_steer_collide_dev.
 *   rrtk_nearest_f64_dev: get_nearest_node_index (rrt_04:1196-1202) of B samples [B][2] over n nodes [n][2]:
 *     idx [B] = dlist.index(min(dlist)) (first minimum), d2 [B] (optional) = that squared distance.
 *   rrtk_near_f64_dev: find_near_nodes (rrt_04:1314-1338) around (cx, cy) with squared radius r2: out_idx[0 ..
 *     min(*out_n, cap)) in the reference's order (ascending, each hit replaced by the first node with the same squared
 *     distance -- `dist_list.index(i)`); *out_n = the number of hits; scratch_d2 [cap] doubles of device memory. */
RRTK_API int rrtk_steer_points_dev(int64_t n_req, const double *from_xy, const double *to_xy, const double *extend_length,
                                   double extend_all, double path_resolution, int32_t pt_cap, double *points,
                                   int32_t *n_points, void *stream);
RRTK_API int rrtk_points_collide_dev(int32_t n_req, const double *points, const int32_t *n_points, int32_t pt_cap,
                                     const int32_t *obs_set, const double *obstacles, int32_t obs_stride,
                                     const int32_t *n_obs, uint8_t *free_flag, void *stream);
RRTK_API int rrtk_nearest_f64_dev(const double *xy, int64_t n, const double *samples, int32_t n_samples, int32_t *idx,
                                  double *d2, void *stream);
RRTK_API int rrtk_near_f64_dev(const double *xy, int32_t n, double cx, double cy, double r2, int32_t *out_idx,
                               double *scratch_d2, int32_t cap, int32_t *out_n, void *stream);

/* generate_final_course (rrt_04:1117-1125) for every query on the device:
 *   path [Q][path_cap][2] = goal, node(goal_index), ..., root;  path_len [Q] (0 = no path) */
RRTK_API int rrtk_extract_paths_dev(int32_t n_queries, int32_t node_cap, int32_t path_cap,
                           const double *start_goal, const double *xy, const int32_t *parent,
                           const int32_t *goal_index, double *path, int32_t *path_len, void *stream);

/* path_smoothing (rrt_04:1447-1479, with get_path_length :1390-1398, get_target_point :1401-1420 and
 * line_collision_check :1423-1444) for Q paths at once, in place: random shortcutting of the final course.
 *   path [Q][path_cap][2] in/out (goal -> start as generate_final_course returns it), path_len [Q] in/out (<= 512);
 *   draws [Q][max_iter][2]: the unit uniforms behind the two `random.uniform(0, le)` calls of each iteration
 *   (CPython evaluates uniform(a, b) as a + (b - a) * random());
 *   obstacles3 [Q][obs_stride][3] = x, y, size (no robot radius: the reference passes obstacle_list as is); n_obs [Q];
 *   status [Q]: RRTK_Q_PATH_OVERFLOW / RRTK_Q_DIV_ZERO bits; iters_done [Q].  A path can grow by one point per
 *   accepted shortcut, so path_cap >= path_len + max_iter is always enough. */
RRTK_API int rrtk_path_smoothing_dev(int32_t n_queries, int32_t path_cap, int32_t max_iter, double *path,
                                     int32_t *path_len, const double *draws, const double *obstacles3,
                                     int32_t obs_stride, const int32_t *n_obs, int32_t *status, int32_t *iters_done,
                                     void *stream);

/* The in-kernel samplers, exposed so a sample stream can be materialised (tests, CPU baseline):
 *   out [Q][max_iter][2]; uses p->sampler, seed, goal_sample_rate, min/max_rand, start_goal */
RRTK_API int rrtk_sample_stream_dev(const rrtk_rrtstar_params *p, const double *start_goal,
                           const int64_t *sobol_offset, double *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Batched Informed RRT* -- `informed_rrt_star_search` of rrt_07:1044-1108 for Q independent queries:
 * informed_sample (:1145-1159), get_nearest_list_index (:1210-1214), get_new_node (:1216-1224),
 * check_collision / check_segment_collision / distance_squared_point_to_segment (:1249-1276),
 * find_near_nodes (:1137-1143), choose_parent (:1110-1135), rewire (:1232-1246), is_near_goal,
 * get_final_course, get_path_len and the c_best bookkeeping (:1094-1103).
 * ------------------------------------------------------------------------------------------- */
typedef struct rrtk_informed_params {
    int32_t n_queries;
    int32_t max_iter;
    int32_t node_cap;    /* >= max_iter + 1 */
    int32_t obs_stride;
    int32_t path_cap;    /* capacity of each best-path snapshot */
    int32_t exec_mode;   /* RRTK_EXEC_AUTO (0) / _WARP / _CTA: a warp or a CTA of 4 warps per query (same trees); AUTO takes
                            the CTA while the whole batch is resident at once */
    double expand_dis;
    double coord_bound;  /* upper bound on |coordinate| of anything in the scenes (samples, nodes, circles): sets the
                            tolerance band inside which a near edge falls back to the reference's exact end-point
                            arithmetic; results do not depend on it */
} rrtk_informed_params;

/*   start_goal [Q][4]; rot [Q][4] = c00, c01, c10, c11 of the rotation matrix C (rrt_07:1063-1068, host SVD);
 *   obstacles [Q][obs_stride][4] = x, y, size, size**2 (host-evaluated square, rrt_07:1267); n_obs [Q];
 *   near_rr2 [node_cap + 1][2] = (r, r**2), r = 50*sqrt(log(n)/n), indexed by n = len(node_list) (:1139);
 *   free_samples [Q][max_iter][2] = what sample_free_space[_sobol] returns in iteration i (:1173-1191);
 *   ball_draws [Q][max_iter][2]   = the two random.random() draws of sample_unit_ball (:1162-1171);
 * outputs: xy [Q][node_cap][2], cost, parent [Q][node_cap], n_nodes [Q],
 *   path [Q][path_cap][2] + path_len [Q] (0 = None): snapshot of the best path, goal -> start; c_best [Q];
 *   status [Q];  scratch: ws_idx [Q][node_cap] + [RRTK_WS_TAIL_INTS] int32, ws_d [Q][node_cap] double */
RRTK_API int rrtk_informed_run_dev(const rrtk_informed_params *p, const double *start_goal, const double *rot,
                                   const double *obstacles, const int32_t *n_obs, const double *near_rr2,
                                   const double *free_samples, const double *ball_draws, double *xy,
                                   double *cost, int32_t *parent, int32_t *n_nodes, double *path,
                                   int32_t *path_len, double *c_best, int32_t *status, int32_t *ws_idx,
                                   double *ws_d, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Informed RRT* on ONE large tree (BASELINE config 3: grow to ~10^6 nodes): the same search as
 * rrtk_informed_run_dev (rrt_07:1044-1108), bit-identical results, but the whole GPU works on a single query:
 * get_nearest_list_index (:1210-1214) and find_near_nodes (:1137-1143) stream the FP64 node array (16 B / node)
 * across all SMs in one fused pass per iteration, choose_parent (:1110-1135) and rewire (:1232-1246) are
 * spread over the SMs that own the near nodes; one grid-wide reduction per iteration (persistent cooperative
 * kernel, one CTA per SM).
 *   obstacles [n_obs][4] = x, y, size, size**2 (n_obs <= 512);  near_rr2 [node_cap + 1][2] = (r, r**2) as above;
 *   free_samples, ball_draws [max_iter][2] as for rrtk_informed_run_dev;
 *   coord_bound: an upper bound on |coordinate| of anything in the scene (sets the tolerance band inside which a
 *     near edge falls back to the reference's exact end-point arithmetic; results do not depend on it);
 *   grid: 0 = one CTA per SM; tests pass small values to exercise the multi-pass scans on small trees.
 *   In batched mode result.reextends counts the batches cut short, result.cycles[0] the batches run.
 * outputs: xy [node_cap][2], cost [node_cap], parent [node_cap], path [path_cap][2] (best-path snapshot),
 *   result (device struct).  workspace: rrtk_informed_tree_workspace_bytes(node_cap) bytes of device memory.
 * ------------------------------------------------------------------------------------------- */
typedef struct rrtk_informed_tree_params {
    int32_t max_iter;
    int32_t node_cap;
    int32_t n_obs;
    int32_t path_cap;
    int32_t grid;
    int32_t batch;        /* samples per pass: 0 / 1 = one (speculative-extension kernel); 2..8 = batched kernel, which
                             advances up to `batch` iterations per pair of tree scans and cuts a batch short wherever the
                             sequential semantics would differ (results are identical for every value) */
    double expand_dis;
    double start_goal[4]; /* sx, sy, gx, gy */
    double rot[4];        /* c00, c01, c10, c11 of C (rrt_07:1063-1068) */
    double coord_bound;
} rrtk_informed_tree_params;

typedef struct rrtk_informed_tree_result {
    int32_t n_nodes, path_len, status, iters_done;
    double c_best;
    int64_t total_hits;   /* sum over iterations of len(near_inds) */
    int32_t slow_paths;   /* iterations that needed the exact equal-d^2 resolution */
    int32_t goal_events;  /* iterations whose new node connected to the goal */
    int32_t resamples;    /* iterations after which c_best changed (next sample and nearest redone) */
    int32_t grid;         /* CTAs used */
    int32_t reextends;    /* iterations whose new node was itself the nearest of the next sample */
    int32_t pad_;
    int64_t cycles[6];    /* SM clock cycles CTA 0 spent in each phase (cull, scan, candidate extension + hits,
                             exchange, append + rewire, goal / redo); diagnostics for profiles/ */
    int64_t cycles_max[6];    /* the same, maximum over the CTAs */
    int64_t cycles_negmin[6]; /* minus the minimum over the CTAs */
} rrtk_informed_tree_result;

RRTK_API int64_t rrtk_informed_tree_workspace_bytes(int32_t node_cap, int32_t grid);
RRTK_API int rrtk_informed_tree_run_dev(const rrtk_informed_tree_params *p, const double *obstacles,
                                        const double *near_rr2, const double *free_samples,
                                        const double *ball_draws, double *xy, double *cost, int32_t *parent,
                                        double *path, rrtk_informed_tree_result *result, void *workspace,
                                        int64_t workspace_bytes, void *stream);

/* Diagnostic: the kernel's grid-wide exchange (one 128-byte record per CTA, all-to-all) run `iters` times on its
 * own; cycles_per_exchange [grid] (device) gets each CTA's average SM cycles per exchange.  workspace as above. */
RRTK_API int rrtk_tree_exchange_probe_dev(int32_t grid, int32_t iters, int64_t *cycles_per_exchange, void *workspace,
                                          int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Batched Dubins steering: plan_dubins_path (rrt_05:1021-1109 == dub00) + the sampled collision test of
 * the course (check_collision rrt_05:1625-1638) -- one edge of RRT*-Dubins' `steer` (rrt_05:1458-1479).
 *   from3, to3 [N][3] = x, y, yaw;  curvature, step_size (0.1 in the reference)
 *   obs_set [N] or NULL (all requests use set 0): obstacle set of each request
 *   obstacles [S][obs_stride][4] = x, y, size + robot_radius, (size + robot_radius)**2 ; n_obs [S] (or NULL: none)
 * outputs: mode [N] (0..5 = LSL,RSR,LSR,RSL,RLR,LRL in _PATH_TYPE_MAP order, -1 = none),
 *   lengths [N][3] (segment lengths / curvature), end [N][3] (last course point x, y, yaw),
 *   n_pts [N] (len(px); steer returns None when <= 1), free [N] uint8 (1 = no course point inside a circle),
 *   pts [N][max_pts][3] or NULL (course points x, y, yaw in order)
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_dubins_steer_dev(int32_t n_req, double curvature, double step_size, const double *from3,
                                   const double *to3, const int32_t *obs_set, const double *obstacles,
                                   int32_t obs_stride, const int32_t *n_obs, int32_t *mode, double *lengths,
                                   double *end, int32_t *n_pts, uint8_t *free_flag, double *pts,
                                   int32_t max_pts, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Batched Reeds-Shepp steering: reeds_shepp_path_planning (rs00:496-515 == rrt_06:1426-1437) -- the 12 path functions
 * (rs00:166-363) under the 4 symmetries of generate_path (:366-428), set_path's de-duplication (:141-160), the first
 * shortest path and its sampled course (:431-493) -- plus the sampled collision test of the course, one edge of
 * RRT*-Reeds-Shepp's steer (rrt_06:1584-1606).
 *   from3, to3 [N][3] = x, y, yaw;  maxc = max curvature, step_size (0.2 default in the reference)
 *   obs_set / obstacles / obs_stride / n_obs as for rrtk_dubins_steer_dev
 * outputs: types [N][5] (0 = L, 1 = S, 2 = R, -1 = unused; all -1 when the reference returns None), lengths [N][5]
 *   (signed, / maxc), L [N] (total length / maxc), n_paths [N] = len(paths), end [N][3] (last course point),
 *   n_pts [N] = len(x), free [N], pts [N][max_pts][4] or NULL = x, y, yaw, direction (+1 / -1)
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_reeds_shepp_steer_dev(int32_t n_req, double maxc, double step_size, const double *from3,
                                        const double *to3, const int32_t *obs_set, const double *obstacles,
                                        int32_t obs_stride, const int32_t *n_obs, int32_t *types, double *lengths,
                                        double *L, int32_t *n_paths, double *end, int32_t *n_pts, uint8_t *free_flag,
                                        double *pts, int32_t max_pts, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Batched RRT*-Dubins -- `planning()` of rrt_05:1416-1456 for Q independent queries: nearest on xy
 * (:1605-1610), steer = full Dubins course (:1458-1479), check_collision over the course (:1625-1638),
 * find_near_nodes (:1715-1739), choose_parent (:1648-1689), rewire + propagate with the EUCLIDEAN
 * calc_new_cost (:1741-1779), search_best_goal_node (:1691-1712).
 * ------------------------------------------------------------------------------------------- */
typedef struct rrtk_dubins_params {
    int32_t n_queries;
    int32_t max_iter;
    int32_t node_cap;               /* >= max_iter + 1 */
    int32_t obs_stride;
    int32_t near_cap;               /* multiple of 32 */
    int32_t search_until_max_iter;  /* the `planning(search_until_max_iter=True)` argument (rrt_05:1416) */
    double curvature;
    double step_size;               /* 0.1 in the reference (rrt_05:1022) */
    double goal_xy_th, goal_yaw_th;
    int32_t rs_cost;                /* rrtk_rrtstar_rs_run_dev only.  0: rrt_06 (Euclidean costs, the later calc_new_cost wins);
                                     * 1: rrt_10:1005-1207 -- choose_parent / rewire / propagate_cost_to_leaves cost an edge by
                                     * its Reeds-Shepp length (calc_new_cost :1153-1161) */
    int32_t exec_mode;              /* the RRT* planners (rrtk_rrtstar_dubins_run_dev, rrtk_rrtstar_rs_run_dev): RRTK_EXEC_WARP =
                                     * one warp per query, RRTK_EXEC_CTA = one CTA of 4 warps per query (the candidates of
                                     * choose_parent / rewire spread over 128 threads), RRTK_EXEC_AUTO (0) = CTA while the batch
                                     * is at most one wave of CTAs.  Results are bit-identical. */
} rrtk_dubins_params;

/*   start_goal6 [Q][6] = sx, sy, syaw, gx, gy, gyaw;  obstacles [Q][obs_stride][4] = x, y, size + rr, (size + rr)**2;
 *   n_obs [Q];  near_r2 [node_cap + 2] as for rrtk_rrtstar_run_dev;  stream3 [Q][max_iter][3] = (x, y, yaw) samples
 *   (get_random_node rrt_05:1528-1538)
 * outputs: xy [Q][node_cap][2], yaw, cost, parent [Q][node_cap]; edge_from / edge_to [Q][node_cap][3]: the pose pair
 *   whose Dubins course is the node's path_x / path_y / path_yaw (rrtk_dubins_steer_dev regenerates it);
 *   n_nodes, iters_done, goal_index (-1 = none; index 0 counts as none like the reference), status [Q];
 *   scratch workspace [Q][4][node_cap] + [RRTK_WS_TAIL_INTS] int32, 16-byte aligned (children lists + propagation
 *   frontier, then the work-queue counter) */
RRTK_API int rrtk_rrtstar_dubins_run_dev(const rrtk_dubins_params *p, const double *start_goal6,
                                         const double *obstacles, const int32_t *n_obs, const double *near_r2,
                                         const double *stream3, double *xy, double *yaw, double *cost,
                                         int32_t *parent, double *edge_from, double *edge_to, int32_t *n_nodes,
                                         int32_t *iters_done, int32_t *goal_index, int32_t *status,
                                         int32_t *workspace, void *stream);

/* Batched RRT-Dubins -- `planning()` of rrt_03:1402-1456 for Q independent queries: plain RRT whose steer is the whole Dubins
 * course from the nearest node (xy metric, :1612-1618) to the sample (:1458-1479).  The new node is kept iff its end pose is
 * inside the play area (check_if_outside_play_area :1621-1631) and no course point is inside a circle (check_collision
 * :1634-1648); its cost is the parent's plus `sum([abs(c) for c in course_lengths])` (:1470 -- CPython >= 3.12 evaluates that
 * sum with Neumaier's compensated addition, restated here).  No near / choose_parent / rewire.  search_best_goal_node
 * (:1491-1512) as for RRT*-Dubins; with search_until_max_iter = 0 it runs after every iteration whose steer returned a node.
 * p->near_cap is ignored.  play_area: 4 doubles xmin, xmax, ymin, ymax on the device (one box for the launch) or NULL;
 * with a play area a steer that returns None sets RRTK_Q_NONE_STEER (the reference raises AttributeError there).
 * Other arguments as for rrtk_rrtstar_dubins_run_dev; workspace [RRTK_WS_TAIL_INTS] int32 (the work-queue counter). */
RRTK_API int rrtk_rrt_dubins_run_dev(const rrtk_dubins_params *p, const double *start_goal6, const double *obstacles,
                                     const int32_t *n_obs, const double *play_area, const double *stream3, double *xy,
                                     double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                                     int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                                     int32_t *workspace, void *stream);

/* Batched RRT*-Reeds-Shepp -- `planning()` of rrt_06:1530-1570 for Q independent queries: the same loop as RRT*-Dubins
 * with reeds_shepp_path_planning as steer (:1584-1604; p->step_size is its step_size, 0.2 by default in the reference),
 * a sampler without goal bias (:1658-1666, the caller's stream), and try_goal_path after every append (:1572-1582), which
 * can add a second node per iteration: node_cap >= 2 * max_iter + 1.  Arguments as for rrtk_rrtstar_dubins_run_dev;
 * edge_from / edge_to regenerate a node's course with rrtk_reeds_shepp_steer_dev.
 * p->rs_cost = 1 is `RRTStarReedsShepp.planning` of rrt_10:1050-1090, the planner under ClosedLoopRRTStar: same loop,
 * Reeds-Shepp-length costs; its find_near_nodes does not clip the radius (rrt_10:521-523), so the caller's near_r2 table
 * is (ccd * sqrt(log(k) / k))**2. */
RRTK_API int rrtk_rrtstar_rs_run_dev(const rrtk_dubins_params *p, const double *start_goal6, const double *obstacles,
                                     const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                                     double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                                     int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                                     int32_t *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Closed-loop RRT* feasibility filter -- check_tracking_path_is_feasible of rrt_10:1521-1559 for P candidate courses:
 * extend_path (:1432-1447), calc_speed_profile / set_stop_point (:1375-1429), closed_loop_prediction (:1307-1372, unicycle
 * `update` :1224-1232 + pure_pursuit_control :1255-1283 + PIDControl :1243-1252; constants :1592-1607) and the four
 * checks (goal reached, final angle, travel ratio, collision of the tracked trajectory).
 *   course     [P][course_cap][3]  x, y, yaw in DRIVING order (start -> goal: generate_final_course reversed), n_course [P] >= 3
 *   obstacles  rows x, y, size + rr, (size + rr)**2;  course k tests rows obs_offset[k] .. obs_offset[k] + n_obs[k] - 1
 *   work       [P][course_cap + 6][4] doubles of scratch (the extended course and its speed profile)
 *   traj       [P][traj_cap][7]    x, y, yaw, v, t, a, d of the prediction (traj_cap >= 2003 always suffices: T / dt steps)
 *   n_traj [P], bits [P]: RRTK_CL_* reject reasons, 0 = feasible.
 * search_best_feasible_path (:1494-1519) is then: the LAST course with bits == 0 and minimal traj[n_traj - 1][4].
 * ------------------------------------------------------------------------------------------- */
#define RRTK_CL_NOT_REACHED 1
#define RRTK_CL_BAD_ANGLE 2
#define RRTK_CL_TOO_LONG 4
#define RRTK_CL_COLLISION 8
#define RRTK_CL_TRAJ_OVERFLOW 16
typedef struct rrtk_closed_loop_params {
    int32_t n_courses, course_cap, traj_cap, pad_;
    double target_speed;            /* 10 / 3.6 in the script */
    double yaw_th;                  /* the final-angle check is |yaw - goal yaw| >= 10 * yaw_th (:1538) */
    double invalid_travel_ratio;
} rrtk_closed_loop_params;
RRTK_API int rrtk_closed_loop_dev(const rrtk_closed_loop_params *p, const double *course, const int32_t *n_course,
                                  const double *obstacles, const int32_t *obs_offset, const int32_t *n_obs, double *work,
                                  double *traj, int32_t *n_traj, int32_t *bits, void *stream);

/* ---------------------------------------------------------------------------------------------
 * BIT* -- `BITStar.plan` of rrt_08:236-331 (+ setup_planning :186-216, setup_sample :218-234, informed_sample :385-419,
 * the queue scoring :439-474, expand_vertex :476-501, connect / _collision_check :359-383, update_graph :524-552,
 * remove_queue :343-351, find_final_path :333-341, and RTree's id <-> coordinate maps :65-135) for Q independent queries.
 *   start_goal [Q][4]; rot [Q][4] = the 2 x 2 block of C (:205-213, host-evaluated); obstacles [Q][obs_stride][4] = x, y,
 *   size, size**2; n_obs [Q]; draws [Q][n_draws] = the unit draws `random.random()` returns, in the reference's order of
 *   consumption (`random.uniform(a, b)` is a + (b - a) * draw); num_cells = np.ceil((randArea[1] - randArea[0]) / 0.01).
 * outputs: path [Q][path_cap][2] start -> goal (find_final_path), counts [Q][12] = vertices, edges, parent entries, sample
 *   slots, vertex queue, edge queue, path length (0 = "cannot find Path"), draws used, batches, "Nothing good" resets,
 *   skipped edges, vertex expansions; g_goal [Q] = g_scores[goalId]; status [Q] = RRTK_BIT_* bits (INDEX_ERROR is where
 *   the reference raises IndexError: both queues ran empty inside the expansion loop).
 * workspace: ws_d [Q][RRTK_BITSTAR_WS_DOUBLES], ws_i [Q][RRTK_BITSTAR_WS_INTS] -- the ordered containers of the run (samples,
 *   score table, tree, queues), left in place for inspection (layout: csrc/rrtk_bitstar.cu).
 * ------------------------------------------------------------------------------------------- */
#define RRTK_BIT_SAMPLE_OVERFLOW 1
#define RRTK_BIT_EDGE_OVERFLOW 2
#define RRTK_BIT_VERTEX_OVERFLOW 4
#define RRTK_BIT_DRAWS_EXHAUSTED 8
#define RRTK_BIT_INDEX_ERROR 16
#define RRTK_BIT_PATH_OVERFLOW 32
#define RRTK_BIT_LIVELOCK 64        /* every edge of the first batch was skipped: `iterations` stays 0, the reference loops forever */
typedef struct rrtk_bitstar_params {
    int32_t n_queries, max_iter;
    int32_t vertex_cap;             /* >= max_iter + 1 (one vertex per counted iteration at most) */
    int32_t sample_cap, edge_cap, path_cap, obs_stride, n_draws;
    double min_rand, max_rand, num_cells;
} rrtk_bitstar_params;
#define RRTK_BITSTAR_WS_DOUBLES(vertex_cap, sample_cap, edge_cap) \
    (6 * (size_t)(sample_cap) + 7 * ((size_t)(vertex_cap) + 2) + 3 * (size_t)(edge_cap))
#define RRTK_BITSTAR_WS_INTS(vertex_cap, sample_cap, edge_cap) (11 * ((size_t)(vertex_cap) + 2) + (size_t)(edge_cap))
RRTK_API int rrtk_bitstar_run_dev(const rrtk_bitstar_params *p, const double *start_goal, const double *rot,
                                  const double *obstacles, const int32_t *n_obs, const double *draws, double *ws_d,
                                  int32_t *ws_i, double *path, int32_t *counts, double *g_goal, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Large-tree mode (BASELINE config 3): brute-force searches over an HBM-resident float2 node array.
 *   rrtk_nearest_f32_dev: get_nearest_node_index (rrt_04:1196-1202, rrt_07:1210-1214) for B samples in
 *     one pass over the n nodes; idx[b] = argmin_i |xy[i] - samples[b]|^2 with the LOWEST index on exact
 *     ties, d2[b] = that squared distance (FP32).  8*n bytes of HBM traffic per pass of <= 8 samples.
 *   rrtk_near_f32_dev: find_near_nodes (rrt_04:1314-1338, rrt_07:1137-1143): all i with d2 <= r2, written
 *     unordered to out_idx[0 .. min(*out_n, cap)); *out_n is the total hit count.
 *   xy must be 16-byte aligned; scratch is B uint64 of device memory.
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_nearest_f32_dev(const float *xy, int64_t n, const float *samples, int32_t n_samples,
                                  uint64_t *scratch, int32_t *idx, float *d2, void *stream);
RRTK_API int rrtk_near_f32_dev(const float *xy, int64_t n, float cx, float cy, float r2, int32_t *out_idx,
                               int32_t cap, int32_t *out_n, void *stream);

/* FMA-loop pipe probes (roofline denominators): blocks x 256 threads x 8 chains x iters FMAs.
 * fp64 != 0 -> DFMA, else FFMA.  flops = blocks * 256 * 8 * 2 * iters. */
RRTK_API int rrtk_fma_peak_dev(int fp64, int32_t iters, int32_t blocks, void *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Arm C-space occupancy grid: get_occupancy_grid (arm02:79-110) with NLinkArm.update_points
 * (:257-262, joint k uses theta1 for k = 1 and theta1 + theta2 for k >= 2) and detect_collision (:46-76),
 * for S obstacle sets at once and rows [row0, row0 + n_rows) of the M x M joint grid.
 *   theta        [M]  (device)  the reference's theta_list, evaluated by the host (arm02:95)
 *   link_lengths [n_links] (HOST, n_links <= 16)
 *   obstacles    [S][O][3] (device)  x, y, radius
 *   grid         [S][n_rows][M] uint8 (device)  1 = the arm touches an obstacle of that set
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_arm_grid_dev(int32_t M, const double *theta, int32_t row0, int32_t n_rows, int32_t n_links,
                               const double *link_lengths, const double *obstacles, int32_t n_sets,
                               int32_t n_obs, uint8_t *grid, void *stream);
/* The same grid, every cell evaluated on its own in the reference's order (one thread per cell; rrtk_arm_grid_dev
 * rasterises whole rows and evaluates only the cells it cannot decide).  Same arguments, same result; kept as the
 * cross-check of the rasteriser (tests) -- about 20x slower at M = 8192. */
RRTK_API int rrtk_arm_grid_cells_dev(int32_t M, const double *theta, int32_t row0, int32_t n_rows, int32_t n_links,
                                     const double *link_lengths, const double *obstacles, int32_t n_sets,
                                     int32_t n_obs, uint8_t *grid, void *stream);

/* ---------------------------------------------------------------------------------------------
 * astar_torus (arm02:113-184, with calc_heuristic_map :221-233 and find_neighbors :187-209) for Q queries at once:
 * the reference's greedy best-first search on the M x M joint-space torus, same expansion order (smallest
 * heuristic, ties to the smallest row-major index), same route, same marks left in the grid.
 *   start_goal [Q][4] int32 = start row, start col, goal row, goal col
 *   grids [Q][M][M] uint8 in/out: 0 free, 1 occupied on entry; on return 2 expanded, 3 frontier, 4 start, 5 goal,
 *         6 route (what the reference leaves in `grid`)
 *   routes [Q][route_cap][2] int32: cells from start to goal; route_len [Q] (0 = "No route found", negative = the
 *         route has -route_len cells and did not fit); expanded [Q] = cells expanded
 *   scratch (device): heur [Q][M*M] int32, parents [Q][M*M] int32, heaps [Q][M*M + 8] uint64
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_astar_torus_dev(int32_t M, int32_t n_queries, const int32_t *start_goal, uint8_t *grids,
                                  int32_t *routes, int32_t route_cap, int32_t *route_len, int32_t *expanded,
                                  int32_t *heur, int32_t *parents, uint64_t *heaps, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Leaf-function probes (tests): the correctly-rounded device functions of csrc/crmath.h.
 *   kind 0: out[i] = hypot(a[i], b[i])   kind 1: atan2(a[i], b[i])   kind 2: sin(a[i])
 *   kind 3: cos(a[i])                     kind 4/5: sin/cos(atan2(a[i], b[i])) (fused steer form)
 *   kind 6: acos(a[i])                    kind 7: asin(a[i])                kind 8: tan(a[i])
 * ------------------------------------------------------------------------------------------- */
RRTK_API int rrtk_crmath_probe_dev(int kind, int64_t n, const double *a, const double *b, double *out,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RRTK_H */
