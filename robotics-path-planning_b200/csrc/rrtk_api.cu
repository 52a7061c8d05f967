// rrtk_api.cu -- the extern "C" boundary (include/rrtk.h): argument checks, error strings, host-buffer
// wrappers.  No torch types; plain pointers and sizes.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/rrtk.h"
#include "rrtk_device.cuh"

namespace rrtk {

static thread_local std::string g_err;

int set_error(int code, const char *msg) {
    g_err = msg ? msg : "";
    return code;
}
int set_cuda_error(cudaError_t e, const char *where) {
    g_err = std::string(where ? where : "cuda") + ": " + cudaGetErrorString(e);
    return RRTK_ERR_CUDA;
}

// launchers (defined next to their kernels)
int launch_rrtstar(const rrtk_rrtstar_params &p, const double *start_goal, const double *obstacles,
                   const int32_t *n_obs, const double *near_r2, const double *sample_stream,
                   const int64_t *sobol_offset, double *xy, double *cost, int32_t *parent,
                   int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                   int32_t *trace, int32_t *workspace, unsigned int *counter, cudaStream_t s);
int launch_extract_paths(int32_t nq, int32_t node_cap, int32_t path_cap, const double *start_goal,
                         const double *xy, const int32_t *parent, const int32_t *goal_index,
                         double *path, int32_t *path_len, cudaStream_t s);
int launch_sample_stream(const rrtk_rrtstar_params &p, const double *start_goal,
                         const int64_t *sobol_offset, double *out, cudaStream_t s);
int launch_crmath_probe(int kind, int64_t n, const double *a, const double *b, double *out, cudaStream_t s);
int launch_sobol_fill(int dim, int64_t first, int64_t count, double *out, cudaStream_t s);
int launch_nearest(const float *xy, long long n, const float *samples, int B, unsigned long long *scratch,
                   int *idx, float *d2, cudaStream_t s);
int launch_near(const float *xy, long long n, float cx, float cy, float r2, int *out_idx, int cap, int *out_n,
                cudaStream_t s);
int launch_fma_peak(int fp64, int iters, int blocks, void *out, cudaStream_t s);
int launch_rrtstar_dubins(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles,
                          const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                          double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                          int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                          int32_t *workspace, unsigned int *counter, cudaStream_t s);
int launch_rrtstar_rs(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles,
                      const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                      double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                      int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                      int32_t *workspace, unsigned int *counter, cudaStream_t s);
int launch_rrt_dubins(const rrtk_dubins_params &p, const double *start_goal6, const double *obstacles, const int32_t *n_obs,
                      const double *play, const double *stream3, double *xy, double *yaw, double *cost, int32_t *parent,
                      double *edge_from, double *edge_to, int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index,
                      int32_t *status, unsigned int *counter, cudaStream_t s);
int launch_dubins_steer(int n_req, double kappa, double step, const double *from3, const double *to3,
                        const int32_t *obs_set, const double *obstacles, int obs_stride, const int32_t *n_obs,
                        int32_t *mode, double *lengths, double *end, int32_t *n_pts, uint8_t *free_flag,
                        double *pts, int max_pts, cudaStream_t s);
int launch_informed(const rrtk_informed_params &p, const double *start_goal, const double *rot,
                    const double *obstacles, const int32_t *n_obs, const double *near_rr2, const double *free_s,
                    const double *ball, double *xy, double *cost, int32_t *parent, int32_t *n_nodes, double *path,
                    int32_t *path_len, double *c_best, int32_t *status, int32_t *ws_idx, double *ws_d,
                    unsigned int *counter, cudaStream_t s);
int informed_tree_workspace_bytes(int node_cap, int grid, size_t *bytes);
int launch_informed_tree(const rrtk_informed_tree_params &p, const double *obstacles, const double *near_rr2,
                         const double *free_s, const double *ball, double *xy, double *cost, int32_t *parent,
                         double *path, rrtk_informed_tree_result *res, void *workspace, size_t workspace_bytes,
                         cudaStream_t s);
int launch_tree_exchange_probe(int grid, int iters, long long *out_dev, void *workspace, size_t workspace_bytes, cudaStream_t s);
int launch_closed_loop(const rrtk_closed_loop_params &p, const double *course, const int32_t *n_course, const double *obstacles,
                       const int32_t *obs_offset, const int32_t *n_obs, double *work, double *traj, int32_t *n_traj,
                       int32_t *bits, cudaStream_t s);
int launch_bitstar(const rrtk_bitstar_params &p, const double *start_goal, const double *rot, const double *obstacles,
                   const int32_t *n_obs, const double *draws, double *ws_d, int32_t *ws_i, double *path, int32_t *counts,
                   double *g_goal, int32_t *status, cudaStream_t s);
int launch_steer_collide(long long n_req, const double *from_xy, const double *to_xy, double extend, double res,
                         const int32_t *obs_set, const double *obstacles, int obs_stride, const int32_t *n_obs, const double *play,
                         double *new_xy, double *dist, int32_t *n_points, uint8_t *free_flag, uint8_t *inside_flag, cudaStream_t s);
int launch_steer_points(long long n_req, const double *from_xy, const double *to_xy, const double *extend, double extend_all,
                        double res, int pt_cap, double *points, int32_t *n_points, cudaStream_t s);
int launch_points_collide(int n_req, const double *points, const int32_t *n_points, int pt_cap, const int32_t *obs_set,
                          const double *obstacles, int obs_stride, const int32_t *n_obs, uint8_t *free_flag, cudaStream_t s);
int launch_nearest_f64(const double *xy, long long n, const double *samples, int n_samples, int32_t *out_idx, double *out_d2,
                       cudaStream_t s);
int launch_near_f64(const double *xy, int n, double cx, double cy, double r2, int32_t *out_idx, double *scratch_d2, int cap,
                    int32_t *out_n, cudaStream_t s);
int launch_smooth_paths(int n_queries, int path_cap, int max_iter, double *path, int32_t *path_len, const double *draws,
                        const double *obs3, int obs_stride, const int32_t *n_obs, int32_t *status, int32_t *iters_done,
                        cudaStream_t s);
int launch_astar_torus(int M, int n_queries, const int32_t *start_goal, uint8_t *grids, int32_t *heur, int32_t *parents,
                       unsigned long long *heaps, int32_t *routes, int route_cap, int32_t *route_len, int32_t *expanded,
                       cudaStream_t s);
int launch_rs_steer(int n_req, double maxc, double step_size, const double *from3, const double *to3, const int32_t *obs_set,
                    const double *obstacles, int obs_stride, const int32_t *n_obs, int32_t *types, double *lengths, double *L,
                    int32_t *n_paths, double *end, int32_t *n_pts, uint8_t *free_flag, double *pts, int max_pts, cudaStream_t s);
int launch_arm_grid(int M, const double *theta, int row0, int n_rows, int n_links, const double *link_host,
                    const double *obstacles, int S, int O, uint8_t *grid, cudaStream_t s, int cells_only);

static int check_params(const rrtk_rrtstar_params *p) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0) return set_error(RRTK_ERR_INVALID, "n_queries/max_iter negative");
    if (p->node_cap < 1) return set_error(RRTK_ERR_INVALID, "node_cap < 1");
    if (p->near_cap < 32 || (p->near_cap & 31)) return set_error(RRTK_ERR_INVALID, "near_cap must be a positive multiple of 32");
    if (!(p->path_resolution > 0.0)) return set_error(RRTK_ERR_INVALID, "path_resolution must be > 0");
    if (!(p->expand_dis >= 0.0)) return set_error(RRTK_ERR_INVALID, "expand_dis must be >= 0");
    if (p->sampler < 0 || p->sampler > 2) return set_error(RRTK_ERR_INVALID, "unknown sampler");
    if (p->obs_stride < 0 || p->obs_stride > 65535) return set_error(RRTK_ERR_INVALID, "obs_stride out of range");
    if (p->grid_nx < 0 || p->grid_ny < 0 || p->grid_nx > 64 || p->grid_ny > 64 || ((p->grid_nx == 0) != (p->grid_ny == 0)))
        return set_error(RRTK_ERR_INVALID, "grid_nx / grid_ny must both be 0 or both in 1..64");
    if (p->grid_nx > 0 && !(p->grid_cell > 0.0)) return set_error(RRTK_ERR_INVALID, "grid_cell must be > 0");
    if ((p->resume != 0 && p->resume != 1) || p->iter_offset < 0) return set_error(RRTK_ERR_INVALID, "resume must be 0 / 1, iter_offset >= 0");
    if (!(p->near_r_max >= 0.0)) return set_error(RRTK_ERR_INVALID, "near_r_max must be >= 0 (0 = expand_dis)");
    if (p->exec_mode < RRTK_EXEC_AUTO || p->exec_mode > RRTK_EXEC_CTA) return set_error(RRTK_ERR_INVALID, "unknown exec_mode");
    if (p->query_base < 0) return set_error(RRTK_ERR_INVALID, "query_base must be >= 0");
    return RRTK_OK;
}

// the work-queue counter of a persistent planner grid: the RRTK_WS_TAIL_INTS ints that end the caller's workspace
static inline unsigned int *ws_tail(int32_t *ws, size_t ints_before) {
    return reinterpret_cast<unsigned int *>(ws + ints_before);
}

}  // namespace rrtk

using namespace rrtk;

extern "C" {

int rrtk_version(void) { return RRTK_VERSION; }
const char *rrtk_last_error(void) { return g_err.c_str(); }

int rrtk_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return set_cuda_error(e, "cudaGetDeviceCount"); }
    return n;
}

int rrtk_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(rrtk_rrtstar_params);
        case 1: return (int)sizeof(rrtk_informed_params);
        case 2: return (int)sizeof(rrtk_informed_tree_params);
        case 3: return (int)sizeof(rrtk_informed_tree_result);
        case 4: return (int)sizeof(rrtk_dubins_params);
        case 5: return (int)sizeof(rrtk_closed_loop_params);
        case 6: return (int)sizeof(rrtk_bitstar_params);
        default: return -1;
    }
}

int rrtk_sobol_table(int dim, uint32_t *v_host) {
    if (dim < 1 || dim > SOBOL_DIM_MAX || !v_host) return set_error(RRTK_ERR_INVALID, "1 <= dim <= 40");
    static const SobolTable t = make_sobol_table();
    std::memcpy(v_host, t.v, sizeof(uint32_t) * SOBOL_BITS * dim);
    return RRTK_OK;
}

int rrtk_sobol_fill_dev(int dim, int64_t first_index, int64_t count, double *out_dev, void *stream) {
    if (dim < 1 || dim > SOBOL_DIM_MAX) return set_error(RRTK_ERR_INVALID, "1 <= dim <= 40 (rrt_04:382-387)");
    if (count < 0 || (count > 0 && !out_dev)) return set_error(RRTK_ERR_INVALID, "bad count/out");
    if (count == 0) return RRTK_OK;
    return launch_sobol_fill(dim, first_index, count, out_dev, (cudaStream_t)stream);
}

int rrtk_sobol_fill_host(int dim, int64_t first_index, int64_t count, double *out_host) {
    if (dim < 1 || dim > SOBOL_DIM_MAX) return set_error(RRTK_ERR_INVALID, "1 <= dim <= 40 (rrt_04:382-387)");
    if (count < 0 || (count > 0 && !out_host)) return set_error(RRTK_ERR_INVALID, "bad count/out");
    if (count == 0) return RRTK_OK;
    double *d = nullptr;
    size_t bytes = sizeof(double) * (size_t)count * dim;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc");
    int rc = launch_sobol_fill(dim, first_index, count, d, 0);
    if (rc == RRTK_OK) {
        e = cudaMemcpy(out_host, d, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_cuda_error(e, "cudaMemcpy D2H");
    }
    cudaFree(d);
    return rc;
}

int rrtk_rrtstar_run_dev(const rrtk_rrtstar_params *p, const double *start_goal,
                         const double *obstacles, const int32_t *n_obs, const double *near_r2,
                         const double *sample_stream, const int64_t *sobol_offset, double *xy,
                         double *cost, int32_t *parent, int32_t *n_nodes, int32_t *iters_done,
                         int32_t *goal_index, int32_t *status, int32_t *trace, int32_t *workspace,
                         void *stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal || !n_obs || !xy || !cost || !parent || !n_nodes || !iters_done || !goal_index || !status)
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (!p->rrt_only && !near_r2) return set_error(RRTK_ERR_INVALID, "near_r2 is NULL");
    if (!workspace) return set_error(RRTK_ERR_INVALID, "workspace is NULL (need n_queries * RRTK_RRTSTAR_WS_INTS int32)");
    if ((uintptr_t)workspace & 15) return set_error(RRTK_ERR_INVALID, "workspace must be 16-byte aligned");
    if (p->obs_stride > 0 && !obstacles) return set_error(RRTK_ERR_INVALID, "obstacles is NULL");
    if (p->sampler == RRTK_SAMPLER_STREAM && !sample_stream && p->max_iter > 0)
        return set_error(RRTK_ERR_INVALID, "sampler = STREAM needs sample_stream");
    if (p->resume && trace) return set_error(RRTK_ERR_INVALID, "trace is not recorded with resume = 1");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int *ctr = ws_tail(workspace, (size_t)p->n_queries * RRTK_RRTSTAR_WS_INTS(p->node_cap, p->grid_nx, p->grid_ny));
    return launch_rrtstar(*p, start_goal, obstacles, n_obs, near_r2, sample_stream, sobol_offset, xy, cost,
                          parent, n_nodes, iters_done, goal_index, status, trace, workspace, ctr, s);
}

int rrtk_informed_run_dev(const rrtk_informed_params *p, const double *start_goal, const double *rot,
                          const double *obstacles, const int32_t *n_obs, const double *near_rr2,
                          const double *free_samples, const double *ball_draws, double *xy, double *cost,
                          int32_t *parent, int32_t *n_nodes, double *path, int32_t *path_len, double *c_best,
                          int32_t *status, int32_t *ws_idx, double *ws_d, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0 || p->node_cap < 1 || p->path_cap < 2 || p->obs_stride < 0)
        return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (!(p->expand_dis > 0.0)) return set_error(RRTK_ERR_INVALID, "expand_dis must be > 0");
    if (!(p->coord_bound > 0.0)) return set_error(RRTK_ERR_INVALID, "coord_bound must be > 0");
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal || !rot || !n_obs || !near_rr2 || !xy || !cost || !parent || !n_nodes || !path || !path_len ||
        !c_best || !status || !ws_idx || !ws_d || (p->max_iter > 0 && (!free_samples || !ball_draws)) ||
        (p->obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    return launch_informed(*p, start_goal, rot, obstacles, n_obs, near_rr2, free_samples, ball_draws, xy, cost, parent,
                           n_nodes, path, path_len, c_best, status, ws_idx, ws_d,
                           ws_tail(ws_idx, (size_t)p->n_queries * (size_t)p->node_cap), s);
}

int64_t rrtk_informed_tree_workspace_bytes(int32_t node_cap, int32_t grid) {
    if (node_cap < 1 || grid < 0) return set_error(RRTK_ERR_INVALID, "node_cap < 1 or grid < 0");
    size_t b = 0;
    int rc = informed_tree_workspace_bytes(node_cap, grid, &b);
    return rc ? rc : (int64_t)b;
}

int rrtk_informed_tree_run_dev(const rrtk_informed_tree_params *p, const double *obstacles, const double *near_rr2,
                               const double *free_samples, const double *ball_draws, double *xy, double *cost,
                               int32_t *parent, double *path, rrtk_informed_tree_result *result, void *workspace,
                               int64_t workspace_bytes, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->max_iter < 0 || p->node_cap < 1 || p->path_cap < 2 || p->n_obs < 0 || p->grid < 0)
        return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (p->n_obs > 512) return set_error(RRTK_ERR_INVALID, "n_obs > 512 (shared-memory obstacle stage)");
    if (!(p->expand_dis > 0.0)) return set_error(RRTK_ERR_INVALID, "expand_dis must be > 0");
    if (!(p->coord_bound > 0.0)) return set_error(RRTK_ERR_INVALID, "coord_bound must be > 0");
    if (!near_rr2 || !xy || !cost || !parent || !path || !result || !workspace ||
        (p->max_iter > 0 && (!free_samples || !ball_draws)) || (p->n_obs > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)xy | (uintptr_t)path | (uintptr_t)workspace) & 15)
        return set_error(RRTK_ERR_INVALID, "xy, path and workspace must be 16-byte aligned");
    return launch_informed_tree(*p, obstacles, near_rr2, free_samples, ball_draws, xy, cost, parent, path, result,
                                workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
}

int rrtk_tree_exchange_probe_dev(int32_t grid, int32_t iters, int64_t *cycles_per_exchange, void *workspace,
                                 int64_t workspace_bytes, void *stream) {
    if (grid < 0 || iters < 1 || !cycles_per_exchange || !workspace) return set_error(RRTK_ERR_INVALID, "bad arguments");
    return launch_tree_exchange_probe(grid, iters, (long long *)cycles_per_exchange, workspace, (size_t)workspace_bytes,
                                      (cudaStream_t)stream);
}

int rrtk_dubins_steer_dev(int32_t n_req, double curvature, double step_size, const double *from3,
                          const double *to3, const int32_t *obs_set, const double *obstacles, int32_t obs_stride,
                          const int32_t *n_obs, int32_t *mode, double *lengths, double *end, int32_t *n_pts,
                          uint8_t *free_flag, double *pts, int32_t max_pts, void *stream) {
    if (n_req < 0 || obs_stride < 0 || max_pts < 0) return set_error(RRTK_ERR_INVALID, "negative size");
    if (!(curvature > 0.0) || !(step_size > 0.0)) return set_error(RRTK_ERR_INVALID, "curvature and step_size must be > 0");
    if (n_req == 0) return RRTK_OK;
    if (!from3 || !to3 || !mode || !lengths || !end || !n_pts || !free_flag) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (n_obs && !obstacles) return set_error(RRTK_ERR_INVALID, "n_obs given without obstacles");
    return launch_dubins_steer(n_req, curvature, step_size, from3, to3, obs_set, obstacles, obs_stride, n_obs, mode,
                               lengths, end, n_pts, free_flag, pts, max_pts, (cudaStream_t)stream);
}

int rrtk_reeds_shepp_steer_dev(int32_t n_req, double maxc, double step_size, const double *from3, const double *to3,
                               const int32_t *obs_set, const double *obstacles, int32_t obs_stride, const int32_t *n_obs,
                               int32_t *types, double *lengths, double *L, int32_t *n_paths, double *end, int32_t *n_pts,
                               uint8_t *free_flag, double *pts, int32_t max_pts, void *stream) {
    if (n_req < 0 || obs_stride < 0 || max_pts < 0) return set_error(RRTK_ERR_INVALID, "negative size");
    if (!(maxc > 0.0) || !(step_size > 0.0)) return set_error(RRTK_ERR_INVALID, "maxc and step_size must be > 0");
    if (n_req == 0) return RRTK_OK;
    if (!from3 || !to3 || !types || !lengths || !L || !n_paths || !end || !n_pts || !free_flag)
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (n_obs && !obstacles) return set_error(RRTK_ERR_INVALID, "n_obs given without obstacles");
    return launch_rs_steer(n_req, maxc, step_size, from3, to3, obs_set, obstacles, obs_stride, n_obs, types, lengths, L,
                           n_paths, end, n_pts, free_flag, pts, max_pts, (cudaStream_t)stream);
}

int rrtk_rrtstar_dubins_run_dev(const rrtk_dubins_params *p, const double *start_goal6, const double *obstacles,
                                const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                                double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                                int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                                int32_t *workspace, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0 || p->node_cap < 1 || p->obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (p->near_cap < 32 || (p->near_cap & 31)) return set_error(RRTK_ERR_INVALID, "near_cap must be a positive multiple of 32");
    if (!(p->curvature > 0.0) || !(p->step_size > 0.0)) return set_error(RRTK_ERR_INVALID, "curvature and step_size must be > 0");
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal6 || !n_obs || !near_r2 || !xy || !yaw || !cost || !parent || !edge_from || !edge_to || !n_nodes ||
        !iters_done || !goal_index || !status || !workspace || (p->max_iter > 0 && !stream3) ||
        (p->obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    if ((uintptr_t)workspace & 15) return set_error(RRTK_ERR_INVALID, "workspace must be 16-byte aligned");
    return launch_rrtstar_dubins(*p, start_goal6, obstacles, n_obs, near_r2, stream3, xy, yaw, cost, parent, edge_from,
                                 edge_to, n_nodes, iters_done, goal_index, status, workspace,
                                 ws_tail(workspace, (size_t)p->n_queries * 4 * (size_t)p->node_cap), s);
}

int rrtk_rrt_dubins_run_dev(const rrtk_dubins_params *p, const double *start_goal6, const double *obstacles,
                            const int32_t *n_obs, const double *play_area, const double *stream3, double *xy, double *yaw,
                            double *cost, int32_t *parent, double *edge_from, double *edge_to, int32_t *n_nodes,
                            int32_t *iters_done, int32_t *goal_index, int32_t *status, int32_t *workspace, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0 || p->node_cap < 1 || p->obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (!(p->curvature > 0.0) || !(p->step_size > 0.0)) return set_error(RRTK_ERR_INVALID, "curvature and step_size must be > 0");
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal6 || !n_obs || !xy || !yaw || !cost || !parent || !edge_from || !edge_to || !n_nodes || !iters_done ||
        !goal_index || !status || !workspace || (p->max_iter > 0 && !stream3) || (p->obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)xy | (uintptr_t)obstacles) & 15) return set_error(RRTK_ERR_INVALID, "xy and obstacles must be 16-byte aligned");
    return launch_rrt_dubins(*p, start_goal6, obstacles, n_obs, play_area, stream3, xy, yaw, cost, parent, edge_from, edge_to,
                             n_nodes, iters_done, goal_index, status, ws_tail(workspace, 0), (cudaStream_t)stream);
}

int rrtk_rrtstar_rs_run_dev(const rrtk_dubins_params *p, const double *start_goal6, const double *obstacles,
                                const int32_t *n_obs, const double *near_r2, const double *stream3, double *xy,
                                double *yaw, double *cost, int32_t *parent, double *edge_from, double *edge_to,
                                int32_t *n_nodes, int32_t *iters_done, int32_t *goal_index, int32_t *status,
                                int32_t *workspace, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0 || p->node_cap < 1 || p->obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (p->near_cap < 32 || (p->near_cap & 31)) return set_error(RRTK_ERR_INVALID, "near_cap must be a positive multiple of 32");
    if (!(p->curvature > 0.0) || !(p->step_size > 0.0)) return set_error(RRTK_ERR_INVALID, "curvature and step_size must be > 0");
    if (p->rs_cost != 0 && p->rs_cost != 1) return set_error(RRTK_ERR_INVALID, "rs_cost must be 0 or 1");
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal6 || !n_obs || !near_r2 || !xy || !yaw || !cost || !parent || !edge_from || !edge_to || !n_nodes ||
        !iters_done || !goal_index || !status || !workspace || (p->max_iter > 0 && !stream3) ||
        (p->obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    if ((uintptr_t)workspace & 15) return set_error(RRTK_ERR_INVALID, "workspace must be 16-byte aligned");
    return launch_rrtstar_rs(*p, start_goal6, obstacles, n_obs, near_r2, stream3, xy, yaw, cost, parent, edge_from,
                                 edge_to, n_nodes, iters_done, goal_index, status, workspace,
                                 ws_tail(workspace, (size_t)p->n_queries * 4 * (size_t)p->node_cap), s);
}

int rrtk_extract_paths_dev(int32_t n_queries, int32_t node_cap, int32_t path_cap,
                           const double *start_goal, const double *xy, const int32_t *parent,
                           const int32_t *goal_index, double *path, int32_t *path_len, void *stream) {
    if (n_queries < 0 || node_cap < 1 || path_cap < 2) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (n_queries == 0) return RRTK_OK;
    if (!start_goal || !xy || !parent || !goal_index || !path || !path_len)
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_extract_paths(n_queries, node_cap, path_cap, start_goal, xy, parent, goal_index, path,
                                path_len, (cudaStream_t)stream);
}

int rrtk_path_smoothing_dev(int32_t n_queries, int32_t path_cap, int32_t max_iter, double *path, int32_t *path_len,
                            const double *draws, const double *obstacles3, int32_t obs_stride, const int32_t *n_obs,
                            int32_t *status, int32_t *iters_done, void *stream) {
    if (n_queries < 0 || path_cap < 1 || max_iter < 0 || obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (n_queries == 0) return RRTK_OK;
    if (!path || !path_len || !status || !iters_done || (max_iter > 0 && !draws) || (n_obs && obs_stride > 0 && !obstacles3))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)path & 15) return set_error(RRTK_ERR_INVALID, "path must be 16-byte aligned");
    return launch_smooth_paths(n_queries, path_cap, max_iter, path, path_len, draws, obstacles3, obs_stride, n_obs, status,
                               iters_done, (cudaStream_t)stream);
}

int rrtk_closed_loop_dev(const rrtk_closed_loop_params *p, const double *course, const int32_t *n_course,
                         const double *obstacles, const int32_t *obs_offset, const int32_t *n_obs, double *work, double *traj,
                         int32_t *n_traj, int32_t *bits, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_courses < 0 || p->course_cap < 3 || p->traj_cap < 1) return set_error(RRTK_ERR_INVALID, "bad sizes (course_cap >= 3, traj_cap >= 1)");
    if (p->n_courses == 0) return RRTK_OK;
    if (!course || !n_course || !obs_offset || !n_obs || !work || !traj || !n_traj || !bits)
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)obstacles & 15) return set_error(RRTK_ERR_INVALID, "obstacles must be 16-byte aligned");
    return launch_closed_loop(*p, course, n_course, obstacles, obs_offset, n_obs, work, traj, n_traj, bits, (cudaStream_t)stream);
}

int rrtk_bitstar_run_dev(const rrtk_bitstar_params *p, const double *start_goal, const double *rot, const double *obstacles,
                         const int32_t *n_obs, const double *draws, double *ws_d, int32_t *ws_i, double *path, int32_t *counts,
                         double *g_goal, int32_t *status, void *stream) {
    if (!p) return set_error(RRTK_ERR_INVALID, "params is NULL");
    if (p->n_queries < 0 || p->max_iter < 0 || p->vertex_cap < 2 || p->sample_cap < 1 || p->edge_cap < 1 || p->path_cap < 2 ||
        p->obs_stride < 0 || p->n_draws < 0)
        return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (!(p->num_cells >= 1.0) || !(p->max_rand > p->min_rand)) return set_error(RRTK_ERR_INVALID, "bad randArea / num_cells");
    if (p->n_queries == 0) return RRTK_OK;
    if (!start_goal || !rot || !n_obs || !draws || !ws_d || !ws_i || !path || !counts || !g_goal || !status ||
        (p->obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)obstacles & 15) return set_error(RRTK_ERR_INVALID, "obstacles must be 16-byte aligned");
    return launch_bitstar(*p, start_goal, rot, obstacles, n_obs, draws, ws_d, ws_i, path, counts, g_goal, status,
                          (cudaStream_t)stream);
}

int rrtk_steer_collide_dev(int64_t n_req, const double *from_xy, const double *to_xy, double extend_length, double path_resolution,
                           const int32_t *obs_set, const double *obstacles, int32_t obs_stride, const int32_t *n_obs,
                           const double *play_area, double *new_xy, double *dist, int32_t *n_points, uint8_t *free_flag,
                           uint8_t *inside_flag, void *stream) {
    if (n_req < 0 || obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (!(path_resolution > 0.0) || !(extend_length >= 0.0)) return set_error(RRTK_ERR_INVALID, "path_resolution must be > 0, extend_length >= 0");
    if (n_req == 0) return RRTK_OK;
    if (!from_xy || !to_xy || !new_xy || !dist || !n_points || !free_flag || !inside_flag || (n_obs && obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)from_xy | (uintptr_t)to_xy | (uintptr_t)new_xy | (uintptr_t)obstacles) & 15)
        return set_error(RRTK_ERR_INVALID, "from_xy / to_xy / new_xy / obstacles must be 16-byte aligned");
    return launch_steer_collide(n_req, from_xy, to_xy, extend_length, path_resolution, obs_set, obstacles, obs_stride, n_obs,
                                play_area, new_xy, dist, n_points, free_flag, inside_flag, (cudaStream_t)stream);
}

int rrtk_steer_points_dev(int64_t n_req, const double *from_xy, const double *to_xy, const double *extend_length,
                          double extend_all, double path_resolution, int32_t pt_cap, double *points, int32_t *n_points,
                          void *stream) {
    if (n_req < 0 || pt_cap < 1) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (!(path_resolution > 0.0) || (!extend_length && !(extend_all >= 0.0)))
        return set_error(RRTK_ERR_INVALID, "path_resolution must be > 0, extend_length >= 0");
    if (n_req == 0) return RRTK_OK;
    if (!from_xy || !to_xy || !points || !n_points) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)from_xy | (uintptr_t)to_xy | (uintptr_t)points) & 15)
        return set_error(RRTK_ERR_INVALID, "from_xy / to_xy / points must be 16-byte aligned");
    return launch_steer_points(n_req, from_xy, to_xy, extend_length, extend_all, path_resolution, pt_cap, points, n_points,
                               (cudaStream_t)stream);
}

int rrtk_points_collide_dev(int32_t n_req, const double *points, const int32_t *n_points, int32_t pt_cap,
                            const int32_t *obs_set, const double *obstacles, int32_t obs_stride, const int32_t *n_obs,
                            uint8_t *free_flag, void *stream) {
    if (n_req < 0 || pt_cap < 1 || obs_stride < 0) return set_error(RRTK_ERR_INVALID, "bad sizes");
    if (n_req == 0) return RRTK_OK;
    if (!points || !n_points || !free_flag || (n_obs && obs_stride > 0 && !obstacles))
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)points | (uintptr_t)obstacles) & 15) return set_error(RRTK_ERR_INVALID, "points / obstacles must be 16-byte aligned");
    return launch_points_collide(n_req, points, n_points, pt_cap, obs_set, obstacles, obs_stride, n_obs, free_flag,
                                 (cudaStream_t)stream);
}

int rrtk_nearest_f64_dev(const double *xy, int64_t n, const double *samples, int32_t n_samples, int32_t *idx, double *d2,
                         void *stream) {
    if (n < 1 || n > 0x7fffffffll || n_samples < 0) return set_error(RRTK_ERR_INVALID, "need 1 <= n < 2^31 and n_samples >= 0");
    if (n_samples == 0) return RRTK_OK;
    if (!xy || !samples || !idx) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if (((uintptr_t)xy | (uintptr_t)samples) & 15) return set_error(RRTK_ERR_INVALID, "xy / samples must be 16-byte aligned");
    return launch_nearest_f64(xy, n, samples, n_samples, idx, d2, (cudaStream_t)stream);
}

int rrtk_near_f64_dev(const double *xy, int32_t n, double cx, double cy, double r2, int32_t *out_idx, double *scratch_d2,
                      int32_t cap, int32_t *out_n, void *stream) {
    if (n < 1 || cap < 0) return set_error(RRTK_ERR_INVALID, "need n >= 1 and cap >= 0");
    if (!xy || !out_n || (cap > 0 && (!out_idx || !scratch_d2))) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)xy & 15) return set_error(RRTK_ERR_INVALID, "xy must be 16-byte aligned");
    return launch_near_f64(xy, n, cx, cy, r2, out_idx, scratch_d2, cap, out_n, (cudaStream_t)stream);
}

int rrtk_sample_stream_dev(const rrtk_rrtstar_params *p, const double *start_goal,
                           const int64_t *sobol_offset, double *out, void *stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (p->sampler == RRTK_SAMPLER_STREAM) return set_error(RRTK_ERR_INVALID, "sampler = STREAM has nothing to generate");
    if (p->n_queries == 0 || p->max_iter == 0) return RRTK_OK;
    if (!start_goal || !out) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_sample_stream(*p, start_goal, sobol_offset, out, (cudaStream_t)stream);
}

int rrtk_crmath_probe_dev(int kind, int64_t n, const double *a, const double *b, double *out, void *stream) {
    if (kind < 0 || kind > 8 || n < 0) return set_error(RRTK_ERR_INVALID, "bad kind/n");
    if (n == 0) return RRTK_OK;
    if (!a || !out || (!b && (kind == 0 || kind == 1 || kind == 4 || kind == 5))) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_crmath_probe(kind, n, a, b, out, (cudaStream_t)stream);
}

int rrtk_nearest_f32_dev(const float *xy, int64_t n, const float *samples, int32_t n_samples,
                         uint64_t *scratch, int32_t *idx, float *d2, void *stream) {
    if (n < 1 || n > 0xfffffffell || n_samples < 1) return set_error(RRTK_ERR_INVALID, "need 1 <= n < 2^32 and n_samples >= 1");
    if (!xy || !samples || !scratch || !idx || !d2) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)xy & 15) return set_error(RRTK_ERR_INVALID, "xy must be 16-byte aligned");
    return launch_nearest(xy, n, samples, n_samples, (unsigned long long *)scratch, idx, d2, (cudaStream_t)stream);
}

int rrtk_near_f32_dev(const float *xy, int64_t n, float cx, float cy, float r2, int32_t *out_idx, int32_t cap,
                      int32_t *out_n, void *stream) {
    if (n < 1 || n > 0x7fffffffll || cap < 0) return set_error(RRTK_ERR_INVALID, "need 1 <= n < 2^31 and cap >= 0");
    if (!xy || !out_n || (cap > 0 && !out_idx)) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    if ((uintptr_t)xy & 15) return set_error(RRTK_ERR_INVALID, "xy must be 16-byte aligned");
    return launch_near(xy, n, cx, cy, r2, out_idx, cap, out_n, (cudaStream_t)stream);
}

int rrtk_arm_grid_dev(int32_t M, const double *theta, int32_t row0, int32_t n_rows, int32_t n_links,
                      const double *link_lengths, const double *obstacles, int32_t n_sets, int32_t n_obs,
                      uint8_t *grid, void *stream) {
    if (M < 1 || row0 < 0 || n_rows < 0 || row0 + n_rows > M) return set_error(RRTK_ERR_INVALID, "bad M / row range");
    if (n_links < 1 || n_links > 16) return set_error(RRTK_ERR_INVALID, "1 <= n_links <= 16");
    if (n_sets < 0 || n_obs < 0) return set_error(RRTK_ERR_INVALID, "negative set / obstacle count");
    if (n_rows == 0 || n_sets == 0) return RRTK_OK;
    if (!theta || !link_lengths || !grid || (n_obs > 0 && !obstacles)) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_arm_grid(M, theta, row0, n_rows, n_links, link_lengths, obstacles, n_sets, n_obs, grid,
                           (cudaStream_t)stream, 0);
}

int rrtk_arm_grid_cells_dev(int32_t M, const double *theta, int32_t row0, int32_t n_rows, int32_t n_links,
                            const double *link_lengths, const double *obstacles, int32_t n_sets, int32_t n_obs,
                            uint8_t *grid, void *stream) {
    if (M < 1 || row0 < 0 || n_rows < 0 || row0 + n_rows > M) return set_error(RRTK_ERR_INVALID, "bad M / row range");
    if (n_links < 1 || n_links > 16) return set_error(RRTK_ERR_INVALID, "1 <= n_links <= 16");
    if (n_sets < 0 || n_obs < 0) return set_error(RRTK_ERR_INVALID, "negative set / obstacle count");
    if (n_rows == 0 || n_sets == 0) return RRTK_OK;
    if (!theta || !link_lengths || !grid || (n_obs > 0 && !obstacles)) return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_arm_grid(M, theta, row0, n_rows, n_links, link_lengths, obstacles, n_sets, n_obs, grid,
                           (cudaStream_t)stream, 1);
}

int rrtk_astar_torus_dev(int32_t M, int32_t n_queries, const int32_t *start_goal, uint8_t *grids, int32_t *routes,
                         int32_t route_cap, int32_t *route_len, int32_t *expanded, int32_t *heur, int32_t *parents,
                         uint64_t *heaps, void *stream) {
    if (M < 1 || M > 46340 || n_queries < 0 || route_cap < 1) return set_error(RRTK_ERR_INVALID, "need 1 <= M <= 46340, route_cap >= 1");
    if (n_queries == 0) return RRTK_OK;
    if (!start_goal || !grids || !routes || !route_len || !expanded || !heur || !parents || !heaps)
        return set_error(RRTK_ERR_INVALID, "NULL pointer argument");
    return launch_astar_torus(M, n_queries, start_goal, grids, heur, parents, (unsigned long long *)heaps, routes, route_cap,
                              route_len, expanded, (cudaStream_t)stream);
}

int rrtk_fma_peak_dev(int fp64, int32_t iters, int32_t blocks, void *out, void *stream) {
    if (iters < 1 || blocks < 1 || !out) return set_error(RRTK_ERR_INVALID, "bad iters/blocks/out");
    return launch_fma_peak(fp64, iters, blocks, out, (cudaStream_t)stream);
}

// ---- host-buffer wrapper: the end-to-end path (H2D, kernel, D2H inside the call) ----
#define RRTK_TRY_CUDA(expr, what)                                      \
    do {                                                               \
        cudaError_t e__ = (expr);                                      \
        if (e__ != cudaSuccess) { rc = set_cuda_error(e__, what); goto cleanup; } \
    } while (0)

int rrtk_rrtstar_run_host(const rrtk_rrtstar_params *p, const double *start_goal,
                          const double *obstacles, const int32_t *n_obs, const double *near_r2,
                          const double *sample_stream, const int64_t *sobol_offset, double *xy,
                          double *cost, int32_t *parent, int32_t *n_nodes, int32_t *iters_done,
                          int32_t *goal_index, int32_t *status, int32_t *trace) {
    int rc = check_params(p);
    if (rc) return rc;
    if (p->n_queries == 0) return RRTK_OK;
    const size_t Q = (size_t)p->n_queries, cap = (size_t)p->node_cap, it = (size_t)p->max_iter;
    const size_t b_sg = Q * 4 * 8, b_obs = Q * (size_t)p->obs_stride * 4 * 8, b_no = Q * 4,
                 b_r2 = (cap + 2) * 8, b_st = sample_stream ? Q * it * 16 : 0,
                 b_so = sobol_offset ? Q * 8 : 0, b_xy = Q * cap * 16, b_c = Q * cap * 8,
                 b_p = Q * cap * 4, b_q = Q * 4, b_tr = trace ? Q * it * 32 : 0, b_ws = (Q * RRTK_RRTSTAR_WS_INTS(cap, p->grid_nx, p->grid_ny) + RRTK_WS_TAIL_INTS) * 4;
    char *d = nullptr;
    size_t off[16], total = 0;
    const size_t sizes[15] = {b_sg, b_obs, b_no, b_r2, b_st, b_so, b_xy, b_c, b_p, b_q, b_q, b_q, b_q, b_tr, b_ws};
    for (int i = 0; i < 15; i++) { off[i] = total; total += (sizes[i] + 255) & ~(size_t)255; }
    cudaStream_t s = 0;
    RRTK_TRY_CUDA(cudaMalloc(&d, total ? total : 256), "cudaMalloc");
    RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[0], start_goal, b_sg, cudaMemcpyHostToDevice, s), "H2D start_goal");
    if (b_obs) RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[1], obstacles, b_obs, cudaMemcpyHostToDevice, s), "H2D obstacles");
    RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[2], n_obs, b_no, cudaMemcpyHostToDevice, s), "H2D n_obs");
    if (near_r2) RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[3], near_r2, b_r2, cudaMemcpyHostToDevice, s), "H2D near_r2");
    if (b_st) RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[4], sample_stream, b_st, cudaMemcpyHostToDevice, s), "H2D stream");
    if (b_so) RRTK_TRY_CUDA(cudaMemcpyAsync(d + off[5], sobol_offset, b_so, cudaMemcpyHostToDevice, s), "H2D sobol_offset");
    rc = rrtk_rrtstar_run_dev(p, (double *)(d + off[0]), (double *)(d + off[1]), (int32_t *)(d + off[2]),
                              near_r2 ? (double *)(d + off[3]) : nullptr,
                              b_st ? (double *)(d + off[4]) : nullptr, b_so ? (int64_t *)(d + off[5]) : nullptr,
                              (double *)(d + off[6]), (double *)(d + off[7]), (int32_t *)(d + off[8]),
                              (int32_t *)(d + off[9]), (int32_t *)(d + off[10]), (int32_t *)(d + off[11]),
                              (int32_t *)(d + off[12]), b_tr ? (int32_t *)(d + off[13]) : nullptr,
                              (int32_t *)(d + off[14]), s);
    if (rc) goto cleanup;
    RRTK_TRY_CUDA(cudaMemcpyAsync(xy, d + off[6], b_xy, cudaMemcpyDeviceToHost, s), "D2H xy");
    RRTK_TRY_CUDA(cudaMemcpyAsync(cost, d + off[7], b_c, cudaMemcpyDeviceToHost, s), "D2H cost");
    RRTK_TRY_CUDA(cudaMemcpyAsync(parent, d + off[8], b_p, cudaMemcpyDeviceToHost, s), "D2H parent");
    RRTK_TRY_CUDA(cudaMemcpyAsync(n_nodes, d + off[9], b_q, cudaMemcpyDeviceToHost, s), "D2H n_nodes");
    RRTK_TRY_CUDA(cudaMemcpyAsync(iters_done, d + off[10], b_q, cudaMemcpyDeviceToHost, s), "D2H iters_done");
    RRTK_TRY_CUDA(cudaMemcpyAsync(goal_index, d + off[11], b_q, cudaMemcpyDeviceToHost, s), "D2H goal_index");
    RRTK_TRY_CUDA(cudaMemcpyAsync(status, d + off[12], b_q, cudaMemcpyDeviceToHost, s), "D2H status");
    if (b_tr) RRTK_TRY_CUDA(cudaMemcpyAsync(trace, d + off[13], b_tr, cudaMemcpyDeviceToHost, s), "D2H trace");
    RRTK_TRY_CUDA(cudaStreamSynchronize(s), "cudaStreamSynchronize");
cleanup:
    if (d) cudaFree(d);
    return rc;
}

}  // extern "C"
