// rrtk_astar.cu -- astar_torus (arm02:113-184) for Q (grid, start, goal) queries at once: the consumer of the arm
// C-space occupancy grids, so grid -> route stays on the GPU.
//
// Despite its name the reference's search is GREEDY BEST-FIRST: the frontier is ordered by the heuristic alone
// (explored_heuristic_map[neighbor] = heuristic_map[neighbor], arm02:163), ties go to the smallest row-major index
// (np.argmin), a cell enters the frontier once (grid value 3), the goal is re-armed every iteration (:140-141).
// The route depends on that exact expansion order, so each query runs the same sequential search -- a binary heap
// keyed by (heuristic << 32 | flat index) replaces the reference's O(M^2) argmin per expansion -- and the
// parallelism is across queries (obstacle sets x start/goal pairs), one thread each.  calc_heuristic_map
// (:221-233) updates its array in place while reading it; which neighbours are read back already updated is fixed
// by the row-major order (row 0 by every later row, column 0 by the rest of its row), giving the closed form of
// heuristic_cell() -- integer arithmetic, exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rrtk.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ long long at_abs(long long v) { return v < 0 ? -v : v; }
__device__ __forceinline__ long long at_orig(int gi, int gj, int i, int j) { return at_abs((long long)j - gj) + at_abs((long long)i - gi); }
// one in-place update step of arm02:226-231 given the (possibly updated) row-0 and column-0 values it reads
__device__ __forceinline__ long long at_step(int M, int gi, int gj, int i, int j, long long row0, long long col0) {
    long long v = at_orig(gi, gj, i, j), t;
    t = i + 1 + at_orig(gi, gj, M - 1, j); v = t < v ? t : v;
    t = M - i + row0; v = t < v ? t : v;
    t = j + 1 + at_orig(gi, gj, i, M - 1); v = t < v ? t : v;
    t = M - j + col0; v = t < v ? t : v;
    return v;
}
__device__ long long heuristic_cell(int M, int gi, int gj, int i, int j) {
    const long long n00 = at_step(M, gi, gj, 0, 0, at_orig(gi, gj, 0, 0), at_orig(gi, gj, 0, 0));
    const long long row0 = i > 0 ? (j == 0 ? n00 : at_step(M, gi, gj, 0, j, at_orig(gi, gj, 0, j), n00)) : at_orig(gi, gj, 0, j);
    const long long col0 = j > 0 ? (i == 0 ? n00 : at_step(M, gi, gj, i, 0, n00, at_orig(gi, gj, i, 0))) : at_orig(gi, gj, i, 0);
    return at_step(M, gi, gj, i, j, row0, col0);
}

__global__ void astar_heuristic_kernel(int M, int n_queries, const int32_t *__restrict__ start_goal, int32_t *heur) {
    const long long cells = (long long)M * M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells * n_queries) return;
    const int q = (int)(t / cells);
    const long long c = t - (long long)q * cells;
    heur[t] = (int32_t)heuristic_cell(M, start_goal[4 * q + 2], start_goal[4 * q + 3], (int)(c / M), (int)(c % M));
}

// Frontier = a 32-ary min-heap of (heuristic << 32 | flat index) keys in global memory, one WARP per query: a node's 32
// children are one coalesced 256-byte load and their minimum one pair of warp reductions, so a pop walks log32(n) <= 4
// levels (a binary heap walked by one thread took ~18 dependent L2 round trips per pop).  Keys are unique (the index is
// part of the key), so any correct priority queue pops the reference's np.argmin order.
__device__ __forceinline__ void heap_push(unsigned long long *heap, int &n, unsigned long long key, int lane) {
    int k = n++;
    while (k > 0) {
        const int p = (k - 1) >> 5;
        const unsigned long long pk = heap[p];
        if (pk <= key) break;
        if (lane == 0) heap[k] = pk;
        k = p;
    }
    if (lane == 0) heap[k] = key;
    __syncwarp();
}
__device__ __forceinline__ void heap_pop(unsigned long long *heap, int &n, int lane) {
    const unsigned long long key = heap[--n];
    int k = 0;
    for (;;) {
        const int c = 32 * k + 1 + lane;
        const unsigned long long ck = c < n ? heap[c] : ~0ull;
        const unsigned hi = (unsigned)(ck >> 32), lo = (unsigned)ck;
        const unsigned mhi = __reduce_min_sync(FULL, hi);
        const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
        const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
        if (mk >= key) break;                      // also when there are no children (all ~0)
        const int src = __ffs(__ballot_sync(FULL, ck == mk)) - 1;
        if (lane == 0) heap[k] = mk;
        k = 32 * k + 1 + src;
    }
    if (n > 0 && lane == 0) heap[k] = key;
    __syncwarp();
}

// one warp per query
constexpr int ASTAR_WARPS = 4;
__global__ void __launch_bounds__(ASTAR_WARPS * 32)
astar_torus_kernel(int M, int n_queries, const int32_t *__restrict__ start_goal, const int32_t *__restrict__ heur,
                   uint8_t *grids, int32_t *parents, unsigned long long *heaps, int32_t *routes,
                   int route_cap, int32_t *route_len, int32_t *expanded) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * ASTAR_WARPS + (threadIdx.x >> 5);
    if (q >= n_queries) return;
    const size_t cells = (size_t)M * M;
    uint8_t *grid = grids + q * cells;
    int32_t *parent = parents + q * cells;
    const int32_t *h = heur + q * cells;
    unsigned long long *heap = heaps + q * (cells + 8);
    int32_t *route = routes + (size_t)q * route_cap * 2;
    const int s = start_goal[4 * q] * M + start_goal[4 * q + 1], g = start_goal[4 * q + 2] * M + start_goal[4 * q + 3];
    int nheap = 0, n_exp = 0;
    bool found = false;
    if (lane == 0) { grid[s] = 4; grid[g] = 5; }
    __syncwarp();
    heap_push(heap, nheap, ((unsigned long long)(unsigned)h[s] << 32) | (unsigned)s, lane);
    for (;;) {
        if (lane == 0) { grid[s] = 4; grid[g] = 5; }     // arm02:140-141
        __syncwarp();
        if (nheap == 0) break;                           // min is inf: no route
        const int cur = (int)(heap[0] & 0xffffffffull);  // np.argmin: smallest (heuristic, row-major index)
        if (cur == g) { found = true; break; }
        heap_pop(heap, nheap, lane);
        n_exp++;
        const int i = cur / M, j = cur - i * M;
        // find_neighbors (:187-209): up, down, left, right on the torus; lanes 0..3 read one neighbour each
        int nb = 0;
        if (lane == 0) nb = (i - 1 >= 0 ? i - 1 : M - 1) * M + j;
        else if (lane == 1) nb = (i + 1 < M ? i + 1 : 0) * M + j;
        else if (lane == 2) nb = i * M + (j - 1 >= 0 ? j - 1 : M - 1);
        else if (lane == 3) nb = i * M + (j + 1 < M ? j + 1 : 0);
        if (lane == 0) grid[cur] = 2;
        __syncwarp();
        uint8_t v = 1;
        int hv = 0;
        if (lane < 4) { v = grid[nb]; hv = h[nb]; }
        // on a 1- or 2-wide torus two of the four neighbours coincide: the first occurrence enters the frontier, the later one
        // then sees value 3 -- resolve duplicates in the reference's order
        bool take = lane < 4 && (v == 0 || v == 5);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int onb = __shfl_sync(FULL, nb, k);
            const bool otake = __shfl_sync(FULL, (int)take, k) != 0;
            if (lane > k && lane < 4 && otake && onb == nb) take = false;
        }
        if (take) { parent[nb] = cur; grid[nb] = 3; }
        const unsigned tm = __ballot_sync(FULL, take);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (tm & (1u << k)) {
                const int knb = __shfl_sync(FULL, nb, k), khv = __shfl_sync(FULL, hv, k);
                heap_push(heap, nheap, ((unsigned long long)(unsigned)khv << 32) | (unsigned)knb, lane);
            }
        }
    }
    int len = 0;
    if (lane == 0) {
        if (found) {
            for (int k = g; k >= 0; k = parent[k]) len++;
            if (len <= route_cap) {
                int w = len - 1;
                for (int k = g; k >= 0; k = parent[k], w--) {
                    route[2 * w] = k / M; route[2 * w + 1] = k % M;
                    if (w >= 1) grid[k] = 6;                 // arm02:172-173
                }
            } else {
                len = -len;                                  // does not fit: report the length negated
            }
        }
        route_len[q] = len;
        expanded[q] = n_exp;
    }
}

int launch_astar_torus(int M, int n_queries, const int32_t *start_goal, uint8_t *grids, int32_t *heur, int32_t *parents,
                       unsigned long long *heaps, int32_t *routes, int route_cap, int32_t *route_len, int32_t *expanded,
                       cudaStream_t s) {
    const size_t cells = (size_t)M * M;
    cudaError_t e = cudaMemsetAsync(parents, 0xff, sizeof(int32_t) * cells * n_queries, s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(parents)");
    const long long total = (long long)cells * n_queries;
    astar_heuristic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(M, n_queries, start_goal, heur);
    astar_torus_kernel<<<(unsigned)((n_queries + ASTAR_WARPS - 1) / ASTAR_WARPS), ASTAR_WARPS * 32, 0, s>>>(
        M, n_queries, start_goal, heur, grids, parents, heaps, routes, route_cap, route_len, expanded);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "astar_torus kernels launch");
    return RRTK_OK;
}

}  // namespace rrtk
