// rrtk_planner.cuh -- device helpers shared by the planner kernels (RRT*, Informed RRT*).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrtk {

constexpr int CULL_CAP = 64;      // culled obstacle list per warp (overflow -> full list)
constexpr unsigned FULL = 0xffffffffu;

struct ObsList {  // SoA view of the obstacles an edge must be tested against
    const double *ox, *oy, *r2;
    int stride;  // element stride (1 for the shared-memory list, 4 for the global AoS rows)
    int m;
};

// warp argmin of (value, index): smaller value wins, ties -> smaller index (list.index(min(..)))
static __device__ __forceinline__ void warp_argmin(double &v, int &i) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        double ov = __shfl_xor_sync(FULL, v, off);
        int oi = __shfl_xor_sync(FULL, i, off);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// Conservative exact cull: keep obstacle o iff |o - c| <= (reach + R_o) * (1 + 1e-9) + 1e-9.
// Returns the list to test edges against (shared-memory survivors, or all obstacles on overflow).
static __device__ __noinline__ ObsList cull_obstacles(const double4 *obs, int n_obs, double cx, double cy,
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
    ObsList L;
    if (!overflow) {
        L.ox = sx; L.oy = sy; L.r2 = sr2; L.stride = 1; L.m = count;
    } else {
        const double *g = reinterpret_cast<const double *>(obs);
        L.ox = g; L.oy = g + 1; L.r2 = g + 3; L.stride = 4; L.m = n_obs;
    }
    return L;
}

}  // namespace rrtk
