// rrtk_arm.cu -- C-space occupancy grid of a planar N-link arm among circles
// (get_occupancy_grid arm02:79-110, NLinkArm.update_points :257-262, detect_collision :46-76).
//
// One thread per joint-space cell (i, j), one CTA per 128 columns of a row at a time; the forward kinematics of the cell
// (prefix-sum angles with the reference's 2-angle rule, correctly rounded cos/sin) is computed ONCE and tested against
// every obstacle set (S sets x O circles staged in shared memory), so trigonometry is amortised over S.
// Per (link, circle): a division-free FP64 filter on squared distances decides all clear cases; only
// circles within a 1e-9 relative band of tangency run the reference's exact operation sequence
// (norm / projection / clamp with numpy's fma dot), so the grid equals the reference's bit for bit.
// Link 1 is tested once per row, links 2..n as one straight stretch first (see arm_grid_kernel).
// Output is uint8 [S][rows][M], written 32 contiguous bytes per warp and set.
#include "crmath.h"
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int ARM_MAX_LINKS = 16;
constexpr int ARM_THREADS = 128;

struct ArmParams {
    int M, row0, n_rows, n_links, S, O;
    int stretch;   // links 2..n are tested as one straight stretch first (all of them longer than 0)
    double link[ARM_MAX_LINKS];
};

// numpy's 2-vector dot on the reference platform: fma(a1, b1, a0 * b0)
__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

// detect_collision (arm02:46-76), the reference's exact operation sequence
__device__ __noinline__ bool detect_collision_exact(double ax, double ay, double bx, double by, double cx,
                                                    double cy, double r) {
    double l0 = bx - ax, l1 = by - ay;
    double mag = sqrt(dot2(l0, l1, l0, l1));
    double v0 = cx - ax, v1 = cy - ay;
    double proj = dot2(v0, v1, l0 / mag, l1 / mag);
    double p0, p1;
    if (proj <= 0) { p0 = ax; p1 = ay; }
    else if (proj >= mag) { p0 = bx; p1 = by; }
    else { p0 = ax + l0 * proj / mag; p1 = ay + l1 * proj / mag; }
    double w0 = p0 - cx, w1 = p1 - cy;
    double dist = sqrt(dot2(w0, w1, w0, w1));
    return !(dist > r);
}

// One (link, circle) test: the division-free filter on squared distances scaled by |l|^2, the reference's sequence inside
// its band of 1e-9 (relative).  The filter takes fma's (it only has to be right to the band).
//   1 touching, 0 clear -- exactly what detect_collision returns
__device__ __forceinline__ int link_circle_filter(double ax, double ay, double bx, double by, double L2, double cx, double cy,
                                                  double r2) {
    const double l0 = bx - ax, l1 = by - ay;
    const double v0 = cx - ax, v1 = cy - ay;
    const double dotp = fma(v1, l1, v0 * l0);
    const double vvL2 = fma(v1, v1, v0 * v0) * L2;
    double sdist;  // = dist^2 * L2
    if (dotp <= 0.0) sdist = vvL2;
    else if (dotp >= L2) {
        const double w0 = cx - bx, w1 = cy - by;
        sdist = fma(w1, w1, w0 * w0) * L2;
    } else sdist = fma(-dotp, dotp, vvL2);
    const double thr = r2 * L2;
    const double t = sdist - thr, band = 1e-9 * (vvL2 + thr);
    if (t > band) return 0;                       // certainly clear
    if (t < -band && L2 > 0.0) return 1;          // certainly touching
    return -1;
}
__device__ __noinline__ bool link_circle_hit(double ax, double ay, double bx, double by, double cx, double cy, double r, double r2) {
    const double l0 = bx - ax, l1 = by - ay;
    const int v = link_circle_filter(ax, ay, bx, by, l0 * l0 + l1 * l1, cx, cy, r2);
    return v >= 0 ? v != 0 : detect_collision_exact(ax, ay, bx, by, cx, cy, r);
}

// MAXL = compile-time bound on the number of links (the per-cell joint positions live in registers).
// The reference's update_points gives joint 1 the angle theta1 and EVERY later joint theta1 + theta2 (arm02:257-262 sums
// joint_angles[:i] of a two-entry list), so
//   * link 1 depends on the row only: its S x O tests run once per row tile (thread s takes obstacle set s) and land in
//     shared memory as one flag per set;
//   * links 2..n lie end to end on one straight line from joint 1: a circle clear of the whole stretch [p1, pn] beyond the
//     band is clear of every link, one that reaches into it beyond the band touches the link its nearest point lies on
//     (the joints computed by the reference's running sums are within ~1e-15 of that line, the band is 1e-9), and only a
//     circle inside the band of the stretch runs the per-link tests.  (Not with a zero-length link among them: the
//     reference's 0 / 0 makes that link "touch" everything -- the host clears `stretch` then.)
// One CTA works on 128 consecutive columns of one row at a time.
template <int MAXL>
__global__ void __launch_bounds__(ARM_THREADS)
arm_grid_kernel(ArmParams p, const double *__restrict__ theta, const double *__restrict__ obstacles,
                uint8_t *__restrict__ grid) {
    extern __shared__ double s_obs[];  // [S * O][4]: x, y, r, r * r; then S row flags
    const int n_circ = p.S * p.O;
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_obs + 4 * (size_t)n_circ);
    for (int t = threadIdx.x; t < n_circ; t += blockDim.x) {
        double x = obstacles[3 * t], y = obstacles[3 * t + 1], r = obstacles[3 * t + 2];
        s_obs[4 * t] = x; s_obs[4 * t + 1] = y; s_obs[4 * t + 2] = r; s_obs[4 * t + 3] = r < 0.0 ? -1.0 : r * r;   // (a negative radius touches nothing: dist > r always, arm02:72)
    }
    const int tiles_per_row = (p.M + ARM_THREADS - 1) / ARM_THREADS;
    const long long tiles = (long long)p.n_rows * tiles_per_row;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int ir = (int)(tile / tiles_per_row), j = (int)(tile % tiles_per_row) * ARM_THREADS + threadIdx.x;
        const int i = p.row0 + ir;
        // forward kinematics (arm02:257-262): joint k uses theta1 (k = 1) or theta1 + theta2 (k >= 2)
        const double a1 = theta[i];
        const double c1 = crm_cos(a1), s1 = crm_sin(a1);
        double px[MAXL + 1], py[MAXL + 1];
        px[0] = 0.0; py[0] = 0.0;
        px[1] = px[0] + p.link[0] * c1;
        py[1] = py[0] + p.link[0] * s1;
        __syncthreads();   // the circles are staged / the previous tile's flags have been read
        for (int s = threadIdx.x; s < p.S; s += blockDim.x) {   // link 1 against obstacle set s
            const double *ob = s_obs + (size_t)s * p.O * 4;
            bool hit = false;
            for (int o = 0; o < p.O && !hit; o++)
                hit = link_circle_hit(px[0], py[0], px[1], py[1], ob[4 * o], ob[4 * o + 1], ob[4 * o + 2], ob[4 * o + 3]);
            s_row[s] = hit ? 1 : 0;
        }
        __syncthreads();
        if (j >= p.M) continue;
        const double a2 = a1 + theta[j];
        const double c2 = crm_cos(a2), s2 = crm_sin(a2);
#pragma unroll
        for (int k = 2; k <= MAXL; k++) {
            if (k <= p.n_links) {
                px[k] = px[k - 1] + p.link[k - 1] * c2;
                py[k] = py[k - 1] + p.link[k - 1] * s2;
            }
        }
        // the stretch of links 2..n
        double ex = px[1], ey = py[1];
#pragma unroll
        for (int k = 2; k <= MAXL; k++)
            if (k == p.n_links) { ex = px[k]; ey = py[k]; }
        const double u0 = ex - px[1], u1 = ey - py[1], U2 = fma(u1, u1, u0 * u0);
        for (int s = 0; s < p.S; s++) {
            const double *ob = s_obs + (size_t)s * p.O * 4;
            bool hit = s_row[s] != 0;
            if (p.n_links >= 2) {
                for (int o = 0; o < p.O && !hit; o++) {
                    const double cx = ob[4 * o], cy = ob[4 * o + 1], r2 = ob[4 * o + 3];
                    const int v = p.stretch ? link_circle_filter(px[1], py[1], ex, ey, U2, cx, cy, r2) : -1;
                    if (v >= 0) { hit = v != 0; continue; }
#pragma unroll
                    for (int k = 1; k < MAXL; k++)
                        if (k < p.n_links && !hit) hit = link_circle_hit(px[k], py[k], px[k + 1], py[k + 1], cx, cy, ob[4 * o + 2], r2);
                }
            }
            grid[((size_t)s * p.n_rows + ir) * p.M + j] = hit ? 1 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row rasteriser.  With joint 1 fixed (one row of the grid) links 2..n are one stretch of length Ls = sum(link[1:]) that
// turns about joint 1 with the column angle, and a circle (centre at distance d > r from joint 1, direction psi) is
// touched exactly for the column angles within +-alpha of psi:
//     alpha = asin(r / d)                                  if Ls reaches the tangent point (Ls >= sqrt(d^2 - r^2)),
//     alpha = acos((d^2 + Ls^2 - r^2) / (2 d Ls))          if only the far end reaches in (d - r < Ls < that), none otherwise.
// The columns are theta_list[j] = 2 (j - M // 2) pi / M, so each circle is a run of columns (mod M).  Per (row, set,
// circle) the thread of that set computes the run twice: the CERTAIN one (alpha shrunk by 1e-7 rad and the ratios by 1e-9,
// orders of magnitude above every rounding involved) and the POSSIBLE one (alpha grown by the same); the grid row of a set
// is then written 32 cells at a time as the OR of the certain runs, and the few cells that are possible but not certain
// (at most a couple per circle and row) are evaluated like the reference does -- arm_cell_hit: the joints by the
// reference's running sums, per link the filter / exact sequence above.  The kernel writes M^2 S bytes and is bound by
// that store stream instead of the FP64 pipe.
// A theta array that is not the reference's list (checked on the device), or links that the stretch argument does not
// cover, make every cell "possible": the result is still the reference's, cell by cell.
constexpr int ARM_RASTER_THREADS = 256;

// cell (joint 1 at p1, column angle a2) against the O circles of one set, links 2..n only -- the reference's sequence
__device__ __noinline__ bool arm_cell_hit(const double *link, int n_links, int stretch, double p1x, double p1y, double a2,
                                          const double *ob, int O) {
    const double c2 = crm_cos(a2), s2 = crm_sin(a2);
    double px[ARM_MAX_LINKS + 1], py[ARM_MAX_LINKS + 1];
    px[1] = p1x; py[1] = p1y;
    for (int k = 2; k <= n_links; k++) {
        px[k] = px[k - 1] + link[k - 1] * c2;
        py[k] = py[k - 1] + link[k - 1] * s2;
    }
    const double ex = px[n_links], ey = py[n_links];
    const double u0 = ex - p1x, u1 = ey - p1y, U2 = fma(u1, u1, u0 * u0);
    for (int o = 0; o < O; o++) {
        const double cx = ob[4 * o], cy = ob[4 * o + 1], r2 = ob[4 * o + 3];
        const int v = stretch ? link_circle_filter(p1x, p1y, ex, ey, U2, cx, cy, r2) : -1;
        if (v > 0) return true;
        if (v == 0) continue;
        for (int k = 1; k < n_links; k++)
            if (link_circle_hit(px[k], py[k], px[k + 1], py[k + 1], cx, cy, ob[4 * o + 2], r2)) return true;
    }
    return false;
}

// the cells k of [j0, j0 + 32) with ((j0 + k - lo) mod M) <= w, as a bit mask (w < 0: none)
__device__ __forceinline__ unsigned span_bits(int lo, int w, int j0, int M) {
    if (w < 0) return 0u;
    if (w >= M - 1) return ~0u;
    int d0 = j0 - lo;
    if (d0 < 0) d0 += M;
    unsigned m = 0u;
    if (d0 <= w) { const int n = w - d0 + 1; m = n >= 32 ? ~0u : (1u << n) - 1u; }      // the run continues into the chunk
    const int kb = M - d0;                                                                 // the run starts at cell kb
    if (kb < 32) { const int n = w + 1; m |= (n >= 32 ? ~0u : (1u << n) - 1u) << kb; }
    return m;
}

#ifndef ARM_RASTER_MINB
#define ARM_RASTER_MINB 3
#endif
__global__ void __launch_bounds__(ARM_RASTER_THREADS, ARM_RASTER_MINB)
arm_grid_rows_kernel(ArmParams p, const double *__restrict__ theta, const double *__restrict__ obstacles,
                     uint8_t *__restrict__ grid, int vec_ok, int G) {
    // [S * O][4] doubles: x, y, r, r * r | [S * O] int4 runs: certain (lo, w), possible (lo, w) | [G][W] row bitmaps of a
    // group of G sets | S row flags | S "evaluate every cell" flags
    extern __shared__ double s_obs[];
    const int n_circ = p.S * p.O, M = p.M, W = (M + 31) >> 5;
    int4 *s_run = reinterpret_cast<int4 *>(s_obs + 4 * (size_t)n_circ);
    unsigned *s_bm = reinterpret_cast<unsigned *>(s_run + n_circ);
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_bm + (size_t)G * W);
    uint8_t *s_all = s_row + p.S;
    __shared__ double s_link[ARM_MAX_LINKS];
    __shared__ int s_any_all;   // some set of this row is evaluated cell by cell
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int t = threadIdx.x; t < n_circ; t += blockDim.x) {
        double x = obstacles[3 * t], y = obstacles[3 * t + 1], r = obstacles[3 * t + 2];
        s_obs[4 * t] = x; s_obs[4 * t + 1] = y; s_obs[4 * t + 2] = r; s_obs[4 * t + 3] = r < 0.0 ? -1.0 : r * r;   // (a negative radius touches nothing: dist > r always, arm02:72)
    }
    for (int t = threadIdx.x; t < G * W; t += blockDim.x) s_bm[t] = 0u;
    if (threadIdx.x < ARM_MAX_LINKS) s_link[threadIdx.x] = p.link[threadIdx.x];
    const int off = -((M + 1) / 2);                     // Python's -M // 2
    // is theta the reference's list?  (2 * i * pi / M, i = j - M // 2 ...: the host evaluates it in Python floats)
    bool bad = false;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const double want = 2.0 * (double)(j + off) * 3.141592653589793 / (double)M;
        if (!(fabs(theta[j] - want) <= 1e-12)) bad = true;
    }
    const bool raster = !__syncthreads_or(bad) && p.stretch && p.n_links >= 2;
    double Ls = 0.0;
    for (int k = 1; k < p.n_links; k++) Ls += p.link[k];
    const double m_per_rad = (double)M / 6.283185307179586;
    for (int ir = blockIdx.x; ir < p.n_rows; ir += gridDim.x) {
        const int i = p.row0 + ir;
        const double a1 = theta[i];
        const double c1 = crm_cos(a1), s1 = crm_sin(a1);
        const double p1x = 0.0 + p.link[0] * c1, p1y = 0.0 + p.link[0] * s1;   // points[1] (arm02:259-260)
        // ---- per (set, circle): link 1 (the same for the whole row), then the circle's runs ----
        for (int t = threadIdx.x; t < p.S; t += blockDim.x) { s_row[t] = 0; s_all[t] = 0; }
        if (threadIdx.x == 0) s_any_all = 0;
        __syncthreads();
        for (int u = threadIdx.x; u < n_circ; u += blockDim.x) {
            const int s = u / p.O;
            const double *ob = s_obs + (size_t)u * 4;
            if (link_circle_hit(0.0, 0.0, p1x, p1y, ob[0], ob[1], ob[2], ob[3])) s_row[s] = 1;   // (every writer writes 1)
            if (p.n_links < 2) continue;
            int4 run = make_int4(0, -1, 0, M);   // certain: none, possible: every column
            const double vx = ob[0] - p1x, vy = ob[1] - p1y, r = ob[2];
            const double d2 = vx * vx + vy * vy, d = sqrt(d2);
            if (raster && d > r * (1.0 + 1e-9) + 1e-12 && d < 1e4 && Ls > 1e-6 && Ls < 1e4 && r >= 0.0) {
                const double em = 1e-9, ea = 1e-7;
                const double q = r / d, Lt = sqrt(d2 - r * r);
                const double at_out = asin(fmin(1.0, q + em)) + ea, at_in = q - em > 0.0 ? asin(q - em) - ea : -1.0;
                double a_out, a_in;
                if (Ls >= Lt * (1.0 + 1e-9)) { a_out = at_out; a_in = at_in; }
                else if (Ls <= (d - r) * (1.0 - 1e-9) - 1e-12) { a_out = -1.0; a_in = -1.0; }   // out of reach
                else {
                    const double x = (d2 + Ls * Ls - r * r) / (2.0 * d * Ls);
                    const double ae_out = x - em > 1.0 ? -1.0 : acos(fmax(-1.0, x - em)) + ea;
                    const double ae_in = x + em >= 1.0 ? -1.0 : acos(fmax(-1.0, x + em)) - ea;
                    a_out = Ls < Lt * (1.0 - 1e-9) ? ae_out : fmax(at_out, ae_out);
                    a_in = fmin(at_in, ae_in);
                }
                const double psi = atan2(vy, vx);
                auto cells = [&](double a, int &lo, int &w) {   // columns j with |a1 + theta[j] - psi| <= a (mod 2 pi)
                    if (a < 0.0) { lo = 0; w = -1; return; }
                    if (a >= 3.14159) { lo = 0; w = M; return; }
                    const double tl = (psi - a - a1) * m_per_rad - (double)off, th = (psi + a - a1) * m_per_rad - (double)off;
                    const double fl = ceil(tl), fh = floor(th);
                    if (fh < fl) { lo = 0; w = -1; return; }
                    w = (int)(fh - fl);
                    double lm = fl - floor(fl / (double)M) * (double)M;   // fl mod M (|fl| < 2^31: exact)
                    if (lm < 0.0) lm += (double)M;
                    lo = (int)lm;
                    if (lo >= M) lo -= M;
                };
                cells(a_in, run.x, run.y);
                cells(a_out, run.z, run.w);
            }
            if (run.w >= M - 1 && run.y < M - 1) { s_all[s] = 1; s_any_all = 1; }   // undecided everywhere: the set's row cell by cell
            s_run[u] = run;
        }
        __syncthreads();
        uint8_t *row_out = grid + (size_t)ir * M;
        const size_t set_stride = (size_t)p.n_rows * M;
        for (int g0 = 0; g0 < p.S; g0 += G) {
            const int gn = p.S - g0 < G ? p.S - g0 : G;
            // ---- paint: one warp per (set, circle); the certain run word by word, the few undecided cells one by one ----
            if (p.n_links >= 2 && p.O > 0) {
                // (set, circle) pairs warp, warp + n_warps, ...: the pair index is split without a division
                const int dsl_p = n_warps / p.O, do_p = n_warps - dsl_p * p.O;
                for (int sl = warp / p.O, o = warp - (warp / p.O) * p.O; sl < gn;) {
                    const int s = g0 + sl, o_cur = o;
                    const bool skip_pair = s_row[s] || s_all[s];
                    const int4 run = s_run[s * p.O + o_cur];
                    sl += dsl_p; o += do_p;
                    if (o >= p.O) { o -= p.O; sl++; }
                    if (skip_pair) continue;
                    const int sl_cur = s - g0;
                    unsigned *bm = s_bm + (size_t)sl_cur * W;
                    if (run.y >= 0) {
                        // words the run can reach into, from the one holding lo (one more when the last word of the row is
                        // partial: fewer cells there before the run wraps to column 0)
                        const int nw = ((run.y + 32) >> 5) + 2;
                        for (int t = lane; t < nw && t < W; t += 32) {
                            int wi = (run.x >> 5) + t;
                            if (wi >= W) wi -= W;
                            const unsigned m = span_bits(run.x, run.y, wi << 5, M);
                            if (m) atomicOr(&bm[wi], m);
                        }
                    }
                    if (run.z != run.x || run.w != run.y) {
                        // possible but not certain: gl cells before the certain run, the rest after it (all of the possible
                        // run when nothing is certain)
                        int gl = run.w + 1, skip = 0;
                        if (run.y >= 0) { gl = run.x - run.z; if (gl < 0) gl += M; skip = run.y + 1; }
                        const int n_und = run.w + 1 - skip;
                        for (int t = lane; t < n_und; t += 32) {
                            int j = run.z + (t < gl ? t : t + skip);
                            if (j >= M) j -= M;
                            if (j >= M) j -= M;
                            if (arm_cell_hit(s_link, p.n_links, p.stretch, p1x, p1y, a1 + theta[j], s_obs + (size_t)s * p.O * 4, p.O))
                                atomicOr(&bm[j >> 5], 1u << (j & 31));
                        }
                    }
                }
                for (int sl = 0; s_any_all && sl < gn; sl++) {   // sets to evaluate cell by cell (inputs the runs do not cover)
                    const int s = g0 + sl;
                    if (s_row[s] || !s_all[s]) continue;
                    for (int j = threadIdx.x; j < M; j += blockDim.x)
                        if (arm_cell_hit(s_link, p.n_links, p.stretch, p1x, p1y, a1 + theta[j], s_obs + (size_t)s * p.O * 4, p.O))
                            atomicOr(&s_bm[(size_t)sl * W + (j >> 5)], 1u << (j & 31));
                }
            }
            __syncthreads();
            // ---- emit: 32 cells per thread and step, bit -> byte; the bitmap is left cleared for the next group / row ----
            const int dsl_e = blockDim.x / W, dwi_e = blockDim.x - dsl_e * W;
            for (int sl = threadIdx.x / W, wi = threadIdx.x - (threadIdx.x / W) * W; sl < gn;) {
                const int s = g0 + sl, j0 = wi << 5, u = sl * W + wi;
                sl += dsl_e; wi += dwi_e;
                if (wi >= W) { wi -= W; sl++; }
                const int ncell = M - j0 < 32 ? M - j0 : 32;
                unsigned bits = s_bm[u];
                s_bm[u] = 0u;
                if (s_row[s]) bits = ~0u;
                uint8_t *out = row_out + (size_t)s * set_stride + j0;
                if (vec_ok && ncell == 32) {
                    uint4 v[2];
                    unsigned *w = reinterpret_cast<unsigned *>(v);
#pragma unroll
                    for (int b = 0; b < 8; b++)   // 4 cells per word: bit k -> byte k (the shifted copies do not overlap)
                        w[b] = (((bits >> (4 * b)) & 15u) * 0x00204081u) & 0x01010101u;
                    reinterpret_cast<uint4 *>(out)[0] = v[0];
                    reinterpret_cast<uint4 *>(out)[1] = v[1];
                } else {
                    for (int k = 0; k < ncell; k++) out[k] = (bits >> k) & 1u;
                }
            }
            __syncthreads();
        }
    }
}

int launch_arm_grid(int M, const double *theta, int row0, int n_rows, int n_links, const double *link_host,
                    const double *obstacles, int S, int O, uint8_t *grid, cudaStream_t s, int cells_only) {
    ArmParams p;
    p.M = M; p.row0 = row0; p.n_rows = n_rows; p.n_links = n_links; p.S = S; p.O = O;
    for (int k = 0; k < ARM_MAX_LINKS; k++) p.link[k] = k < n_links ? link_host[k] : 0.0;
    p.stretch = 1;
    for (int k = 1; k < n_links; k++)
        if (!(link_host[k] > 0.0)) p.stretch = 0;
    size_t smem = (size_t)S * O * 4 * sizeof(double) + (((size_t)S + 15) & ~(size_t)15);
    if (smem > 200 * 1024) return set_error(RRTK_ERR_INVALID, "S * O circles do not fit in shared memory (max 6400)");
    if (!cells_only) {   // the row rasteriser (every input: what it cannot rasterise it evaluates cell by cell)
        const size_t base_r = (size_t)S * O * (4 * sizeof(double) + sizeof(int4)) + (((size_t)2 * S + 15) & ~(size_t)15);
        const size_t W = ((size_t)M + 31) / 32;
        // row bitmaps of as many sets at a time as fit beside the circles (64 KB keeps three CTAs on an SM)
        size_t budget = 64 * 1024;
        if (base_r + W * 4 > budget) budget = 200 * 1024;
        long long G = base_r + W * 4 <= budget ? (long long)((budget - base_r) / (W * 4)) : 0;
        if (G > S) G = S;
        if (G >= 1) {
            const size_t smem_r = base_r + (size_t)G * W * 4;
            cudaError_t e = cudaFuncSetAttribute(arm_grid_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
            if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(arm_grid_rows_kernel)");
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, arm_grid_rows_kernel, ARM_RASTER_THREADS, smem_r);
            if (per_sm < 1) per_sm = 1;
            long long grid_dim = (long long)sms * per_sm;
            if (grid_dim > n_rows) grid_dim = n_rows;
            const int vec_ok = (M % 32 == 0) && (((uintptr_t)grid) % 16 == 0);
            arm_grid_rows_kernel<<<(unsigned)grid_dim, ARM_RASTER_THREADS, smem_r, s>>>(p, theta, obstacles, grid, vec_ok, (int)G);
            e = cudaGetLastError();
            if (e != cudaSuccess) return set_cuda_error(e, "arm_grid_rows_kernel launch");
            return RRTK_OK;
        }
    }
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ARM_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        long long want = (long long)n_rows * ((M + ARM_THREADS - 1) / ARM_THREADS);   // row tiles
        long long grid_dim = (long long)sms * per_sm;  // persistent, a multiple of the SM count
        if (grid_dim > want) grid_dim = want;
        if (grid_dim < 1) grid_dim = 1;
        kernel<<<(unsigned)grid_dim, ARM_THREADS, smem, s>>>(p, theta, obstacles, grid);
        return cudaSuccess;
    };
    cudaError_t e;
    if (n_links <= 2) e = launch(arm_grid_kernel<2>);
    else if (n_links <= 5) e = launch(arm_grid_kernel<5>);
    else if (n_links <= 8) e = launch(arm_grid_kernel<8>);
    else e = launch(arm_grid_kernel<ARM_MAX_LINKS>);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(arm_grid_kernel)");
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "arm_grid_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
