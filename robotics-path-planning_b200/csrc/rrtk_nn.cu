// rrtk_nn.cu -- brute-force nearest-node and near-radius search over a large SoA float2 node array
// (get_nearest_node_index rrt_04:1196-1202 / rrt_07:1210-1214, find_near_nodes rrt_04:1314-1338 /
// rrt_07:1137-1143) for the large-tree mode (BASELINE config 3: one tree of up to 10^6..10^8 nodes,
// B samples per pass).  HBM-bound streaming kernels: 8 bytes per node per pass, independent of B.
//
//   * persistent grid (multiple of the SM count), 16-byte loads (two float2 nodes per LDG.128),
//     4 loads in flight per thread, per-thread running best for each of the B samples;
//   * warp-shuffle argmin, one 64-bit atomicMin per warp and sample on a packed (d2 bits, index)
//     key: the minimum distance wins and exact ties resolve to the LOWEST index, like list.index();
//   * near: same scan, ballot + popc compaction per warp, one atomicAdd per warp to reserve output
//     slots (output order is per-warp ascending; sort by index on the consumer side when order matters).
// FP32 here is a FILTER for the FP64 planner: the caller re-checks candidates whose distance is within
// the float rounding bound of the decision in FP64 (see DESIGN.md "Large-tree mode").
#include "rrtk_device.cuh"

namespace rrtk {

constexpr int NN_THREADS = 256;
constexpr int NN_MAX_B = 8;  // samples per pass held in registers

__device__ __forceinline__ unsigned long long pack_key(float d2, unsigned idx) {
    // d2 >= 0 so its bit pattern is monotone in its value
    return ((unsigned long long)__float_as_uint(d2) << 32) | idx;
}

template <int B>
__global__ void __launch_bounds__(NN_THREADS)
nearest_kernel(const float4 *__restrict__ xy2, long long n_pairs, const float2 *__restrict__ tail,
               long long n, const float2 *__restrict__ samples, unsigned long long *__restrict__ best) {
    float sx[B], sy[B], bd[B];
    unsigned bi[B];
#pragma unroll
    for (int b = 0; b < B; b++) {
        float2 s = samples[b];
        sx[b] = s.x; sy[b] = s.y; bd[b] = CUDART_INF_F; bi[b] = 0xffffffffu;
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // main loop: 4 independent 16-byte loads in flight per thread
    for (; i + 3 * stride < n_pairs; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldcs(&xy2[i + u * stride]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            unsigned idx = (unsigned)(2 * (i + u * stride));
#pragma unroll
            for (int b = 0; b < B; b++) {
                float dx0 = v[u].x - sx[b], dy0 = v[u].y - sy[b];
                float d0 = fmaf(dx0, dx0, dy0 * dy0);
                float dx1 = v[u].z - sx[b], dy1 = v[u].w - sy[b];
                float d1 = fmaf(dx1, dx1, dy1 * dy1);
                if (d0 < bd[b]) { bd[b] = d0; bi[b] = idx; }
                if (d1 < bd[b]) { bd[b] = d1; bi[b] = idx + 1; }
            }
        }
    }
    for (; i < n_pairs; i += stride) {
        float4 v = __ldcs(&xy2[i]);
        unsigned idx = (unsigned)(2 * i);
#pragma unroll
        for (int b = 0; b < B; b++) {
            float dx0 = v.x - sx[b], dy0 = v.y - sy[b];
            float d0 = fmaf(dx0, dx0, dy0 * dy0);
            float dx1 = v.z - sx[b], dy1 = v.w - sy[b];
            float d1 = fmaf(dx1, dx1, dy1 * dy1);
            if (d0 < bd[b]) { bd[b] = d0; bi[b] = idx; }
            if (d1 < bd[b]) { bd[b] = d1; bi[b] = idx + 1; }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail node
        float2 t = tail[0];
#pragma unroll
        for (int b = 0; b < B; b++) {
            float dx = t.x - sx[b], dy = t.y - sy[b];
            float d = fmaf(dx, dx, dy * dy);
            if (d < bd[b]) { bd[b] = d; bi[b] = (unsigned)(n - 1); }
        }
    }
    // a thread visits indices in ascending order, so strict '<' already kept its lowest index;
    // across threads the packed key orders by (d2, index)
#pragma unroll
    for (int b = 0; b < B; b++) {
        unsigned long long key = pack_key(bd[b], bi[b]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
            key = o < key ? o : key;
        }
        if ((threadIdx.x & 31) == 0) atomicMin(&best[b], key);
    }
}

__global__ void nearest_unpack_kernel(const unsigned long long *best, int B, int *idx, float *d2) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    unsigned long long k = best[b];
    idx[b] = (int)(unsigned)(k & 0xffffffffull);
    d2[b] = __uint_as_float((unsigned)(k >> 32));
}

// near-radius search for ONE centre: out_idx gets every i with d2 <= r2 (unordered), count in *out_n.
// Four 16-byte loads in flight per thread like nearest_kernel (one load per round left the scan at ~60 % of the HBM rate);
// hits are rare for the radii a planner asks for, so a round first votes whether ANY lane has one and only then compacts.
__global__ void __launch_bounds__(NN_THREADS)
near_kernel(const float4 *__restrict__ xy2, long long n_pairs, const float2 *__restrict__ tail, long long n,
            float cx, float cy, float r2, int *__restrict__ out_idx, int cap, int *__restrict__ out_n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rounds = (n_pairs + 4 * stride - 1) / (4 * stride);   // the same for every thread: the votes are warp-wide
    for (long long r = 0; r < rounds; r++) {
        const long long i0 = base + 4 * r * stride;
        float4 v[4];
        unsigned hits = 0u;   // bit 2u / 2u + 1: first / second node of load u
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * stride < n_pairs) v[u] = __ldcs(&xy2[i0 + u * stride]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i0 + u * stride < n_pairs) {
                const float dx0 = v[u].x - cx, dy0 = v[u].y - cy, dx1 = v[u].z - cx, dy1 = v[u].w - cy;
                hits |= (fmaf(dx0, dx0, dy0 * dy0) <= r2 ? 1u : 0u) << (2 * u);
                hits |= (fmaf(dx1, dx1, dy1 * dy1) <= r2 ? 2u : 0u) << (2 * u);
            }
        }
        if (!__any_sync(0xffffffffu, hits != 0u)) continue;
        const int mine = __popc(hits);
        int incl = mine;   // inclusive prefix sum of the lanes' hit counts
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        int slot = 0;
        if (lane == 0) slot = atomicAdd(out_n, tot);
        slot = __shfl_sync(0xffffffffu, slot, 0) + incl - mine;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * stride;
            if (hits & (1u << (2 * u))) { if (slot < cap) out_idx[slot] = (int)(2 * i); slot++; }
            if (hits & (2u << (2 * u))) { if (slot < cap) out_idx[slot] = (int)(2 * i + 1); slot++; }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        float2 t = tail[0];
        float dx = t.x - cx, dy = t.y - cy;
        if (fmaf(dx, dx, dy * dy) <= r2) {
            int p = atomicAdd(out_n, 1);
            if (p < cap) out_idx[p] = (int)(n - 1);
        }
    }
}

static int nn_grid(const void *kernel) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NN_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;  // a multiple of the SM count
}

template <int B>
static int launch_nearest_b(const float *xy, long long n, const float *samples, unsigned long long *best,
                            cudaStream_t s) {
    long long n_pairs = n / 2;
    int grid = nn_grid((const void *)nearest_kernel<B>);
    long long want = (n_pairs + NN_THREADS - 1) / NN_THREADS;
    if (want < 1) want = 1;
    if (grid > want) grid = (int)want;
    nearest_kernel<B><<<grid, NN_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(xy), n_pairs,
                                                   reinterpret_cast<const float2 *>(xy) + (n - 1), n,
                                                   reinterpret_cast<const float2 *>(samples), best);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "nearest_kernel launch");
    return RRTK_OK;
}

// xy: [n][2] float32 (16-byte aligned), samples: [B][2] float32, scratch: [B] uint64 (device)
int launch_nearest(const float *xy, long long n, const float *samples, int B, unsigned long long *scratch,
                   int *idx, float *d2, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(scratch, 0xff, sizeof(unsigned long long) * B, s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
    int rc = RRTK_OK;
    int done = 0;
    while (done < B && rc == RRTK_OK) {  // groups of up to NN_MAX_B samples share one pass
        int b = B - done;
        const float *sp = samples + 2 * done;
        unsigned long long *bp = scratch + done;
        if (b >= NN_MAX_B) { rc = launch_nearest_b<NN_MAX_B>(xy, n, sp, bp, s); done += NN_MAX_B; }
        else if (b >= 4) { rc = launch_nearest_b<4>(xy, n, sp, bp, s); done += 4; }
        else if (b >= 2) { rc = launch_nearest_b<2>(xy, n, sp, bp, s); done += 2; }
        else { rc = launch_nearest_b<1>(xy, n, sp, bp, s); done += 1; }
    }
    if (rc) return rc;
    nearest_unpack_kernel<<<(B + 63) / 64, 64, 0, s>>>(scratch, B, idx, d2);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "nearest_unpack_kernel launch");
    return RRTK_OK;
}

int launch_near(const float *xy, long long n, float cx, float cy, float r2, int *out_idx, int cap, int *out_n,
                cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_n, 0, sizeof(int), s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
    long long n_pairs = n / 2;
    int grid = nn_grid((const void *)near_kernel);
    long long want = (n_pairs + NN_THREADS - 1) / NN_THREADS;
    if (want < 1) want = 1;
    if (grid > want) grid = (int)want;
    near_kernel<<<grid, NN_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(xy), n_pairs,
                                            reinterpret_cast<const float2 *>(xy) + (n - 1), n, cx, cy, r2,
                                            out_idx, cap, out_n);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "near_kernel launch");
    return RRTK_OK;
}

// ---- pipe-peak probes: dependent-free FMA chains, for the FP64 / FP32 roofline denominators ----
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, T *out) {
    T a[8], b = (T)1.000000119, c = (T)1e-9;
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == (T)-1) out[0] = s;  // never true; keeps the chains alive
}

int launch_fma_peak(int fp64, int iters, int blocks, void *out, cudaStream_t s) {
    if (fp64) fma_peak_kernel<double><<<blocks, 256, 0, s>>>(iters, (double *)out);
    else fma_peak_kernel<float><<<blocks, 256, 0, s>>>(iters, (float *)out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "fma_peak_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
