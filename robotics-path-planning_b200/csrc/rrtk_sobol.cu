// rrtk_sobol.cu -- Sobol sample generator: i4_sobol (rrt_04:230-503) in closed form.
// One thread per (point, dimension-group); the direction table sits in constant memory.
#include "rrtk_device.cuh"

namespace rrtk {

extern "C" __global__ void sobol_fill_kernel(int dim, int64_t first, int64_t count, double *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int64_t idx = first + i;
    uint64_t n = idx < 0 ? 0ull : (uint64_t)idx;
    uint64_t g = n ^ (n >> 1);
    const double recipd = 1.0 / 1073741824.0;  // rrt_04:440
    for (int d = 0; d < dim; d++) {
        uint32_t q = 0;
        uint64_t gg = g;
        for (int j = 0; gg != 0 && j < SOBOL_BITS; j++, gg >>= 1)
            if (gg & 1ull) q ^= c_sobol.v[d][j];
        out[i * dim + d] = (double)q * recipd;
    }
}

int launch_sobol_fill(int dim, int64_t first, int64_t count, double *out, cudaStream_t s) {
    int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    sobol_fill_kernel<<<(unsigned)blocks, threads, 0, s>>>(dim, first, count, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "sobol_fill_kernel launch");
    return RRTK_OK;
}

}  // namespace rrtk
