// rrtk_device.cuh -- shared device helpers: error plumbing, counter-based RNG, Sobol table.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rrtk.h"

namespace rrtk {

// ---- error plumbing (thread-local message, C-ABI status codes) ----
int set_error(int code, const char *msg);
int set_cuda_error(cudaError_t e, const char *where);

// ---- counter-based RNG for the in-kernel samplers (splitmix64 finaliser) ----
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// key of (seed, query, iteration); draws c = 0, 1, 2 are splitmix64(key + c)
__host__ __device__ __forceinline__ uint64_t rng_key(uint64_t seed, uint64_t q, uint64_t it) {
    return splitmix64(seed ^ splitmix64(q * 0x100000001B3ull + it * 0x9E3779B1ull + 0x51ull));
}
// 53-bit uniform in [0, 1), like CPython's random.random()
__host__ __device__ __forceinline__ double u01(uint64_t r) {
    return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

// ---- Sobol direction integers (rrt_04:320-364 data, :393-436 recurrence), built at compile time ----
constexpr int SOBOL_BITS = 30;
constexpr int SOBOL_DIM_MAX = 40;

struct SobolTable {
    uint32_t v[SOBOL_DIM_MAX][SOBOL_BITS];
};

constexpr SobolTable make_sobol_table() {
    constexpr int poly[SOBOL_DIM_MAX] = {1,   3,   7,   11,  13,  19,  25,  37,  59,  47,
                                         61,  55,  41,  67,  97,  91,  109, 103, 115, 131,
                                         193, 137, 145, 143, 241, 157, 185, 167, 229, 171,
                                         213, 191, 253, 203, 211, 239, 247, 285, 369, 299};
    // initial m_j, one row per bit column j = 1..7, starting at dimension index `first[j]`
    constexpr int first[8] = {0, 2, 3, 5, 7, 13, 19, 37};
    constexpr int init[8][38] = {
        {0},
        {1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3, 3, 1, 3, 1, 3, 1, 3, 1, 1, 3, 1, 3, 1, 3, 1, 3},
        {7, 5, 1, 3, 3, 7, 5, 5, 7, 7, 1, 3, 3, 7, 5, 1, 1, 5, 3, 3, 1, 7, 5, 1, 3, 3, 7, 5, 1, 1, 5, 7, 7, 5, 1, 3, 3},
        {1, 7, 9, 13, 11, 1, 3, 7, 9, 5, 13, 13, 11, 3, 15, 5, 3, 15, 7, 9, 13, 9, 1, 11, 7, 5, 15, 1, 15, 11, 5, 3, 1, 7, 9},
        {9, 3, 27, 15, 29, 21, 23, 19, 11, 25, 7, 13, 17, 1, 25, 29, 3, 31, 11, 5, 23, 27, 19, 21, 5, 1, 17, 13, 7, 15, 9, 31, 9},
        {37, 33, 7, 5, 11, 39, 63, 27, 17, 15, 23, 29, 3, 21, 13, 31, 25, 9, 49, 33, 19, 29, 11, 19, 27, 15, 25},
        {13, 33, 115, 41, 79, 17, 29, 119, 75, 73, 105, 7, 59, 65, 21, 3, 113, 61, 89, 45, 107},
        {7, 23, 39}};
    SobolTable t{};
    for (int d = 0; d < SOBOL_DIM_MAX; d++) {
        for (int j = 0; j < SOBOL_BITS; j++) t.v[d][j] = 0;
        t.v[d][0] = 1;
    }
    for (int c = 1; c < 8; c++)
        for (int d = first[c]; d < SOBOL_DIM_MAX; d++) t.v[d][c] = (uint32_t)init[c][d - first[c]];
    for (int j = 0; j < SOBOL_BITS; j++) t.v[0][j] = 1;
    for (int d = 1; d < SOBOL_DIM_MAX; d++) {
        int deg = 0;
        for (int pp = poly[d] >> 1; pp; pp >>= 1) deg++;
        for (int j = deg; j < SOBOL_BITS; j++) {
            uint32_t nv = t.v[d][j - deg];
            for (int k = 0; k < deg; k++)
                if ((poly[d] >> (deg - 1 - k)) & 1) nv ^= (2u << k) * t.v[d][j - k - 1];
            t.v[d][j] = nv;
        }
    }
    for (int d = 0; d < SOBOL_DIM_MAX; d++)
        for (int j = 0; j < SOBOL_BITS; j++) t.v[d][j] <<= (SOBOL_BITS - 1 - j);
    return t;
}

#ifdef __CUDACC__
__constant__ const SobolTable c_sobol = make_sobol_table();

// point `index` of the 2-D sequence, as 30-bit integers (Gray-code order, rrt_04:494-503)
__device__ __forceinline__ void sobol2(int64_t index, uint32_t &q0, uint32_t &q1) {
    uint64_t n = index < 0 ? 0ull : (uint64_t)index;
    uint64_t g = n ^ (n >> 1);
    q0 = 0; q1 = 0;
#pragma unroll 1
    for (int j = 0; g != 0 && j < SOBOL_BITS; j++, g >>= 1)
        if (g & 1ull) { q0 ^= c_sobol.v[0][j]; q1 ^= c_sobol.v[1][j]; }
}
#endif

}  // namespace rrtk
