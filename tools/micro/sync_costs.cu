// Micro-costs of the primitives a hand-rolled grid exchange is made of (B200).  nvcc -arch=sm_100a sync_costs.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ldr(const uint4 *p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(uint4 *p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ unsigned ldr32(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str32(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

__global__ void k_fence(int mode, int iters, unsigned *buf, long long *out) {
    long long t0 = clock64();
    for (int k = 0; k < iters; k++) {
        if (mode & 4) buf[threadIdx.x + 32 * blockIdx.x] = k;  // an outstanding store before the fence
        if ((mode & 3) == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        else if ((mode & 3) == 2) __threadfence();
        else if ((mode & 3) == 3) asm volatile("fence.acq_rel.cta;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
// ping-pong between CTA 0 and CTA 1: one-way store -> remote poll latency (cycles per round trip / 2)
__global__ void k_pingpong(int iters, unsigned *flag, long long *out) {
    if (threadIdx.x != 0) return;
    unsigned *mine = flag + 64 * blockIdx.x, *other = flag + 64 * (1 - blockIdx.x);
    long long t0 = clock64();
    for (unsigned k = 1; k <= (unsigned)iters; k++) {
        if (blockIdx.x == 0) { str32(mine, k); while (ldr32(other) != k) { } }
        else { while (ldr32(other) != k) { } str32(mine, k); }
    }
    long long t1 = clock64();
    out[blockIdx.x] = (t1 - t0) / iters;
}
// all-to-all: every CTA publishes a tagged 16-byte piece, thread t polls CTA t's piece; G CTAs x T threads
__global__ void k_all2all(int iters, uint4 *rec, long long *out, int npiece, int fence) {
    const int G = gridDim.x, t = threadIdx.x;
    long long t0 = clock64();
    for (unsigned k = 1; k <= (unsigned)iters; k++) {
        uint4 *r = rec + (size_t)(k & 1) * G * 8;
        if (t < npiece) { if (fence) asm volatile("fence.acq_rel.gpu;" ::: "memory"); str(r + blockIdx.x * 8 + t, make_uint4(k, blockIdx.x, t, k)); }
        if (t < G) {
            for (int p = 0; p < npiece; p++) while (ldr(r + t * 8 + p).w != k) { }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (t == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
// counter barrier: atomicAdd + one poller per CTA
__global__ void k_counter(int iters, unsigned long long *ctr, long long *out, int fence) {
    const int G = gridDim.x;
    long long t0 = clock64();
    for (unsigned k = 1; k <= (unsigned)iters; k++) {
        if (threadIdx.x == 0) {
            if (fence == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (fence == 2) __threadfence();
            atomicAdd(ctr, 1ull);
            const unsigned long long target = (unsigned long long)k * G;
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
static void report(const char *name, long long *d_out, int G) {
    long long h[256];
    cudaMemcpy(h, d_out, sizeof(long long) * G, cudaMemcpyDeviceToHost);
    long long mn = h[0], mx = h[0];
    for (int i = 1; i < G; i++) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
    printf("%-44s min %6lld  max %6lld cycles\n", name, mn, mx);
}
int main() {
    unsigned *buf; long long *out; uint4 *rec; unsigned long long *ctr;
    cudaMalloc(&buf, 1 << 20); cudaMalloc(&out, 256 * 8); cudaMalloc(&rec, 2 * 256 * 8 * 16); cudaMalloc(&ctr, 256);
    const int it = 2000;
    const char *fn[] = {"no fence", "fence.acq_rel.gpu (MEMBAR.ALL.GPU)", "__threadfence (MEMBAR.SC.GPU)", "fence.acq_rel.cta"};
    for (int st = 0; st < 2; st++) for (int m = 0; m < 4; m++) {
        char name[96]; snprintf(name, 96, "%s%s", fn[m], st ? " after a store" : "");
        k_fence<<<1, 32>>>(m | (st << 2), it, buf, out); cudaDeviceSynchronize(); report(name, out, 1);
    }
    { char name[96]; snprintf(name, 96, "fence.acq_rel.gpu after a store, 148 CTAs x 32"); k_fence<<<148, 32>>>(1 | 4, it, buf, out); cudaDeviceSynchronize(); report(name, out, 148); }
    cudaMemset(buf, 0, 1 << 20);
    k_pingpong<<<2, 32>>>(it, buf, out); cudaDeviceSynchronize(); report("ping-pong round trip (2 one-way hops)", out, 2);
    int grids[] = {2, 16, 74, 148};
    for (int g : grids) for (int np = 1; np <= 5; np += 4) for (int f = 0; f < 2; f++) {
        cudaMemset(rec, 0, 2 * 256 * 8 * 16);
        void *args[] = {(void *)&it, (void *)&rec, (void *)&out, (void *)&np, (void *)&f};
        cudaLaunchCooperativeKernel((const void *)k_all2all, dim3(g), dim3(512), args, 0, 0); cudaDeviceSynchronize();
        char name[96]; snprintf(name, 96, "all-to-all tagged pieces G=%d pieces=%d fence=%d", g, np, f); report(name, out, g);
    }
    for (int g : grids) for (int f = 0; f < 3; f++) {
        cudaMemset(ctr, 0, 256);
        void *args[] = {(void *)&it, (void *)&ctr, (void *)&out, (void *)&f};
        cudaLaunchCooperativeKernel((const void *)k_counter, dim3(g), dim3(512), args, 0, 0); cudaDeviceSynchronize();
        char name[96]; snprintf(name, 96, "counter barrier G=%d fence=%d", g, f); report(name, out, g);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
