// Latency of a flag-in-data hand-off between two CTAs on different SMs (B200): the strip-to-strip halo exchange of the
// fused sweep is exactly this. Prints nanoseconds per one-way hand-off for several store / load flavours, and for a
// 2-CTA cluster exchanging through distributed shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong tools/microbench/pingpong.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int MODE>
__device__ __forceinline__ void put(unsigned long long* p, unsigned long long v) {
    if (MODE == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 2) atomicExch(p, v);
    if (MODE == 3) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned long long get(const unsigned long long* p) {
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (MODE == 3) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// two CTAs (blockIdx 0 and `far`), one thread each: 0 sends i, the other answers i
template <int MODE, int SLEEP>
__global__ void k_pp(unsigned long long* a, unsigned long long* b, int iters, int far, long long* out) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; i++) {
            put<MODE>(a, (unsigned long long)i);
            while (get<MODE>(b) != (unsigned long long)i) if (SLEEP) __nanosleep(SLEEP);
        }
        out[0] = clock64() - t0;
    } else if ((int)blockIdx.x == far) {
        for (int i = 1; i <= iters; i++) {
            while (get<MODE>(a) != (unsigned long long)i) if (SLEEP) __nanosleep(SLEEP);
            put<MODE>(b, (unsigned long long)i);
        }
    }
}
// cluster of 2: each CTA polls a word in its OWN shared memory, the partner writes it remotely
__global__ void __cluster_dims__(2, 1, 1) k_pp_dsmem(int iters, long long* out) {
    __shared__ unsigned long long box;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    if (threadIdx.x == 0) box = 0;
    cl.sync();
    unsigned long long* remote = cl.map_shared_rank(&box, rank ^ 1);
    if (threadIdx.x == 0) {
        volatile unsigned long long* mine = &box;
        long long t0 = clock64();
        for (int i = 1; i <= iters; i++) {
            if (rank == 0) {
                *(volatile unsigned long long*)remote = (unsigned long long)i;
                while (*mine != (unsigned long long)i) {}
            } else {
                while (*mine != (unsigned long long)i) {}
                *(volatile unsigned long long*)remote = (unsigned long long)i;
            }
        }
        if (rank == 0) out[0] = clock64() - t0;
    }
    cl.sync();
}
template <int MODE, int SLEEP>
void run(const char* name, int far) {
    unsigned long long *a, *b; long long* out;
    cudaMalloc(&a, 256); cudaMalloc(&b, 256); cudaMalloc(&out, 8);
    const int iters = 20000;
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(a, 0, 256); cudaMemset(b, 0, 256);
        k_pp<MODE, SLEEP><<<148, 32>>>(a, b + 16, iters, far, out);   // words in different 128-byte lines
        cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-44s partner CTA %3d: %7.0f ns per one-way hand-off (%s)\n", name, far, h / (double)iters / 2 / (clk * 1e-6), cudaGetErrorString(cudaGetLastError()));
    cudaFree(a); cudaFree(b); cudaFree(out);
}
int main() {
    for (int far : {1, 2, 74, 147}) {
        run<0, 0>("st/ld.volatile", far);
        run<1, 0>("st/ld.relaxed.gpu", far);
        run<1, 40>("st/ld.relaxed.gpu + nanosleep(40)", far);
        run<2, 0>("atomicExch + ld.relaxed.gpu", far);
        run<3, 0>("st.release.gpu / ld.acquire.gpu", far);
    }
    long long* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    for (int rep = 0; rep < 2; rep++) { k_pp_dsmem<<<2, 32>>>(iters, out); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-44s %7.0f ns per one-way hand-off (%s)\n", "cluster of 2, distributed shared memory", h / (double)iters / 2 / (clk * 1e-6), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
