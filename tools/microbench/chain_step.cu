// Latency of ONE dependent SGM path step (previous pixel -> minimum over D -> this pixel) for a single row chain, the
// quantity that bounds a horizontal path fused into a strip sweep: a strip's row is COLS such steps in sequence.
// D = 128 disparities over G lanes x NP packed int16x2 registers. Prints clocks per step for
//   8 lanes x 8 regs (the layout of the row-parallel kernels), xor-shuffle reduction
//   16 lanes x 4 regs, 32 lanes x 2 regs, xor-shuffle reduction
//   32 lanes x 2 regs with redux.sync (one instruction for the cross-lane minimum)
//   the same with the minimum taken one step late (exact: see sgm_step_lag in csrc/sgbm.cu)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_step tools/microbench/chain_step.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

template <int NP, int G, int RED>
__device__ __forceinline__ void step(uint32_t (&P)[NP], uint32_t& m, const uint32_t (&c)[NP], uint32_t P1x2, uint32_t P2x2, uint32_t PADx2, int lg) {
    uint32_t up = __shfl_up_sync(0xffffffffu, P[NP - 1], 1, G);
    uint32_t dn = __shfl_down_sync(0xffffffffu, P[0], 1, G);
    if (lg == 0) up = PADx2;
    if (lg == G - 1) dn = PADx2;
    uint32_t T[NP + 1];
    T[0] = prmt(up, P[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NP; k++) T[k] = prmt(P[k - 1], P[k], 0x5432);
    T[NP] = prmt(P[NP - 1], dn, 0x5432);
    const uint32_t negm = __vadd2(~m, 0x00010001u);
    uint32_t mm = 0x7FFF7FFFu;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        uint32_t t = __vmins2(T[k], T[k + 1]);
        t = __viaddmin_s16x2(t, P1x2, P[k]);
        t = __viaddmin_s16x2(t, negm, P2x2);
        P[k] = __vadd2(c[k], t);
        mm = __vmins2(mm, P[k]);
    }
    mm = __vmins2(mm, prmt(mm, mm, 0x1032));
    if (RED == 0) {
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) mm = __vmins2(mm, __shfl_xor_sync(0xffffffffu, mm, o, G));
    } else {
        uint32_t r;
        asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(mm));
        mm = r;
    }
    m = mm;
}

// exact reformulation with the cross-lane minimum off the critical path: B = C + A(L) needs no minimum, L' = min(B - m, C + P2),
// m' = min(min_d B - m, min_d C + P2)
template <int NP>
__device__ __forceinline__ void step_lag(uint32_t (&P)[NP], uint32_t& negm, const uint32_t (&c)[NP], uint32_t cminP2, uint32_t P1x2, uint32_t P2x2,
                                         uint32_t PADx2, int lg) {
    uint32_t up = __shfl_up_sync(0xffffffffu, P[NP - 1], 1);
    uint32_t dn = __shfl_down_sync(0xffffffffu, P[0], 1);
    if (lg == 0) up = PADx2;
    if (lg == 31) dn = PADx2;
    uint32_t T[NP + 1], B[NP];
    T[0] = prmt(up, P[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NP; k++) T[k] = prmt(P[k - 1], P[k], 0x5432);
    T[NP] = prmt(P[NP - 1], dn, 0x5432);
    uint32_t mm = 0x7FFF7FFFu;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        uint32_t t = __vmins2(T[k], T[k + 1]);
        t = __viaddmin_s16x2(t, P1x2, P[k]);
        B[k] = __vadd2(c[k], t);
        mm = __vmins2(mm, B[k]);
    }
    mm = __vmins2(mm, prmt(mm, mm, 0x1032));
    uint32_t r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(mm));
#pragma unroll
    for (int k = 0; k < NP; k++) P[k] = __viaddmin_s16x2(B[k], negm, __vadd2(c[k], P2x2));
    const uint32_t mnew = __viaddmin_s16x2(r, negm, cminP2);
    negm = __vadd2(~mnew, 0x00010001u);
}

template <int NP, int G, int RED>
__global__ void k_chain(int steps, long long* out, uint32_t* sink) {
    __shared__ uint32_t C[64][64];   // 64 pixels x 128 int16
    for (int i = threadIdx.x; i < 64 * 64; i += 32) ((uint32_t*)C)[i] = ((i * 2654435761u) >> 20) & 0x03FF03FFu;
    __syncwarp();
    const int lane = threadIdx.x, lg = lane & (G - 1);
    if (lane >= G) {   // the other lanes of the warp idle along (same instruction stream)
    }
    uint32_t P[NP], c[NP], cn[NP], m = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) { P[k] = 0; cn[k] = C[0][lg * NP + k]; }
    const uint32_t P1 = 0x02580258u, P2 = 0x09600960u, PAD = 0x7DA77DA7u;
    long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int k = 0; k < NP; k++) { c[k] = cn[k]; cn[k] = C[(s + 1) & 63][lg * NP + k]; }
        if constexpr (RED == 2) {
            uint32_t cm = c[0];
#pragma unroll
            for (int k = 1; k < NP; k++) cm = __vmins2(cm, c[k]);
            cm = __vmins2(cm, prmt(cm, cm, 0x1032));
            uint32_t r;
            asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(cm));
            step_lag<NP>(P, m, c, __vadd2(r, P2), P1, P2, PAD, lg);
        } else {
            step<NP, G, RED>(P, m, c, P1, P2, PAD, lg);
        }
    }
    long long t1 = clock64();
    uint32_t acc = m;
#pragma unroll
    for (int k = 0; k < NP; k++) acc ^= P[k];
    sink[threadIdx.x] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int NP, int G, int RED>
void run(const char* name) {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 256);
    const int steps = 20000;
    for (int r = 0; r < 2; r++) { k_chain<NP, G, RED><<<1, 32>>>(steps, out, sink); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %6.1f clocks per step (%s)\n", name, h / (double)steps, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(sink);
}

int main() {
    run<8, 8, 0>("8 lanes x 8 regs, shuffle reduction");
    run<4, 16, 0>("16 lanes x 4 regs, shuffle reduction");
    run<2, 32, 0>("32 lanes x 2 regs, shuffle reduction");
    run<2, 32, 1>("32 lanes x 2 regs, redux.sync.min");
    run<2, 32, 2>("32 lanes x 2 regs, redux.sync.min one step late");
    run<4, 32, 1>("32 lanes x 4 regs (D=256), redux.sync.min");
    run<4, 32, 2>("32 lanes x 4 regs (D=256), redux one step late");
    run<1, 32, 1>("32 lanes x 1 reg (D=64), redux.sync.min");
    return 0;
}
