// Issue-rate microbenchmark of the integer ops the SGBM kernels are made of (B200, sm_100a).
// Each kernel runs 8 independent dependency chains per thread of ONE op; prints warp-instructions / clk / SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates tools/microbench/alu_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    if (OP == 0) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == 2) asm volatile("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == 3) r = __vadd2(a, b);
    if (OP == 4) r = __vmins2(a, b);
    if (OP == 5) r = __viaddmin_s16x2(a, b, c);
    if (OP == 6) r = __vimin3_s16x2(a, b, c);
    if (OP == 7) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == 8) r = __viaddmax_s16x2_relu(a, b, c);
    if (OP == 9) asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == 10) r = __vsub2(a, b);
    if (OP == 11) r = __vminu2(a, b);
    if (OP == 12) asm volatile("shf.r.wrap.b32 %0, %1, %2, 7;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <int OP, int MIX>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, long long* clk) {
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = seed * (threadIdx.x + i + 1);
    uint32_t b = seed ^ 0x00010001u, c = seed + 0x7fff7fffu;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = op<OP>(x[i], b, c);
            if (MIX >= 0) x[i] = op < MIX < 0 ? 0 : MIX > (x[i], c, b);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP, int MIX>
void run(const char* name, uint32_t* out, long long* clk) {
    int sms = 148;
    k<OP, MIX><<<sms, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    k<OP, MIX><<<sms, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += h[i];
    avg /= sms;
    double winstr = 32.0 * ITER * 8 * (MIX >= 0 ? 2 : 1);   // warp-instructions per SM (32 warps)
    printf("%-34s %6.2f warp-instr/clk/SM  (%.2f per scheduler)\n", name, winstr / avg, winstr / avg / 4);
}
int main() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    run<7, 0>("imad + add", out, clk);
    run<7, 1>("imad + lop3", out, clk);
    run<7, 2>("imad + prmt", out, clk);
    run<7, 3>("imad + vadd2", out, clk);
    run<7, 4>("imad + vmins2", out, clk);
    run<7, 5>("imad + viaddmin", out, clk);
    run<7, 6>("imad + vimin3", out, clk);
    run<7, 8>("imad + viaddmax_relu", out, clk);
    run<7, 9>("imad + min.s32", out, clk);
    run<7, 10>("imad + vsub2", out, clk);
    run<7, 11>("imad + vminu2", out, clk);
    run<7, 12>("imad + shf", out, clk);
    run<2, 0>("prmt + add", out, clk);
    run<2, 1>("prmt + lop3", out, clk);
    run<2, 3>("prmt + vadd2", out, clk);
    run<2, 4>("prmt + vmins2", out, clk);
    run<2, 5>("prmt + viaddmin", out, clk);
    run<2, 6>("prmt + vimin3", out, clk);
    run<2, 8>("prmt + viaddmax_relu", out, clk);
    run<2, 9>("prmt + min.s32", out, clk);
    run<2, 10>("prmt + vsub2", out, clk);
    run<2, 11>("prmt + vminu2", out, clk);
    run<2, 12>("prmt + shf", out, clk);
    run<3, 4>("vadd2 + vmins2", out, clk);
    run<3, 5>("vadd2 + viaddmin", out, clk);
    run<3, 6>("vadd2 + vimin3", out, clk);
    run<4, 5>("vmins2 + viaddmin", out, clk);
    run<1, 5>("lop3 + viaddmin", out, clk);
    run<1, 4>("lop3 + vmins2", out, clk);
    run<12, 5>("shf + viaddmin", out, clk);
    run<0, 5>("add + viaddmin", out, clk);
    run<0, 4>("add + vmins2", out, clk);
    run<3, 10>("vadd2 + vsub2", out, clk);
    return 0;
}
