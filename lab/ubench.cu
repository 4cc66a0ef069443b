// Integer-pipe microbenchmarks for B200 (sm_100a): issue rates of the instructions the field arithmetic is made of.
// Prints thread-ops per clock per SM for each instruction kind (148 SMs, all resident warps busy).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define ITERS 4096
#define CHAINS 8

template <int KIND>
__global__ void __launch_bounds__(256) bench(u64* out, u32 b) {
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 acc[CHAINS]; u32 a32[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { acc[k] = t * 77 + k; a32[k] = t + k * 3; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a32[k]), "r"(b));
            if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a32[k]) : "r"(b), "r"(t));
            if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a32[k]) : "r"(b), "r"(t));
            if (KIND == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a32[k]) : "r"(t), "r"(b));
            if (KIND == 4) {  // 64-bit add with carry chain: add.cc + addc
                u32 lo = (u32)acc[k], hi = (u32)(acc[k] >> 32);
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(b), "r"(t));
                acc[k] = ((u64)hi << 32) | lo;
            }
            if (KIND == 5) {  // mix: 1 wide mad + 2 carry adds (like the field mul)
                u32 lo = (u32)acc[k], hi = (u32)(acc[k] >> 32);
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a32[k]), "r"(b));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(b), "r"(t));
                a32[k] ^= lo + hi;
            }
            if (KIND == 6) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a32[k]) : "r"(b));
            if (KIND == 7) asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(acc[k]) : "r"(a32[k]));
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += acc[k] + a32[k];
    out[t] = s;
}

template <int KIND>
void run(const char* name, int ops_per_iter, u64* d) {
    int blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<KIND><<<blocks, threads>>>(d, 12345u);
    cudaEventRecord(e0);
    bench<KIND><<<blocks, threads>>>(d, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * ITERS * CHAINS * ops_per_iter;
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double per_clk_sm = ops / (ms * 1e-3) / (clk_khz * 1e3) / 148.0;
    printf("%-28s %8.3f ms  %7.2f T thread-ops/s  %6.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / 1e12, per_clk_sm, clk_khz / 1000);
}
int main() {
    u64* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    run<0>("IMAD.WIDE.U32 (reg x reg)", 1, d);
    run<7>("IMAD.WIDE.U32 (reg x imm)", 1, d);
    run<1>("IMAD.LO", 1, d);
    run<6>("IMAD.HI (mul.hi.u32)", 1, d);
    run<2>("LOP3", 1, d);
    run<3>("SHF", 1, d);
    run<4>("IADD3 + IADD3.X (64b add)", 2, d);
    run<5>("1 WIDE + 2 carry adds", 3, d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
