// Integer-pipe microbenchmarks for B200 (sm_100a): issue rates of the instructions the field arithmetic is made of.
// Prints thread-ops per clock per SM for each instruction kind (148 SMs, 8 CTAs x 256 threads per SM, 8 independent
// dependency chains per thread). Every op's multiplier/addend depends on the previous result of its chain so that
// ptxas cannot hoist or strength-reduce it (checked in SASS: cuobjdump -sass lab/ubench).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lab/ubench lab/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define ITERS 2048
#define CHAINS 8

template <int KIND>
__global__ void __launch_bounds__(256) bench(u64* out, u32 b) {
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u32 lo[CHAINS], hi[CHAINS], x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { lo[k] = t * 77 + k; hi[k] = t ^ (k * 977); x[k] = t + k * 3 + b; }
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (KIND == 0) {   // IMAD.WIDE.U32 reg x reg + 64-bit acc; multiplier = previous low word
                u64 acc = ((u64)hi[k] << 32) | lo[k];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(lo[k]), "r"(x[k]));
                lo[k] = (u32)acc; hi[k] = (u32)(acc >> 32);
            }
            if (KIND == 1) {   // IMAD.WIDE.U32 reg x small immediate
                u64 acc = ((u64)hi[k] << 32) | lo[k];
                asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(acc) : "r"(lo[k]));
                lo[k] = (u32)acc; hi[k] = (u32)(acc >> 32);
            }
            if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[k]) : "r"(x[k]), "r"(hi[k]));
            if (KIND == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(x[k]));
            if (KIND == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[k]) : "r"(x[k]), "r"(hi[k]));
            if (KIND == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(lo[k]) : "r"(hi[k]), "r"(x[k]));
            if (KIND == 6)     // 64-bit add with carry: IADD3 + IADD3.X
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(x[k]), "r"(lo[k]));
            if (KIND == 7) {   // 1 IMAD.WIDE + 1 64-bit carry add (1 FMA-pipe : 2 ALU-pipe)
                u64 acc = ((u64)hi[k] << 32) | lo[k];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(lo[k]), "r"(x[k]));
                lo[k] = (u32)acc; hi[k] = (u32)(acc >> 32);
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[k]), "+r"(hi[k]) : "r"(lo[k]), "r"(x[k]));
            }
            if (KIND == 8) {   // 1 IMAD.WIDE + 1 LOP3 (1 FMA-pipe : 1 ALU-pipe)
                u64 acc = ((u64)hi[k] << 32) | lo[k];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(lo[k]), "r"(x[k]));
                lo[k] = (u32)acc; hi[k] = (u32)(acc >> 32);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(lo[k]), "r"(hi[k]));
            }
            if (KIND >= 10 && KIND <= 14) {   // FP64 pipe: DFMA alone, and interleaved 1:1 with an integer instruction
                double dacc = __longlong_as_double(((long long)hi[k] << 32) | lo[k]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dacc) : "d"(1.0000001), "d"(0.5));
                long long bits = __double_as_longlong(dacc);
                lo[k] = (u32)bits; hi[k] = (u32)(bits >> 32);
                if (KIND == 11) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(x[k]), "r"(b));
                if (KIND == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(x[k]), "r"(b));
                if (KIND == 13) {
                    u64 acc = ((u64)b << 32) | x[k];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[k]), "r"(b));
                    x[k] = (u32)acc ^ (u32)(acc >> 32);
                }
                if (KIND == 14) {   // 1 DFMA : 1 IMAD : 1 LOP3
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(x[k]), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(x[k]), "r"(b));
                }
            }
            if (KIND == 9) {   // predicated 64-bit fix-up: setp + 2 predicated adds
                asm volatile("{.reg .pred p;\n\tsetp.lt.u32 p, %0, %2;\n\t@p add.cc.u32 %0, %0, %3;\n\t@p addc.u32 %1, %1, 0;}"
                             : "+r"(lo[k]), "+r"(hi[k]) : "r"(x[k]), "r"(hi[k]));
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += (((u64)hi[k] << 32) | lo[k]) + x[k];
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
    printf("%-44s %8.3f ms  %7.2f T thread-instr/s  %6.1f instr/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / 1e12, per_clk_sm, clk_khz / 1000);
}
int main() {
    u64* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    run<0>("IMAD.WIDE.U32 (reg x reg + acc64)", 1, d);
    run<1>("IMAD.WIDE.U32 (reg x imm + acc64)", 1, d);
    run<2>("IMAD (lo)", 1, d);
    run<3>("IMAD.HI.U32", 1, d);
    run<4>("LOP3", 1, d);
    run<5>("SHF", 1, d);
    run<6>("IADD3 + IADD3.X (64-bit add)", 2, d);
    run<7>("1 IMAD.WIDE + IADD3 + IADD3.X", 3, d);
    run<8>("1 IMAD.WIDE + 1 LOP3", 2, d);
    run<9>("ISETP + 2 predicated IADD3", 3, d);
    run<10>("DFMA only", 1, d);
    run<11>("1 DFMA + 1 IMAD(lo)", 2, d);
    run<12>("1 DFMA + 1 LOP3", 2, d);
    run<13>("1 DFMA + 1 IMAD.WIDE (+ LOP3)", 3, d);
    run<14>("1 DFMA + 1 IMAD(lo) + 1 LOP3", 3, d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
