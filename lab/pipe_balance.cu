// Does moving the carry-in-only steps of the field add/sub/reduce onto the FMA pipe (IMAD.X via madc.lo) pay on B200?
// Hypothesis from lab/ubench.cu: alu-pipe and 32-bit-IMAD instructions issue on alternate cycles (up to ~86 lanes/clk/SM
// measured), IMAD.WIDE blocks both for its two cycles, so a kernel costs ~2 x (WIDE + max(alu, fma32)) cycles per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-circuits_b200/csrc -o lab/pipe_balance lab/pipe_balance.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
using namespace zkb;
#define ITERS 512
#define CH 8

// Reference forms: every carry step as add.cc / addc / sub.cc / subc (ptxas: IADD3 / IADD3.X on the alu pipe) and a compare +
// select canonicalisation — what field.cuh used before. f_sub / f_add / f_mul / f_canon of field.cuh are the FMA-pipe forms.
__device__ __forceinline__ u64 ref_sub(u64 a, u64 b) {
    u32 o0, o1;
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t}"
        : "=&r"(o0), "=&r"(o1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)o1 << 32) | o0;
}
__device__ __forceinline__ u64 ref_add(u64 a, u64 b) {
    u32 o0, o1;
    asm("{\n\t.reg .u32 m, n0, n1;\n\t"
        "sub.cc.u32 n0, 1, %4;\n\t"
        "subc.u32 n1, 0xffffffff, %5;\n\t"
        "sub.cc.u32 %0, %2, n0;\n\t"
        "subc.cc.u32 %1, %3, n1;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t}"
        : "=&r"(o0), "=&r"(o1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)o1 << 32) | o0;
}
__device__ __forceinline__ u64 ref_mul(u64 a, u64 b) {
    u64 r = gl_mul_lazy(a, b);            // gl_mul128_limbs + gl_reduce_limbs: all carry steps on the alu pipe
    u32 lo = (u32)r, hi = (u32)(r >> 32);
    if (hi == 0xFFFFFFFFu && lo != 0) { lo -= 1; hi = 0; }
    return ((u64)hi << 32) | lo;
}
#define f_sub_v2 f_sub
#define f_add_v2 f_add
#define f_mul_v2 f_mul

// every variant against the reference forms on edge values and a pseudo-random stream; *bad counts mismatches
__global__ void check_kernel(unsigned long long* bad) {
    const u64 edge[12] = {0, 1, 2, GL_P - 1, GL_P - 2, 0xffffffffull, 0x100000000ull, 0x100000001ull, 0x8000000000000000ull,
                          GL_P - 0x100000000ull, 0xfffffffe00000002ull, 0x00000001ffffffffull};
    const u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 z = 0x9e3779b97f4a7c15ull * (t + 1);
    for (int i = 0; i < 256; ++i) {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
        u64 a = (i < 144) ? edge[i % 12] : gl_canon(z), b = (i < 144) ? edge[(i / 12) % 12] : gl_canon(z * 0x2545f4914f6cdd1dull + t);
        if (f_sub(a, b) != ref_sub(a, b) || f_add(a, b) != ref_add(a, b) || f_mul(a, b) != ref_mul(a, b)) atomicAdd(bad, 1ull);
    }
}

template <int V>
__global__ void __launch_bounds__(256) bench(u64* out, u64 seed) {
    const u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 a[CH], b[CH], w[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { a[k] = gl_canon(seed * (t + 1) + k); b[k] = gl_canon(seed + 977u * t + k); w[k] = gl_canon(seed ^ (0x9e3779b97f4a7c15ull * (k + 1))); }
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (V == 0) { u64 s = ref_add(a[k], b[k]), d = ref_sub(a[k], b[k]); a[k] = s; b[k] = d; }                   // butterflies only, alu forms
            if (V == 1) { u64 s = f_add(a[k], b[k]), d = f_sub(a[k], b[k]); a[k] = s; b[k] = d; }                       // butterflies only, FMA-pipe forms
            if (V == 2) { u64 s = ref_add(a[k], b[k]), d = ref_sub(a[k], b[k]); a[k] = s; b[k] = ref_mul(d, w[k]); }    // + one multiply, alu forms
            if (V == 3) { u64 s = f_add(a[k], b[k]), d = f_sub(a[k], b[k]); a[k] = s; b[k] = ref_mul(d, w[k]); }        // FMA-pipe add/sub, alu multiply
            if (V == 4) { u64 s = f_add(a[k], b[k]), d = f_sub(a[k], b[k]); a[k] = s; b[k] = f_mul(d, w[k]); }          // all FMA-pipe forms
            if (V == 5) { a[k] = ref_mul(a[k], w[k]); }                                                                // multiplies only, alu forms
            if (V == 6) { a[k] = f_mul(a[k], w[k]); }                                                                  // multiplies only, FMA-pipe forms
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s ^= a[k] + 3 * b[k];
    out[t] = s;
}

template <int V>
static void run(const char* name, u64* out, u64* check) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    bench<V><<<blocks, 256>>>(out, 12345);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) bench<V><<<blocks, 256>>>(out, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    u64 h[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double units = 5.0 * blocks * 256 * (double)ITERS * CH;
    std::printf("%-46s %8.3f ms  %7.2f G iterations/s  %6.2f clk/SM per warp-iteration   check %016llx\n", name, ms, units / ms * 1e-6,
                ms * 1e-3 * 1.965e9 * 148 / (units / 32), (unsigned long long)h[1]);
    *check = h[1];
}

int main() {
    u64* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(u64));
    {
        unsigned long long* bad; unsigned long long h = 0;
        cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
        check_kernel<<<64, 128>>>(bad);
        cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
        std::printf("FMA-pipe forms (field.cuh) vs alu forms: %llu mismatches in %d cases\n", h, 64 * 128 * 256);
    }
    u64 c[7];
    run<0>("butterfly (add + sub), alu forms", out, &c[0]);
    run<1>("butterfly (add + sub), FMA-pipe forms", out, &c[1]);
    run<2>("butterfly + multiply, alu forms", out, &c[2]);
    run<3>("FMA-pipe butterfly + alu multiply", out, &c[3]);
    run<4>("butterfly + multiply, FMA-pipe forms", out, &c[4]);
    run<5>("multiply only, alu forms", out, &c[5]);
    run<6>("multiply only, FMA-pipe forms", out, &c[6]);
    std::printf("results agree: %s\n", (c[0] == c[1] && c[2] == c[3] && c[3] == c[4] && c[5] == c[6]) ? "yes" : "NO");
    return 0;
}
