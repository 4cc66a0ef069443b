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

#if !defined(__CUDA_ARCH__)   // host pass: the device helpers of field.cuh do not exist
namespace zkb { inline void gl_unpack(u64, u32&, u32&) {} inline u64 gl_pack(u32, u32) { return 0; } inline void gl_wide(u32, u32, u32&, u32&) {} }
#endif
// Carry-flag polarity: after sub.cc / subc.cc the flag the next carry-consuming instruction sees is the HARDWARE carry of
// a + ~b + 1, i.e. 1 = NO borrow (subc undoes that itself; a madc after a sub.cc does not). The forms below rely on it and are
// checked against the reference forms by check_kernel.
// a - b mod p, canonical in/out: IADD3, IADD3.X, IMAD.X (mask), IADD3, IMAD.X  = 3 alu + 2 fma (reference form: 4 + 1)
__device__ __forceinline__ u64 f_sub_v2(u64 a, u64 b) {
    u32 a0, a1, b0, b1, o0, o1;
    gl_unpack(a, a0, a1); gl_unpack(b, b0, b1);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"                  // borrow ? 0xffffffff : 0
        "sub.cc.u32 %0, %0, m;\n\t"              // - (2^32 - 1) on borrow
        "madc.lo.u32 %1, %6, 1, %1;\n\t"         // hi + 0xffffffff + (no second borrow) = hi - second borrow
        "}" : "=&r"(o0), "=&r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(0xffffffffu));
    return gl_pack(o0, o1);
}
// p - b for canonical b (result in (0, p])
__device__ __forceinline__ u64 f_negp(u64 b) {
    u32 b0, b1, n0, n1;
    gl_unpack(b, b0, b1);
    asm("{\n\t"
        "sub.cc.u32 %0, 1, %2;\n\t"
        "subc.u32 %1, 0xffffffff, %3;\n\t"
        "}" : "=&r"(n0), "=&r"(n1) : "r"(b0), "r"(b1));
    return gl_pack(n0, n1);
}
__device__ __forceinline__ u64 f_add_v2(u64 a, u64 b) { return f_sub_v2(a, f_negp(b)); }

// product + reduction with the carry-in-only steps as madc / IMAD, canonical out
__device__ __forceinline__ u64 f_mul_v2(u64 a, u64 b) {
    u32 a0, a1, b0, b1;
    gl_unpack(a, a0, a1); gl_unpack(b, b0, b1);
    u32 r0, p00h, p01l, p01h, p10l, p10h, p11l, p11h;
    gl_wide(a0, b0, r0, p00h); gl_wide(a0, b1, p01l, p01h); gl_wide(a1, b0, p10l, p10h); gl_wide(a1, b1, p11l, p11h);
    u32 r1, r2, r3;
    const u32 zero = 0, ones = 0xffffffffu;
    asm("{\n\t"
        "add.cc.u32 %0, %3, %4;\n\t"
        "addc.cc.u32 %1, %5, %7;\n\t"
        "madc.lo.u32 %2, %9, 1, %10;\n\t"       // r3 = p11.hi + c
        "add.cc.u32 %0, %0, %6;\n\t"
        "addc.cc.u32 %1, %1, %8;\n\t"
        "madc.lo.u32 %2, %2, 1, %10;\n\t"
        "}" : "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(p00h), "r"(p01l), "r"(p01h), "r"(p10l), "r"(p10h), "r"(p11l), "r"(p11h), "r"(zero));
    u64 A;
    asm("mad.wide.u32 %0, %1, 0xffffffff, %2;" : "=l"(A) : "r"(r2), "l"(gl_pack(r0, 0)));
    u32 A0, A1, o0, o1;
    gl_unpack(A, A0, A1);
    asm("{\n\t.reg .u32 mb, mc;\n\t"
        "add.cc.u32 %1, %3, %4;\n\t"              // hi + r1 -> carry
        "madc.lo.u32 mc, %6, 0, %6;\n\t"          // mc = carry (0 / 1)
        "sub.cc.u32 %0, %2, %5;\n\t"              // lo - r3 -> borrow
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 mb, 0, 0;\n\t"                  // borrow ? 0xffffffff : 0
        "mul.lo.u32 mc, mc, 0xffffffff;\n\t"      // carry  ? 0xffffffff : 0   (IMAD)
        "add.cc.u32 %0, %0, mc;\n\t"              // + EPS on carry
        "madc.lo.u32 %1, %1, 1, %6;\n\t"
        "sub.cc.u32 %0, %0, mb;\n\t"              // - EPS on borrow
        "madc.lo.u32 %1, %7, 1, %1;\n\t"          // hi + 0xffffffff + (no borrow)
        "}" : "=&r"(o0), "=&r"(o1) : "r"(A0), "r"(A1), "r"(r1), "r"(r3), "r"(zero), "r"(ones));
    u64 r = gl_pack(o0, o1);
    return f_canon(r);
}

// every variant against the reference forms on edge values and a pseudo-random stream; *bad counts mismatches
__global__ void check_kernel(unsigned long long* bad) {
    const u64 edge[12] = {0, 1, 2, GL_P - 1, GL_P - 2, 0xffffffffull, 0x100000000ull, 0x100000001ull, 0x8000000000000000ull,
                          GL_P - 0x100000000ull, 0xfffffffe00000002ull, 0x00000001ffffffffull};
    const u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 z = 0x9e3779b97f4a7c15ull * (t + 1);
    for (int i = 0; i < 256; ++i) {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
        u64 a = (i < 144) ? edge[i % 12] : gl_canon(z), b = (i < 144) ? edge[(i / 12) % 12] : gl_canon(z * 0x2545f4914f6cdd1dull + t);
        if (f_sub_v2(a, b) != f_sub(a, b) || f_add_v2(a, b) != f_add(a, b) || f_mul_v2(a, b) != f_mul(a, b)) atomicAdd(bad, 1ull);
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
            if (V == 0) { u64 s = f_add(a[k], b[k]), d = f_sub(a[k], b[k]); a[k] = s; b[k] = d; }                       // butterflies only, current
            if (V == 1) { u64 s = f_add_v2(a[k], b[k]), d = f_sub_v2(a[k], b[k]); a[k] = s; b[k] = d; }                 // butterflies only, rebalanced
            if (V == 2) { u64 s = f_add(a[k], b[k]), d = f_sub(a[k], b[k]); a[k] = s; b[k] = f_mul(d, w[k]); }          // + one multiply, current
            if (V == 3) { u64 s = f_add_v2(a[k], b[k]), d = f_sub_v2(a[k], b[k]); a[k] = s; b[k] = f_mul(d, w[k]); }    // rebalanced add/sub, current mul
            if (V == 4) { u64 s = f_add_v2(a[k], b[k]), d = f_sub_v2(a[k], b[k]); a[k] = s; b[k] = f_mul_v2(d, w[k]); } // all rebalanced
            if (V == 5) { a[k] = f_mul(a[k], w[k]); }                                                                  // multiplies only, current
            if (V == 6) { a[k] = f_mul_v2(a[k], w[k]); }                                                               // multiplies only, rebalanced
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
        std::printf("IMAD.X forms vs reference forms: %llu mismatches in %d cases\n", h, 64 * 128 * 256);
    }
    u64 c[7];
    run<0>("butterfly (add + sub), current", out, &c[0]);
    run<1>("butterfly (add + sub), IMAD.X forms", out, &c[1]);
    run<2>("butterfly + multiply, current", out, &c[2]);
    run<3>("butterfly IMAD.X forms + current multiply", out, &c[3]);
    run<4>("butterfly + multiply, IMAD.X forms", out, &c[4]);
    run<5>("multiply only, current", out, &c[5]);
    run<6>("multiply only, IMAD.X forms", out, &c[6]);
    std::printf("results agree: %s\n", (c[0] == c[1] && c[2] == c[3] && c[3] == c[4] && c[5] == c[6]) ? "yes" : "NO");
    return 0;
}
