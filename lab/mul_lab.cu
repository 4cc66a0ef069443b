// SASS lab: count instructions for Goldilocks mul variants (compile only)
#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
#define D __device__ __forceinline__

// variant A: current C version
D u64 mulA(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 p00 = (u64)a0 * b0;
    u64 t = (u64)a0 * b1 + (p00 >> 32);
    u64 t2 = (u64)a1 * b0 + (u32)t;
    u64 hi = (u64)a1 * b1 + (t >> 32) + (t2 >> 32);
    u64 lo = (t2 << 32) | (u32)p00;
    u64 hi_hi = hi >> 32, hi_lo = hi & 0xffffffffu;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= 0xffffffffu;
    u64 t1 = (hi_lo << 32) - hi_lo;
    u64 r = t0 + t1;
    if (r < t1) r += 0xffffffffu;
    return r;
}

// variant B: PTX carry chains. product limbs (r0,r1,r2,r3); result = (r0,r1) + r2*EPS - r3
D u64 mulB(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1, r2, r3;
    asm("{\n\t"
        ".reg .u64 p, t, t2, h;\n\t"
        ".reg .u32 plo, phi, tlo, thi, t2lo, t2hi, z;\n\t"
        "mov.u32 z, 0;\n\t"
        "mul.wide.u32 p, %4, %6;\n\t"
        "mov.b64 {plo, phi}, p;\n\t"
        "mov.b64 t, {phi, z};\n\t"
        "mad.wide.u32 t, %4, %7, t;\n\t"
        "mov.b64 {tlo, thi}, t;\n\t"
        "mov.b64 t2, {tlo, z};\n\t"
        "mad.wide.u32 t2, %5, %6, t2;\n\t"
        "mov.b64 {t2lo, t2hi}, t2;\n\t"
        "mov.b64 h, {thi, z};\n\t"
        "mad.wide.u32 h, %5, %7, h;\n\t"
        "mov.b64 t, {t2hi, z};\n\t"
        "add.u64 h, h, t;\n\t"
        "mov.b64 {%2, %3}, h;\n\t"
        "mov.u32 %0, plo;\n\t"
        "mov.u32 %1, t2lo;\n\t"
        "}" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    // reduce: x = (r0,r1) - r3 + r2*2^32 - r2
    u32 o0, o1;
    asm("{\n\t"
        ".reg .u32 m, c;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"      // lo - hi_hi
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"           // m = borrow ? 0xffffffff : 0  -> subtract EPS = add 1 to... (x - EPS)
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        // add r2*EPS = (r2<<32) - r2 : first subtract r2 from low (borrow into high), then add r2 to high
        "sub.cc.u32 %0, %0, %4;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"           // borrow b: value went below 0 -> we are at +2^64, must subtract EPS later
        "add.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 c, 0, 0;\n\t"           // carry c: must add EPS
        // net: + EPS*c - EPS*b where b = (m != 0). c and b: if both, cancel.
        "add.u32 c, c, m;\n\t"            // c + m : 1 + (-1) = 0 ; 1 ; -1 ; 0  (as signed 32-bit k)
        // add k*EPS where k in {-1,0,1}: low += -k ; (k*EPS = k<<32 - k)
        "sub.cc.u32 %0, %0, c;\n\t"       // low -= k  (k=-1: low += 1 ... ) careful with sign: handled as 64-bit below
        "subc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(o0), "=&r"(o1) : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return ((u64)o1 << 32) | o0;
}

__global__ void kA(u64* x) { u64 a = x[threadIdx.x], b = x[threadIdx.x + 32]; for (int i = 0; i < 4; ++i) a = mulA(a, b); x[threadIdx.x] = a; }
__global__ void kB(u64* x) { u64 a = x[threadIdx.x], b = x[threadIdx.x + 32]; for (int i = 0; i < 4; ++i) a = mulB(a, b); x[threadIdx.x] = a; }
