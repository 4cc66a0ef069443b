// SASS lab for the Poseidon permutation: 32-bit limb representation, carry-chain PTX.
#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
#define D __device__ __forceinline__
__constant__ u64 c_rc[360];

struct gl2 { u32 lo, hi; };

// (a1:a0) * (b1:b0) -> r0..r3
D void mul128(u32 a0, u32 a1, u32 b0, u32 b1, u32& r0, u32& r1, u32& r2, u32& r3) {
    u64 p00 = (u64)a0 * b0, p01 = (u64)a0 * b1, p10 = (u64)a1 * b0, p11 = (u64)a1 * b1;
    u32 c0, c1;
    asm("{\n\t"
        "add.cc.u32 %1, %6, %8;\n\t"        // r1 = p00.hi + p01.lo
        "addc.cc.u32 %2, %9, %11;\n\t"      // r2 = p01.hi + p10.hi + c
        "addc.u32 %3, %13, 0;\n\t"          // r3 = p11.hi + c
        "add.cc.u32 %1, %1, %10;\n\t"       // r1 += p10.lo
        "addc.cc.u32 %2, %2, %12;\n\t"      // r2 += p11.lo + c
        "addc.u32 %3, %3, 0;\n\t"
        "}" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(c0), "=r"(c1)
            : "r"((u32)(p00 >> 32)), "r"((u32)p00), "r"((u32)p01), "r"((u32)(p01 >> 32)), "r"((u32)p10), "r"((u32)(p10 >> 32)),
              "r"((u32)p11), "r"((u32)(p11 >> 32)));
    r0 = (u32)p00;
}

// (r1:r0) + r2*EPS - r3  -> lazy 64-bit (exactly one correction suffices)
D gl2 reduce(u32 r0, u32 r1, u32 r2, u32 r3) {
    gl2 o;
    asm("{\n\t"
        ".reg .u32 mb, mc, e0, e1;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"        // U = lo - r3
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 mb, 0, 0;\n\t"            // mb = borrow ? 0xffffffff : 0
        "sub.cc.u32 e0, 0, %4;\n\t"         // E = (r2 << 32) - r2
        "subc.u32 e1, %4, 0;\n\t"
        "add.cc.u32 %0, %0, e0;\n\t"        // W = U + E
        "addc.cc.u32 %1, %1, e1;\n\t"
        "addc.u32 mc, 0, 0;\n\t"            // mc = carry (0/1)
        "neg.s32 mc, mc;\n\t"               // 0 / 0xffffffff
        // W + mc(as u64 zero-ext = EPS if carry) - mb(zero-ext = EPS if borrow)
        "add.cc.u32 %0, %0, mc;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 %0, %0, mb;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(o.lo), "=&r"(o.hi) : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return o;
}
D gl2 mul(gl2 a, gl2 b) { u32 r0, r1, r2, r3; mul128(a.lo, a.hi, b.lo, b.hi, r0, r1, r2, r3); return reduce(r0, r1, r2, r3); }
D gl2 sbox7(gl2 x) { gl2 x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2); return mul(x3, x4); }

// lazy add of a canonical constant (no double wrap possible)
D gl2 add_rc(gl2 a, u64 c) {
    gl2 o;
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "neg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t}" : "=&r"(o.lo), "=&r"(o.hi) : "r"(a.lo), "r"(a.hi), "r"((u32)c), "r"((u32)(c >> 32)));
    return o;
}

#define MC(i) ((i) == 0 ? 17u : (i) == 1 ? 15u : (i) == 2 ? 41u : (i) == 3 ? 16u : (i) == 4 ? 2u : (i) == 5 ? 28u : \
               (i) == 6 ? 13u : (i) == 7 ? 13u : (i) == 8 ? 39u : (i) == 9 ? 18u : (i) == 10 ? 34u : 20u)

// out = MDS * s + rc (rc canonical or zero) ; accumulators initialised with the constant's halves
template <bool WITH_RC>
D void mds(gl2* s, const u64* rc) {
    gl2 o[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        u64 al = 0, ah = 0;
        if (WITH_RC) { u64 c = rc[r]; al = (u32)c; ah = c >> 32; }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            al += (u64)s[(i + r) % 12].lo * MC(i);
            ah += (u64)s[(i + r) % 12].hi * MC(i);
        }
        if (r == 0) { al += (u64)s[0].lo * 8u; ah += (u64)s[0].hi * 8u; }
        // value = al + ah*2^32, al, ah < 2^43.  limbs: al=(a0,a1) ah=(b0,b1): v = a0 + (a1+b0)*2^32 + (b1 + c)*2^64
        u32 a0 = (u32)al, a1 = (u32)(al >> 32), b0 = (u32)ah, b1 = (u32)(ah >> 32);
        asm("{\n\t.reg .u32 k, m;\n\t"
            "add.cc.u32 %1, %3, %4;\n\t"      // m1 = a1 + b0
            "addc.u32 k, %5, 0;\n\t"          // k = b1 + carry  (< 2^12)
            // v = (a0, m1) + k*EPS = (a0 - k, m1 + k) with borrow/carry
            "sub.cc.u32 %0, %2, k;\n\t"
            "subc.cc.u32 %1, %1, 0;\n\t"
            "subc.u32 m, 0, 0;\n\t"           // borrow mask (only if m1 == 0 and a0 < k)
            "add.cc.u32 %1, %1, k;\n\t"
            "addc.u32 k, 0, 0;\n\t"           // carry
            "neg.s32 k, k;\n\t"
            "add.cc.u32 %0, %0, k;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "sub.cc.u32 %0, %0, m;\n\t"
            "subc.u32 %1, %1, 0;\n\t}"
            : "=&r"(o[r].lo), "=&r"(o[r].hi) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    }
#pragma unroll
    for (int r = 0; r < 12; ++r) s[r] = o[r];
}

D void permute(gl2* s) {
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = add_rc(s[i], c_rc[i]);
    int rc = 12;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < 12; ++i) s[i] = sbox7(s[i]);
        mds<true>(s, c_rc + rc);
        rc += 12;
    }
#pragma unroll 1
    for (int r = 0; r < 22; ++r) {
        s[0] = sbox7(s[0]);
        mds<true>(s, c_rc + rc);
        rc += 12;
    }
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int i = 0; i < 12; ++i) s[i] = sbox7(s[i]);
        mds<true>(s, c_rc + rc);
        rc += 12;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = sbox7(s[i]);
    mds<false>(s, nullptr);
}

__global__ void __launch_bounds__(128) kperm(u64* states, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    gl2 s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { u64 v = states[i * 12 + k]; s[k].lo = (u32)v; s[k].hi = (u32)(v >> 32); }
    permute(s);
#pragma unroll
    for (int k = 0; k < 12; ++k) states[i * 12 + k] = ((u64)s[k].hi << 32) | s[k].lo;
}
