// Goldilocks field (p = 2^64 - 2^32 + 1) and its quadratic extension F_p[X]/(X^2 - 7) for sm_100a.
// Replaces qp-plonky2-field 1.1.1 `GoldilocksField` / `QuadraticExtension` (type contract:
// /root/reference/common/src/circuit.rs:10-12) on the device. 64-bit products are built from four
// 32x32->64 IMAD.WIDE.U32 with 64-bit addends (no carry chains needed, see mul128), the reduction uses
// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).
//
// Representation: every value handed across a kernel boundary is canonical (< p). Inside kernels the
// "lazy" functions keep values in [0, 2^64) and canonicalise once at the end.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

namespace zkb {

typedef uint64_t u64;
typedef uint32_t u32;

constexpr u64 GL_P = 0xFFFFFFFF00000001ULL;
constexpr u64 GL_EPS = 0xFFFFFFFFULL;
constexpr u64 GL_GEN = 0xc65c18b67785d900ULL;        // multiplicative generator = coset shift
constexpr u64 GL_TWO_ADIC_ROOT = 0x64fdd1a46201e246ULL;  // order 2^32
constexpr u64 GL_W = 7;                               // X^2 = 7

ZKB_HD u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

// a, b any u64 (value mod p); result any u64 with the same residue as a + b
ZKB_HD u64 gl_add_lazy(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a) {           // wrapped: add 2^64 mod p
        s += GL_EPS;
        if (s < GL_EPS) s += GL_EPS;
    }
    return s;
}
ZKB_HD u64 gl_sub_lazy(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) {           // borrowed: subtract 2^64 mod p
        u64 e = d - GL_EPS;
        if (d < GL_EPS) e -= GL_EPS;
        d = e;
    }
    return d;
}
// canonical in, canonical out
ZKB_HD u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
ZKB_HD u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a - b + GL_P; }
ZKB_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

ZKB_HD void mul128(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 p00 = (u64)a0 * b0;
    u64 t = (u64)a0 * b1 + (p00 >> 32);      // cannot overflow
    u64 t2 = (u64)a1 * b0 + (u32)t;          // cannot overflow
    hi = (u64)a1 * b1 + (t >> 32) + (t2 >> 32);
    lo = (t2 << 32) | (u32)p00;
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    lo = (u64)x;
    hi = (u64)(x >> 64);
#endif
}

#if defined(__CUDACC__)
// 64-bit <-> 2 x 32-bit through mov.b64: a C-level (u32)(v >> 32) makes ptxas emit a stray "VIADD hi, hi, 0" per use
__device__ __forceinline__ void gl_unpack(u64 v, u32& lo, u32& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ u64 gl_pack(u32 lo, u32 hi) { u64 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi)); return v; }
__device__ __forceinline__ void gl_wide(u32 a, u32 b, u32& lo, u32& hi) {
    u64 v;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(v) : "r"(a), "r"(b));
    gl_unpack(v, lo, hi);
}
#endif
#if defined(__CUDA_ARCH__)
// ---- device fast paths: 32-bit limbs, carry flags through add.cc/addc (IADD3/IADD3.X/IMAD.X in SASS), no
// compare+select sequences. The limb algorithms were checked exhaustively on edge values on the host. ----

// (a1:a0) * (b1:b0) -> 128-bit product limbs r0..r3: 4 x IMAD.WIDE.U32 + two 3-limb carry chains
__device__ __forceinline__ void gl_mul128_limbs(u32 a0, u32 a1, u32 b0, u32 b1, u32& r0, u32& r1, u32& r2, u32& r3) {
    u32 p00h, p01l, p01h, p10l, p10h, p11l, p11h;
    gl_wide(a0, b0, r0, p00h);
    gl_wide(a0, b1, p01l, p01h);
    gl_wide(a1, b0, p10l, p10h);
    gl_wide(a1, b1, p11l, p11h);
    asm("{\n\t"
        "add.cc.u32 %0, %3, %4;\n\t"        // r1 = p00.hi + p01.lo
        "addc.cc.u32 %1, %5, %7;\n\t"       // r2 = p01.hi + p10.hi + c
        "addc.u32 %2, %9, 0;\n\t"           // r3 = p11.hi + c
        "add.cc.u32 %0, %0, %6;\n\t"        // r1 += p10.lo
        "addc.cc.u32 %1, %1, %8;\n\t"       // r2 += p11.lo + c
        "addc.u32 %2, %2, 0;\n\t"
        "}" : "=&r"(r1), "=&r"(r2), "=&r"(r3)
            : "r"(p00h), "r"(p01l), "r"(p01h), "r"(p10l), "r"(p10h), "r"(p11l), "r"(p11h));
}
// T = (r1:r0) + r2*(2^32-1) - r3 lies in (-2^32, 2^65): result = (T mod 2^64) + EPS*[T >= 2^64] - EPS*[T < 0],
// which needs exactly one correction and cannot wrap again. r2*EPS + r0 is one IMAD.WIDE
// (cannot overflow), so the fold costs 1 FMA-pipe + 10 ALU-pipe instructions.
__device__ __forceinline__ u64 gl_reduce_limbs(u32 r0, u32 r1, u32 r2, u32 r3) {
    u64 A;
    asm("mad.wide.u32 %0, %1, 0xffffffff, %2;" : "=l"(A) : "r"(r2), "l"(gl_pack(r0, 0)));
    u32 A0, A1, o0, o1;
    gl_unpack(A, A0, A1);
    asm("{\n\t"
        ".reg .u32 mb, mc;\n\t"
        "add.cc.u32 %1, %3, %4;\n\t"        // high limb + r1 -> carry
        "addc.u32 mc, 0, 0;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"        // - r3 -> borrow
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 mb, 0, 0;\n\t"            // borrow ? 0xffffffff : 0
        "neg.s32 mc, mc;\n\t"               // carry  ? 0xffffffff : 0
        "add.cc.u32 %0, %0, mc;\n\t"        // + EPS on carry
        "addc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 %0, %0, mb;\n\t"        // - EPS on borrow
        "subc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(o0), "=&r"(o1) : "r"(A0), "r"(A1), "r"(r1), "r"(r3));
    return gl_pack(o0, o1);
}
#endif

// (hi:lo) mod p into [0, 2^64) — not necessarily canonical
ZKB_HD u64 gl_reduce128_lazy(u64 lo, u64 hi) {
#if defined(__CUDA_ARCH__)
    return gl_reduce_limbs((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
#else
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    u64 t1 = (hi_lo << 32) - hi_lo;           // hi_lo * (2^32 - 1)
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return r;
#endif
}
// inputs: any u64; output lazy
ZKB_HD u64 gl_mul_lazy(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 r0, r1, r2, r3;
    gl_mul128_limbs((u32)a, (u32)(a >> 32), (u32)b, (u32)(b >> 32), r0, r1, r2, r3);
    return gl_reduce_limbs(r0, r1, r2, r3);
#else
    u64 lo, hi;
    mul128(a, b, lo, hi);
    return gl_reduce128_lazy(lo, hi);
#endif
}
// squaring: 3 wide multiplies (the cross product once, doubled with shifts) instead of 4
ZKB_HD u64 gl_sqr_lazy(u64 a) {
#if defined(__CUDA_ARCH__)
    u32 a0, a1, p00l, p00h, m0, m1, p11l, p11h;
    gl_unpack(a, a0, a1);
    gl_wide(a0, a0, p00l, p00h);
    gl_wide(a0, a1, m0, m1);
    gl_wide(a1, a1, p11l, p11h);
    const u32 d0 = m0 << 1, d1 = __funnelshift_l(m0, m1, 1), d2 = m1 >> 31;      // 2m as three limbs
    u32 r1, r2, r3;
    asm("{\n\t"
        "add.cc.u32 %0, %3, %4;\n\t"        // r1 = p00.hi + d0
        "addc.cc.u32 %1, %5, %6;\n\t"       // r2 = p11.lo + d1 + c
        "addc.u32 %2, %7, %8;\n\t"          // r3 = p11.hi + d2 + c   (cannot overflow: a^2 < 2^128)
        "}" : "=&r"(r1), "=&r"(r2), "=&r"(r3)
            : "r"(p00h), "r"(d0), "r"(p11l), "r"(d1), "r"(p11h), "r"(d2));
    return gl_reduce_limbs(p00l, r1, r2, r3);
#else
    return gl_mul_lazy(a, a);
#endif
}
// a any u64, c <= p - 1 (canonical): a + c cannot wrap twice
ZKB_HD u64 gl_add_lazy_c(u64 a, u64 c) {
#if defined(__CUDA_ARCH__)
    u32 o0, o1;
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "neg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t}" : "=&r"(o0), "=&r"(o1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)c), "r"((u32)(c >> 32)));
    return ((u64)o1 << 32) | o0;
#else
    u64 s = a + c;
    if (s < a) s += GL_EPS;
    return s;
#endif
}
// a any u64, c <= p - 1 (canonical): a - c cannot borrow twice
ZKB_HD u64 gl_sub_lazy_c(u64 a, u64 c) {
#if defined(__CUDA_ARCH__)
    u32 o0, o1;
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t}" : "=&r"(o0), "=&r"(o1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)c), "r"((u32)(c >> 32)));
    return ((u64)o1 << 32) | o0;
#else
    u64 d = a - c;
    if (a < c) d -= GL_EPS;
    return d;
#endif
}
ZKB_HD u64 gl_mul(u64 a, u64 b) { return gl_canon(gl_mul_lazy(a, b)); }
ZKB_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }

ZKB_HD u64 gl_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, a);
        a = gl_mul(a, a);
        e >>= 1;
    }
    return r;
}
ZKB_HD u64 gl_sqr_n_lazy(u64 a, int k) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < k; ++i) a = gl_sqr_lazy(a);
    return a;
}
// a^(p-2) with p - 2 = 2^64 - 2^32 - 1, by the chain e(k) = a^(2^k - 1): 74 squarings + 10 multiplies, lazy inside
// (a plain square-and-multiply needs 64 + 63). gl_inv(0) = 0.
ZKB_HD u64 gl_inv(u64 a) {
    const u64 e2 = gl_mul_lazy(gl_sqr_lazy(a), a);
    const u64 e3 = gl_mul_lazy(gl_sqr_lazy(e2), a);
    const u64 e4 = gl_mul_lazy(gl_sqr_n_lazy(e2, 2), e2);
    const u64 e7 = gl_mul_lazy(gl_sqr_n_lazy(e4, 3), e3);
    const u64 e8 = gl_mul_lazy(gl_sqr_n_lazy(e4, 4), e4);
    const u64 e15 = gl_mul_lazy(gl_sqr_n_lazy(e8, 7), e7);
    const u64 e16 = gl_mul_lazy(gl_sqr_n_lazy(e8, 8), e8);
    const u64 e31 = gl_mul_lazy(gl_sqr_n_lazy(e16, 15), e15);
    const u64 b = gl_sqr_lazy(e31);                       // a^(2^32 - 2)
    const u64 c = gl_mul_lazy(b, a);                      // a^(2^32 - 1)
    return gl_canon(gl_mul_lazy(gl_sqr_n_lazy(b, 32), c));
}
inline u64 gl_root_of_unity(unsigned k) {
    u64 r = GL_TWO_ADIC_ROOT;
    for (unsigned i = k; i < 32; ++i) r = gl_mul(r, r);
    return r;
}

#if defined(__CUDACC__)
// ---- device-only canonical arithmetic on carry flags (5 / 7 / ~25 instructions; the portable gl_add / gl_sub / gl_mul
// compile to compare + select sequences of 9-12 on top) ----
// Pipe balance (lab/pipe_balance.cu, measured on B200): alu-pipe instructions (IADD3, IADD3.X, SEL, ISETP, LOP3) and 32-bit
// IMADs issue on alternate cycles, so code made of carry chains is bound by its alu count while the FMA pipe idles. Every step
// of these forms that only CONSUMES a carry is therefore written as madc.lo (IMAD.X on the FMA pipe): butterflies 1.22x,
// butterfly + multiply 1.16x against the all-IADD3.X forms. They rely on the carry flag's polarity after sub.cc / subc.cc as
// seen by a following madc: it is the hardware carry of a + ~b + 1, i.e. 1 = NO borrow (subc compensates by itself, madc does
// not). lab/pipe_balance.cu checks the forms against the plain ones on 2 M cases; the parity tests would catch a toolchain that
// changed it.
// canonical in, canonical out: 3 alu + 2 fma
ZKB_D u64 f_sub(u64 a, u64 b) {
    u32 a0, a1, b0, b1, o0, o1;
    gl_unpack(a, a0, a1);
    gl_unpack(b, b0, b1);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"              // borrow ? 0xffffffff : 0
        "sub.cc.u32 %0, %0, m;\n\t"          // + p  ==  - (2^32 - 1)  (mod 2^64)
        "madc.lo.u32 %1, %6, 1, %1;\n\t"     // hi + 0xffffffff + (no second borrow)  ==  hi - second borrow
        "}" : "=&r"(o0), "=&r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(0xffffffffu));
    return gl_pack(o0, o1);
}
ZKB_D u64 f_add(u64 a, u64 b) {               // a - (p - b); p - b in (0, p], and a < p, so one borrow fix is exact
    u32 b0, b1, n0, n1;
    gl_unpack(b, b0, b1);
    asm("{\n\t"
        "sub.cc.u32 %0, 1, %2;\n\t"
        "subc.u32 %1, 0xffffffff, %3;\n\t"
        "}" : "=&r"(n0), "=&r"(n1) : "r"(b0), "r"(b1));
    return f_sub(a, gl_pack(n0, n1));
}

// r >= p  <=>  r + (2^32 - 1) carries out of 64 bits (high word all ones and low word >= 1); then r - p = (low - 1, 0).
// With c = that carry (0 / 1): low' = low + c * 0xffffffff, high' = high + c (0xffffffff + 1 wraps to 0) — the two carry adds on
// the alu pipe, everything else on the FMA pipe (the compare + select form is 4-5 alu instructions).
ZKB_D u64 f_canon(u64 r) {
    u32 lo, hi, c;
    gl_unpack(r, lo, hi);
    asm("{\n\t.reg .u32 t;\n\t"
        "add.cc.u32 t, %1, 0xffffffff;\n\t"
        "addc.cc.u32 t, %2, 0;\n\t"
        "madc.lo.u32 %0, %3, 0, %3;\n\t"
        "}" : "=r"(c) : "r"(lo), "r"(hi), "r"(0u));
    u32 lo2, hi2;
    asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(lo2) : "r"(c), "r"(lo));
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(hi2) : "r"(c), "r"(hi));
    return gl_pack(lo2, hi2);
}

// canonical product of any two u64: the same product limbs and fold as gl_mul_lazy / gl_reduce_limbs (which the Poseidon code,
// FMA-pipe bound, keeps), with the carry-consuming steps on the FMA pipe: 5 alu + 5 fma in the fold instead of 9 + 2
ZKB_D u64 f_mul(u64 a, u64 b) {
    u32 a0, a1, b0, b1, r0, p00h, p01l, p01h, p10l, p10h, p11l, p11h, r1, r2, r3;
    gl_unpack(a, a0, a1);
    gl_unpack(b, b0, b1);
    gl_wide(a0, b0, r0, p00h);
    gl_wide(a0, b1, p01l, p01h);
    gl_wide(a1, b0, p10l, p10h);
    gl_wide(a1, b1, p11l, p11h);
    const u32 zero = 0, ones = 0xffffffffu;
    asm("{\n\t"
        "add.cc.u32 %0, %3, %4;\n\t"        // r1 = p00.hi + p01.lo
        "addc.cc.u32 %1, %5, %7;\n\t"       // r2 = p01.hi + p10.hi + c
        "madc.lo.u32 %2, %9, 1, %10;\n\t"   // r3 = p11.hi + c
        "add.cc.u32 %0, %0, %6;\n\t"        // r1 += p10.lo
        "addc.cc.u32 %1, %1, %8;\n\t"       // r2 += p11.lo + c
        "madc.lo.u32 %2, %2, 1, %10;\n\t"
        "}" : "=&r"(r1), "=&r"(r2), "=&r"(r3)
            : "r"(p00h), "r"(p01l), "r"(p01h), "r"(p10l), "r"(p10h), "r"(p11l), "r"(p11h), "r"(zero));
    u64 A;
    asm("mad.wide.u32 %0, %1, 0xffffffff, %2;" : "=l"(A) : "r"(r2), "l"(gl_pack(r0, 0)));
    u32 A0, A1, o0, o1;
    gl_unpack(A, A0, A1);
    asm("{\n\t.reg .u32 mb, mc;\n\t"
        "add.cc.u32 %1, %3, %4;\n\t"        // high limb + r1 -> carry
        "madc.lo.u32 mc, %6, 0, %6;\n\t"    // carry (0 / 1)
        "sub.cc.u32 %0, %2, %5;\n\t"        // - r3 -> borrow
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 mb, 0, 0;\n\t"            // borrow ? 0xffffffff : 0
        "mul.lo.u32 mc, mc, 0xffffffff;\n\t" // carry  ? 0xffffffff : 0
        "add.cc.u32 %0, %0, mc;\n\t"        // + EPS on carry
        "madc.lo.u32 %1, %1, 1, %6;\n\t"
        "sub.cc.u32 %0, %0, mb;\n\t"        // - EPS on borrow
        "madc.lo.u32 %1, %7, 1, %1;\n\t"    // hi + 0xffffffff + (no borrow)
        "}" : "=&r"(o0), "=&r"(o1) : "r"(A0), "r"(A1), "r"(r1), "r"(r3), "r"(zero), "r"(ones));
    return f_canon(gl_pack(o0, o1));
}


// x * 2^K (mod p) for a compile-time K in (0, 96), canonical in / canonical out, WITHOUT a 64x64 multiply: the radix-8 / radix-16
// butterflies' internal twiddles are powers of two (w_16 = 2^12, w_8 = 2^24, w_4 = 2^48: SURVEY.md A.1). x << (K mod 32) is three
// limbs (y2:y1:y0); the word offset K / 32 places them at 2^0, 2^32 or 2^64, and 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 fold
// them back with one f_add / f_sub on values that are canonical by construction:
//   K < 32 : canon(y1:y0) + y2 (2^32 - 1)                10-17 instructions instead of the ~30 issue slots of f_mul
//   K < 64 : (y0:0) + y1 (2^32 - 1) - y2
//   K < 96 : y0 (2^32 - 1) - (y2:y1)
// (multiplication by 2^(96 + K) is the same applied to -x: the caller swaps the operands of the butterfly's subtraction)
template <int K>
ZKB_D u64 f_shl(u64 x) {
    static_assert(K > 0 && K < 96, "shift out of range");
    constexpr int r = K & 31, w = K >> 5;
    u32 x0, x1;
    gl_unpack(x, x0, x1);
    const u32 y0 = x0 << r, y1 = r ? __funnelshift_l(x0, x1, r) : x1, y2 = r ? (x1 >> (32 - r)) : 0u;
    auto times_eps = [](u32 v) { return gl_pack(0u - v, v - (v != 0u)); };      // v (2^32 - 1) = (v << 32) - v  < p
    if (w == 0) return f_add(f_canon(gl_pack(y0, y1)), times_eps(y2));
    if (w == 1) return f_sub(f_add(gl_pack(0u, y0), times_eps(y1)), (u64)y2);
    return f_sub(times_eps(y0), gl_pack(y1, y2));                                // (y2:y1) < 2^63: canonical
}
// d * w_16^(+-IDX): the butterfly's twiddle for d = a - b, with the sign of the inverse twiddle (w_16^-i = -2^(96 - 12 i)) taken
// by the subtraction itself
template <int IDX, bool INV>
ZKB_D u64 f_sub_twiddle16(u64 a, u64 b) {
    static_assert(IDX > 0 && IDX < 8, "w_16 exponent");
    if (!INV) return f_shl<12 * IDX>(f_sub(a, b));
    return f_shl<96 - 12 * IDX>(f_sub(b, a));
}
// the same for w_64 = 8 (w_64^IDX = 2^(3 IDX), IDX < 32): every twiddle inside a transform of up to 64 points is a shift
template <int IDX, bool INV>
ZKB_D u64 f_sub_twiddle64(u64 a, u64 b) {
    static_assert(IDX > 0 && IDX < 32, "w_64 exponent");
    if (!INV) return f_shl<3 * IDX>(f_sub(a, b));
    return f_shl<96 - 3 * IDX>(f_sub(b, a));
}
// x * w_64^E for any compile-time E (mod 64): 2^(3 E mod 192), 2^96 = -1
template <int E>
ZKB_D u64 f_mul_w64(u64 x) {
    constexpr int S = 3 * (((E % 64) + 64) % 64);
    if constexpr (S == 0) return x;
    else if constexpr (S < 96) return f_shl<S>(x);
    else if constexpr (S == 96) return f_sub(0, x);
    else return f_sub(0, f_shl<S - 96>(x));
}
#endif

// ---- quadratic extension, canonical components ----
struct ext2 {
    u64 a, b;
};
ZKB_HD ext2 e_make(u64 a, u64 b) { ext2 r; r.a = a; r.b = b; return r; }
ZKB_HD ext2 e_from(u64 a) { return e_make(a, 0); }
ZKB_HD ext2 e_add(ext2 x, ext2 y) { return e_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
ZKB_HD ext2 e_sub(ext2 x, ext2 y) { return e_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
ZKB_HD ext2 e_mul(ext2 x, ext2 y) {
    u64 aa = gl_mul(x.a, y.a), bb = gl_mul(x.b, y.b);
    u64 ab = gl_mul(x.a, y.b), ba = gl_mul(x.b, y.a);
    return e_make(gl_add(aa, gl_mul(bb, GL_W)), gl_add(ab, ba));
}
ZKB_HD ext2 e_mul_base(ext2 x, u64 s) { return e_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
ZKB_HD ext2 e_inv(ext2 x) {
    u64 norm = gl_sub(gl_sqr(x.a), gl_mul(GL_W, gl_sqr(x.b)));
    u64 ni = gl_inv(norm);
    return e_make(gl_mul(x.a, ni), gl_mul(gl_neg(x.b), ni));
}
ZKB_HD bool e_eq(ext2 x, ext2 y) { return x.a == y.a && x.b == y.b; }
ZKB_HD ext2 e_pow2k(ext2 x, unsigned k) {
    for (unsigned i = 0; i < k; ++i) x = e_mul(x, x);
    return x;
}
ZKB_HD ext2 e_pow(ext2 x, u64 e) {
    ext2 r = e_from(1);
    while (e) {
        if (e & 1) r = e_mul(r, x);
        x = e_mul(x, x);
        e >>= 1;
    }
    return r;
}

ZKB_HD u32 bitrev32(u32 x, unsigned bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    u32 r = 0;
    for (unsigned i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace zkb
