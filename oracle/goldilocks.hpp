// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// CPU restatement of the Goldilocks field and its quadratic extension as used by the reference's
// proving dependency qp-plonky2-field 1.1.1 (pinned in /root/reference/Cargo.lock:514-517; the type
// contract is /root/reference/common/src/circuit.rs:10-12: F = GoldilocksField, D = 2).
// The crate source is not vendored in the reference tree; this restates its published algorithm
//  and is pinned by the reference's own fixtures through the verifier
// (oracle/verifier.cpp accepts wormhole/bench-data/proof.bin) and the Poseidon KATs.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <array>
#include <stdexcept>
#include <string>

namespace orc {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using u128 = unsigned __int128;

constexpr u64 P = 0xFFFFFFFF00000001ULL;   // 2^64 - 2^32 + 1
constexpr u64 EPS = 0xFFFFFFFFULL;         // 2^64 mod P
constexpr u64 GEN = 0xc65c18b67785d900ULL; // multiplicative generator g = coset shift (A.1)
constexpr u64 TWO_ADIC_ROOT = 0x64fdd1a46201e246ULL; // order 2^32 (A.1): g^((p-1)/2^32)
constexpr u64 EXT_W = 7;                   // F_{p^2} = F_p[X]/(X^2 - 7)

// (written with conditional expressions rather than branches: the conditions are data-dependent coin flips, and a mispredicted
// branch costs more than the whole reduction)
inline u64 fadd(u64 a, u64 b) {
    u64 s = a + b;
    return s - (((s < a) | (s >= P)) ? P : 0);
}
inline u64 fsub(u64 a, u64 b) { return a - b + ((a < b) ? P : 0); }
inline u64 fneg(u64 a) { return a ? P - a : 0; }
inline u64 reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    // x = lo + hi_lo*2^64 + hi_hi*2^96,  2^64 = 2^32-1,  2^96 = -1  (mod P)
    u64 t0 = lo - hi_hi;
    t0 -= (lo < hi_hi) ? EPS : 0;          // borrow: subtracting 2^64 is subtracting EPS
    u64 t1 = hi_lo * EPS;                  // < 2^64, no overflow
    u64 r = t0 + t1;
    r += (r < t0) ? EPS : 0;               // carry: adding 2^64 is adding EPS (cannot carry twice)
    return r - ((r >= P) ? P : 0);
}
inline u64 fmul(u64 a, u64 b) { return reduce128((u128)a * b); }
inline u64 fsqr(u64 a) { return fmul(a, a); }
inline u64 fpow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = fmul(r, a);
        a = fmul(a, a);
        e >>= 1;
    }
    return r;
}
inline u64 finv(u64 a) {
    if (a == 0) throw std::runtime_error("finv(0)");
    return fpow(a, P - 2);
}
inline u64 from_u64(u64 x) { return x >= P ? x - P : x; }
// primitive 2^k-th root of unity (A.1): root(k) = T^(2^(32-k))
inline u64 root_of_unity(unsigned k) {
    if (k > 32) throw std::runtime_error("root_of_unity: k > 32");
    u64 r = TWO_ADIC_ROOT;
    for (unsigned i = k; i < 32; ++i) r = fsqr(r);
    return r;
}
inline unsigned log2_strict(size_t n) {
    unsigned k = 0;
    while ((size_t(1) << k) < n) ++k;
    if ((size_t(1) << k) != n) throw std::runtime_error("log2_strict: not a power of two");
    return k;
}
inline size_t reverse_bits(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// ---- quadratic extension ----
struct E2 {
    u64 a = 0, b = 0;  // a + b*X
    E2() = default;
    E2(u64 a_, u64 b_) : a(a_), b(b_) {}
    explicit E2(u64 a_) : a(a_), b(0) {}
    bool operator==(const E2& o) const { return a == o.a && b == o.b; }
    bool operator!=(const E2& o) const { return !(*this == o); }
};
inline E2 operator+(E2 x, E2 y) { return {fadd(x.a, y.a), fadd(x.b, y.b)}; }
inline E2 operator-(E2 x, E2 y) { return {fsub(x.a, y.a), fsub(x.b, y.b)}; }
inline E2 operator*(E2 x, E2 y) {
    return {fadd(fmul(x.a, y.a), fmul(EXT_W, fmul(x.b, y.b))), fadd(fmul(x.a, y.b), fmul(x.b, y.a))};
}
inline E2 emul_base(E2 x, u64 s) { return {fmul(x.a, s), fmul(x.b, s)}; }
inline E2 einv(E2 x) {
    u64 norm = fsub(fsqr(x.a), fmul(EXT_W, fsqr(x.b)));
    u64 ni = finv(norm);
    return {fmul(x.a, ni), fmul(fneg(x.b), ni)};
}
inline E2 epow(E2 x, u64 e) {
    E2 r(1);
    while (e) {
        if (e & 1) r = r * x;
        x = x * x;
        e >>= 1;
    }
    return r;
}
inline E2 epow2k(E2 x, unsigned k) {
    for (unsigned i = 0; i < k; ++i) x = x * x;
    return x;
}

// Generic helpers so gate evaluators can be written once for the base field and the extension.
struct BaseOps {
    using T = u64;
    static T zero() { return 0; }
    static T one() { return 1; }
    static T from(u64 c) { return from_u64(c); }
    static T add(T x, T y) { return fadd(x, y); }
    static T sub(T x, T y) { return fsub(x, y); }
    static T mul(T x, T y) { return fmul(x, y); }
    static T mulc(T x, u64 c) { return fmul(x, c); }
};
struct ExtOps {
    using T = E2;
    static T zero() { return E2(); }
    static T one() { return E2(1); }
    static T from(u64 c) { return E2(from_u64(c)); }
    static T add(T x, T y) { return x + y; }
    static T sub(T x, T y) { return x - y; }
    static T mul(T x, T y) { return x * y; }
    static T mulc(T x, u64 c) { return emul_base(x, c); }
};

}  // namespace orc
