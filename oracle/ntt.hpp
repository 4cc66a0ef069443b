// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// CPU restatement of qp-plonky2-field 1.1.1 `fft`, `ifft`, `lde`, `coset_fft`, `coset_ifft`
// (SURVEY.md §8 a2, A.1), reached from /root/reference/wormhole/prover/src/lib.rs:234-236.
// Plain textbook radix-2 (bit-reverse then DIT), natural order in and out; the transforms are
// mathematically unique so only the definition matters:
//   fft(c)[i]        = sum_k c[k] * w^(ik),           w = root_of_unity(log2 n)
//   coset_fft(c,s)[i]= sum_k c[k] * (s*w^i)^k
#pragma once
#include "goldilocks.hpp"

namespace orc {

template <class T, class MulBase>
inline void ntt_inplace(std::vector<T>& a, bool inverse, MulBase mulb, T (*addf)(T, T), T (*subf)(T, T)) {
    size_t n = a.size();
    unsigned lg = log2_strict(n);
    for (size_t i = 0; i < n; ++i) {
        size_t j = reverse_bits(i, lg);
        if (i < j) std::swap(a[i], a[j]);
    }
    for (unsigned s = 1; s <= lg; ++s) {
        size_t m = size_t(1) << s, h = m >> 1;
        u64 wm = root_of_unity(s);
        if (inverse) wm = finv(wm);
        std::vector<u64> tw(h);
        tw[0] = 1;
        for (size_t k = 1; k < h; ++k) tw[k] = fmul(tw[k - 1], wm);
        // one flat loop over the n / 2 butterflies of the stage: a top-level call (FRI's extension-field transforms, the
        // quotient's coset iNTT) spreads it over the host cores; inside an outer parallel loop over columns it runs serially
#pragma omp parallel for schedule(static) if (n >= 8192)
        for (long idx = 0; idx < (long)(n >> 1); ++idx) {
            const size_t base = ((size_t)idx / h) * m, k = (size_t)idx % h;
            T u = a[base + k], v = mulb(a[base + k + h], tw[k]);
            a[base + k] = addf(u, v);
            a[base + k + h] = subf(u, v);
        }
    }
    if (inverse) {
        u64 ninv = finv(from_u64(n));
        for (auto& x : a) x = mulb(x, ninv);
    }
}

inline u64 add_b(u64 a, u64 b) { return fadd(a, b); }
inline u64 sub_b(u64 a, u64 b) { return fsub(a, b); }
inline E2 add_e(E2 a, E2 b) { return a + b; }
inline E2 sub_e(E2 a, E2 b) { return a - b; }

inline void fft(std::vector<u64>& a) { ntt_inplace<u64>(a, false, [](u64 x, u64 s) { return fmul(x, s); }, add_b, sub_b); }
inline void ifft(std::vector<u64>& a) { ntt_inplace<u64>(a, true, [](u64 x, u64 s) { return fmul(x, s); }, add_b, sub_b); }
inline void fft(std::vector<E2>& a) { ntt_inplace<E2>(a, false, [](E2 x, u64 s) { return emul_base(x, s); }, add_e, sub_e); }
inline void ifft(std::vector<E2>& a) { ntt_inplace<E2>(a, true, [](E2 x, u64 s) { return emul_base(x, s); }, add_e, sub_e); }

// coefficients -> evaluations on shift*<w_n>
template <class T>
inline void coset_fft(std::vector<T>& c, u64 shift);
template <>
inline void coset_fft<u64>(std::vector<u64>& c, u64 shift) {
    u64 s = 1;
    for (auto& x : c) { x = fmul(x, s); s = fmul(s, shift); }
    fft(c);
}
template <>
inline void coset_fft<E2>(std::vector<E2>& c, u64 shift) {
    u64 s = 1;
    for (auto& x : c) { x = emul_base(x, s); s = fmul(s, shift); }
    fft(c);
}
// evaluations on shift*<w_n> -> coefficients
inline void coset_ifft(std::vector<u64>& v, u64 shift) {
    ifft(v);
    u64 si = finv(shift), s = 1;
    for (auto& x : v) { x = fmul(x, s); s = fmul(s, si); }
}

// zero-pad coefficients by 2^rate_bits and evaluate on GEN*<w_{n<<rate_bits}>, natural order
template <class T>
inline std::vector<T> lde_coset(const std::vector<T>& coeffs, unsigned rate_bits, u64 shift = GEN) {
    std::vector<T> v(coeffs);
    v.resize(coeffs.size() << rate_bits, T());
    coset_fft<T>(v, shift);
    return v;
}

// ---- fast path of the CPU-baseline arm (orc_set_fast(1)); same results as lde_coset<u64> (tests/test_oracle_prover.py) ----
// The textbook form above transforms the zero-padded 8n-vector and rebuilds its twiddles per call. Here: the padded transform
// splits into 2^rate_bits coset transforms of size n (X[j + R i] = sum_k c[k] (s w_N^j)^k w_n^(i k)), twiddles come from one
// cached table per size, products are branch-free lazy reductions, butterflies keep values canonical with conditional moves.
namespace fastntt {
static inline u64 mul(u64 a, u64 b) {
    u128 x = (u128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh;
    t0 -= (lo < hh) ? EPS : 0;
    u64 t1 = (hl << 32) - hl, r = t0 + t1;
    r += (r < t1) ? EPS : 0;
    return r >= P ? r - P : r;
}
static inline u64 add(u64 a, u64 b) { u64 s = a + b; return (s < a || s >= P) ? s - P : s; }
static inline u64 sub(u64 a, u64 b) { u64 d = a - b; return a < b ? d + P : d; }
// table[k] = w_n^k, k < n / 2 (cached per size; built under a lock)
const std::vector<u64>& twiddles(unsigned lg);
// in-place forward DIF, natural order in, bit-reversed order out
static inline void dif(u64* a, unsigned lg, const std::vector<u64>& tw) {
    const size_t n = size_t(1) << lg;
    for (unsigned s = lg; s >= 1; --s) {
        const size_t m = size_t(1) << s, h = m >> 1, stride = n >> s;
        for (size_t base = 0; base < n; base += m)
            for (size_t k = 0; k < h; ++k) {
                const u64 u = a[base + k], v = a[base + k + h];
                a[base + k] = add(u, v);
                a[base + k + h] = mul(sub(u, v), tw[k * stride]);
            }
    }
}
}  // namespace fastntt

// out[reverse_bits(natural index)] layout is what the Merkle leaves want, so the fast path returns the LDE directly in LEAF
// order: leaf l = bitrev_{lgN}(j + R i) = bitrev_r(j) * n + bitrev_lg(i): block bitrev_r(j) is coset j in bit-reversed order —
// exactly what the DIF transform leaves behind.
inline void lde_coset_leaf_order_fast(const std::vector<u64>& coeffs, unsigned rate_bits, u64 shift, u64* out /* N */) {
    const size_t n = coeffs.size(), R = size_t(1) << rate_bits;
    const unsigned lg = log2_strict(n);
    const auto& tw = fastntt::twiddles(lg);
    const u64 wN = root_of_unity(lg + rate_bits);
    for (size_t j = 0; j < R; ++j) {
        u64* blk = out + reverse_bits(j, rate_bits) * n;
        const u64 sj = fmul(shift, fpow(wN, j));
        u64 p = 1;
        for (size_t k = 0; k < n; ++k) { blk[k] = fastntt::mul(coeffs[k], p); p = fastntt::mul(p, sj); }
        fastntt::dif(blk, lg, tw);
    }
}

}  // namespace orc
