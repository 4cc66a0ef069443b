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
        for (size_t base = 0; base < n; base += m)
            for (size_t k = 0; k < h; ++k) {
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

}  // namespace orc
