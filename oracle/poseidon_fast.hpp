// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// A FAST host implementation of the same width-12 Poseidon permutation as poseidon.hpp, used by the CPU BASELINE arm
// (bench.py --impl reference) so that the GPU / CPU ratio is taken against an optimised CPU prover rather than the
// readable restatement: qp-plonky2's own `PoseidonGoldilocks` is hand-tuned (lazy reductions, unrolled rounds, SIMD on
// x86), and VERDICT r01 asked for "an honest CPU arm". Written independently of the product's device / host-transcript
// code, for AVX2 (the oracle is built with -march=x86-64-v3):
//   * the 12 state words live in three 4 x u64 vectors, lazily reduced (any value in [0, 2^64)) between rounds;
//   * S-box: 64 x 64 -> 128-bit products from four vpmuludq, reduced with 2^64 = 2^32 - 1, 2^96 = -1 (branch-free masks); in
//     the 22 partial rounds the single S-box is scalar;
//   * dense MDS on the 32-bit halves: out = sum_i C[i] * rot_i(x) with the rotations read as unaligned loads from a doubled
//     copy of the state, 64-bit lanes never overflow (13 x 41 x 2^32 < 2^42), one recombination per round.
// The naive form in poseidon.hpp stays the definition; tests/test_oracle_kats.py checks this one against it on random
// states, corner states and the reference's known answers. Selected at run time: orc_set_fast(1).
#pragma once
#include "poseidon.hpp"
#include <immintrin.h>

namespace orc {

namespace fastp {
static inline u64 red128(u128 x) {            // any 128-bit value -> same residue in [0, 2^64)
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & EPS, t0, r;
    t0 = lo - hh;
    t0 -= (lo < hh) ? EPS : 0;                // borrowed 2^64 = EPS (mod p); cannot borrow twice: hh < 2^32
    u64 t1 = (hl << 32) - hl;                 // hl * (2^32 - 1)
    r = t0 + t1;
    r += (r < t1) ? EPS : 0;                  // wrapped: add 2^64 mod p; cannot wrap twice
    return r;
}
static inline u64 mul(u64 a, u64 b) { return red128((u128)a * b); }
static inline u64 sbox(u64 x) {
    u64 x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x2, x);
    return mul(x3, x4);
}

#if defined(__AVX2__)
typedef __m256i V;
static inline V vset(u64 x) { return _mm256_set1_epi64x((long long)x); }
static inline V ult(V a, V b) {               // unsigned a < b per lane -> all-ones mask
    const V s = vset(0x8000000000000000ULL);
    return _mm256_cmpgt_epi64(_mm256_xor_si256(b, s), _mm256_xor_si256(a, s));
}
static inline V vreduce(V hi, V lo) {          // (hi:lo) -> lazy u64
    const V eps = vset(EPS);
    V hh = _mm256_srli_epi64(hi, 32), hl = _mm256_and_si256(hi, eps);
    V t0 = _mm256_sub_epi64(lo, hh);
    t0 = _mm256_sub_epi64(t0, _mm256_and_si256(ult(lo, hh), eps));
    V t1 = _mm256_sub_epi64(_mm256_slli_epi64(hl, 32), hl);
    V r = _mm256_add_epi64(t0, t1);
    return _mm256_add_epi64(r, _mm256_and_si256(ult(r, t1), eps));
}
static inline V vmul(V a, V b) {               // lazy product of lazy inputs
    const V m = vset(EPS);
    V a1 = _mm256_srli_epi64(a, 32), b1 = _mm256_srli_epi64(b, 32);
    V p00 = _mm256_mul_epu32(a, b), p01 = _mm256_mul_epu32(a, b1), p10 = _mm256_mul_epu32(a1, b), p11 = _mm256_mul_epu32(a1, b1);
    V mid = _mm256_add_epi64(p01, _mm256_srli_epi64(p00, 32));
    V mid2 = _mm256_add_epi64(p10, _mm256_and_si256(mid, m));
    V hi = _mm256_add_epi64(_mm256_add_epi64(p11, _mm256_srli_epi64(mid, 32)), _mm256_srli_epi64(mid2, 32));
    V lo = _mm256_or_si256(_mm256_slli_epi64(mid2, 32), _mm256_and_si256(p00, m));
    return vreduce(hi, lo);
}
static inline V vsbox(V x) {
    V x2 = vmul(x, x), x4 = vmul(x2, x2), x3 = vmul(x2, x);
    return vmul(x3, x4);
}
static inline V vadd_lazy(V a, V c) {          // c canonical: at most one wrap
    V s = _mm256_add_epi64(a, c);
    return _mm256_add_epi64(s, _mm256_and_si256(ult(s, a), vset(EPS)));
}
// dense MDS: s[0..2] hold words 0-3, 4-7, 8-11
static inline void vmds(V* s) {
    const V m = vset(EPS);
    alignas(32) u64 lo[24], hi[24];
    for (int k = 0; k < 3; ++k) {
        V l = _mm256_and_si256(s[k], m), h = _mm256_srli_epi64(s[k], 32);
        _mm256_store_si256((V*)(lo + 4 * k), l); _mm256_store_si256((V*)(lo + 12 + 4 * k), l);
        _mm256_store_si256((V*)(hi + 4 * k), h); _mm256_store_si256((V*)(hi + 12 + 4 * k), h);
    }
    V al[3] = {_mm256_setzero_si256(), _mm256_setzero_si256(), _mm256_setzero_si256()}, ah[3] = {al[0], al[0], al[0]};
#pragma GCC unroll 12
    for (int i = 0; i < 12; ++i) {
        const V c = vset(MDS_CIRC[i]);
#pragma GCC unroll 3
        for (int k = 0; k < 3; ++k) {
            al[k] = _mm256_add_epi64(al[k], _mm256_mul_epu32(_mm256_loadu_si256((const V*)(lo + i + 4 * k)), c));
            ah[k] = _mm256_add_epi64(ah[k], _mm256_mul_epu32(_mm256_loadu_si256((const V*)(hi + i + 4 * k)), c));
        }
    }
    // + diag(8, 0, ...): word 0 only
    const V d = _mm256_set_epi64x(0, 0, 0, (long long)MDS_DIAG[0]);
    al[0] = _mm256_add_epi64(al[0], _mm256_mul_epu32(_mm256_load_si256((const V*)lo), d));
    ah[0] = _mm256_add_epi64(ah[0], _mm256_mul_epu32(_mm256_load_si256((const V*)hi), d));
    // value = al + ah 2^32 with al, ah < 2^42: (hi:lo) = (ah >> 32 : (ah << 32) + al) with the add's carry into hi
    for (int k = 0; k < 3; ++k) {
        V lo64 = _mm256_add_epi64(_mm256_slli_epi64(ah[k], 32), al[k]);
        V carry = _mm256_srli_epi64(ult(lo64, al[k]), 63);
        s[k] = vreduce(_mm256_add_epi64(_mm256_srli_epi64(ah[k], 32), carry), lo64);
    }
}
#endif
}  // namespace fastp

inline void poseidon_permute_fast_impl(u64* st) {
    using namespace fastp;
    const u64* rc = poseidon_round_constants();
#if defined(__AVX2__)
    V s[3] = {_mm256_loadu_si256((const V*)st), _mm256_loadu_si256((const V*)(st + 4)), _mm256_loadu_si256((const V*)(st + 8))};
    for (int r = 0; r < N_ROUNDS; ++r) {
        for (int k = 0; k < 3; ++k) s[k] = vadd_lazy(s[k], _mm256_loadu_si256((const V*)(rc + 12 * r + 4 * k)));
        if (r < HALF_N_FULL_ROUNDS || r >= HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS) {
            for (int k = 0; k < 3; ++k) s[k] = vsbox(s[k]);
        } else {
            u64 x0 = sbox((u64)_mm256_extract_epi64(s[0], 0));
            s[0] = _mm256_insert_epi64(s[0], (long long)x0, 0);
        }
        vmds(s);
    }
    alignas(32) u64 out[12];
    for (int k = 0; k < 3; ++k) _mm256_store_si256((V*)(out + 4 * k), s[k]);
    for (int i = 0; i < 12; ++i) st[i] = out[i] >= P ? out[i] - P : out[i];
#else
    u64 s[12];
    for (int i = 0; i < 12; ++i) s[i] = st[i];
    for (int r = 0; r < N_ROUNDS; ++r) {
        const bool full = r < HALF_N_FULL_ROUNDS || r >= HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS;
        for (int i = 0; i < 12; ++i) { u64 v = s[i] + rc[12 * r + i]; s[i] = v + ((v < s[i]) ? EPS : 0); }
        for (int i = 0; i < (full ? 12 : 1); ++i) s[i] = sbox(s[i]);
        u64 o[12];
        for (int q = 0; q < 12; ++q) {
            u128 acc = (u128)s[q] * MDS_DIAG[q];
            for (int i = 0; i < 12; ++i) acc += (u128)s[(i + q) % 12] * MDS_CIRC[i];
            o[q] = red128(acc);
        }
        for (int i = 0; i < 12; ++i) s[i] = o[i];
    }
    for (int i = 0; i < 12; ++i) st[i] = s[i] >= P ? s[i] - P : s[i];
#endif
}

}  // namespace orc
