// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// A 4-lane AVX2 instance of the `Ops` interface that gates.hpp is written over: the CPU-baseline arm (orc_set_fast(1))
// evaluates the vanishing polynomial at FOUR LDE points per call with the SAME generic gate code the readable restatement
// instantiates over u64 (prover) and F_{p^2} (verifier) — nothing about the constraints is restated here, only the field
// operations. Every lane holds a canonical element (< p) between operations, as BaseOps does. qp-plonky2's CPU prover evaluates
// gates in packed batches in the same way (`eval_unfiltered_base_batch` over `PackedField`).
// tests/test_oracle_prover.py::test_fast_mode_proof_is_byte_identical pins it against the scalar path.
#pragma once
#include "poseidon_fast.hpp"

#if defined(__AVX2__)
namespace orc {

struct V4 {
    __m256i v;
};

struct VecOps {
    using T = V4;
    using V = __m256i;
    static V bc(u64 x) { return _mm256_set1_epi64x((long long)x); }
    static V canon(V r) {                       // lazy [0, 2^64) -> canonical: r >= p <=> !(r < p)
        const V p = bc(P);
        return _mm256_sub_epi64(r, _mm256_andnot_si256(fastp::ult(r, p), p));
    }
    static T zero() { return {_mm256_setzero_si256()}; }
    static T one() { return {bc(1)}; }
    static T from(u64 c) { return {bc(from_u64(c))}; }
    static T add(T a, T b) {                    // canonical inputs: at most one subtraction of p
        const V p = bc(P);
        V s = _mm256_add_epi64(a.v, b.v);
        V over = _mm256_or_si256(fastp::ult(s, a.v), _mm256_xor_si256(fastp::ult(s, p), bc(~0ULL)));
        return {_mm256_sub_epi64(s, _mm256_and_si256(over, p))};
    }
    static T sub(T a, T b) {
        V d = _mm256_sub_epi64(a.v, b.v);
        return {_mm256_add_epi64(d, _mm256_and_si256(fastp::ult(a.v, b.v), bc(P)))};
    }
    static T mul(T a, T b) { return {canon(fastp::vmul(a.v, b.v))}; }
    static T mulc(T a, u64 c) { return {canon(fastp::vmul(a.v, bc(from_u64(c))))}; }
    static T lanes(u64 a, u64 b, u64 c, u64 d) { return {_mm256_set_epi64x((long long)d, (long long)c, (long long)b, (long long)a)}; }
    static void store(T x, u64 out[4]) { _mm256_storeu_si256((V*)out, x.v); }
};

// MDS layer on four states at once: products of the 32-bit halves with the small circulant constants never overflow a 64-bit
// lane (13 x 41 x 2^32 < 2^42); one 128-bit recombination and reduction per output word.
template <>
inline void mds_layer<VecOps>(V4* st) {
    using V = __m256i;
    const V m = VecOps::bc(EPS);
    V lo[12], hi[12];
    for (int i = 0; i < 12; ++i) { lo[i] = _mm256_and_si256(st[i].v, m); hi[i] = _mm256_srli_epi64(st[i].v, 32); }
    V4 out[12];
    for (int r = 0; r < 12; ++r) {
        V al = _mm256_mul_epu32(lo[r], VecOps::bc(MDS_DIAG[r])), ah = _mm256_mul_epu32(hi[r], VecOps::bc(MDS_DIAG[r]));
        for (int i = 0; i < 12; ++i) {
            const V c = VecOps::bc(MDS_CIRC[i]);
            al = _mm256_add_epi64(al, _mm256_mul_epu32(lo[(i + r) % 12], c));
            ah = _mm256_add_epi64(ah, _mm256_mul_epu32(hi[(i + r) % 12], c));
        }
        V lo64 = _mm256_add_epi64(_mm256_slli_epi64(ah, 32), al);
        V carry = _mm256_srli_epi64(fastp::ult(lo64, al), 63);
        out[r].v = VecOps::canon(fastp::vreduce(_mm256_add_epi64(_mm256_srli_epi64(ah, 32), carry), lo64));
    }
    for (int r = 0; r < 12; ++r) st[r] = out[r];
}

}  // namespace orc
#endif
