// AVX2 form of the host-side Poseidon permutation behind the Fiat-Shamir transcript (host_transcript.hpp). The transcript is
// strictly serial and sits on every proof's critical path (119 permutations, ~0.33 ms with the scalar form); this file is
// compiled with -mavx2 on its own and selected at run time when the CPU has AVX2, the scalar form stays the fallback.
// State: three 4 x u64 vectors, lazily reduced between rounds. S-box: 64 x 64 products from four vpmuludq, folded with
// 2^64 = 2^32 - 1 and 2^96 = -1 through branch-free masks; the single S-box of a partial round is scalar. MDS: the circulant
// applied to the 32-bit halves, rotations read as unaligned loads from a doubled copy of the half-state (13 x 41 x 2^32 < 2^42
// fits a lane), one recombination per word.
#include <immintrin.h>
#include <cstdint>

namespace zkb {
typedef uint64_t u64;
const u64* host_round_constants_ptr();      // host_transcript.hpp's table (one copy per library)

namespace {
constexpr u64 EPS = 0xFFFFFFFFULL, P = 0xFFFFFFFF00000001ULL;
constexpr u64 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
typedef __m256i V;
inline V bc(u64 x) { return _mm256_set1_epi64x((long long)x); }
inline V below(V a, V b) {       // unsigned a < b
    const V s = bc(0x8000000000000000ULL);
    return _mm256_cmpgt_epi64(_mm256_xor_si256(b, s), _mm256_xor_si256(a, s));
}
inline V fold(V hi, V lo) {      // 128-bit (hi:lo) -> lazy 64-bit
    const V eps = bc(EPS);
    const V hh = _mm256_srli_epi64(hi, 32), hl = _mm256_and_si256(hi, eps);
    V t = _mm256_sub_epi64(lo, hh);
    t = _mm256_sub_epi64(t, _mm256_and_si256(below(lo, hh), eps));
    const V u = _mm256_sub_epi64(_mm256_slli_epi64(hl, 32), hl);
    const V r = _mm256_add_epi64(t, u);
    return _mm256_add_epi64(r, _mm256_and_si256(below(r, u), eps));
}
inline V mul(V a, V b) {
    const V m = bc(EPS);
    const V ah = _mm256_srli_epi64(a, 32), bh = _mm256_srli_epi64(b, 32);
    const V ll = _mm256_mul_epu32(a, b), lh = _mm256_mul_epu32(a, bh), hl = _mm256_mul_epu32(ah, b), hh = _mm256_mul_epu32(ah, bh);
    const V mid = _mm256_add_epi64(lh, _mm256_srli_epi64(ll, 32));
    const V mid2 = _mm256_add_epi64(hl, _mm256_and_si256(mid, m));
    const V hi = _mm256_add_epi64(_mm256_add_epi64(hh, _mm256_srli_epi64(mid, 32)), _mm256_srli_epi64(mid2, 32));
    const V lo = _mm256_or_si256(_mm256_slli_epi64(mid2, 32), _mm256_and_si256(ll, m));
    return fold(hi, lo);
}
inline V pow7(V x) {
    const V x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2);
    return mul(x3, x4);
}
inline u64 mul1(u64 a, u64 b) {
    const unsigned __int128 x = (unsigned __int128)a * b;
    const u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & EPS;
    u64 t = lo - hh;
    t -= (lo < hh) ? EPS : 0;
    const u64 u = (hl << 32) - hl;
    u64 r = t + u;
    r += (r < u) ? EPS : 0;
    return r;
}
}  // namespace

void h_poseidon_permute_avx2(u64* st) {
    const u64* rc = host_round_constants_ptr();
    const V m = bc(EPS), eps = m;
    V s[3] = {_mm256_loadu_si256((const V*)st), _mm256_loadu_si256((const V*)(st + 4)), _mm256_loadu_si256((const V*)(st + 8))};
    alignas(32) u64 lo[24], hi[24];
    for (int r = 0; r < 30; ++r) {
        for (int k = 0; k < 3; ++k) {              // + round constants (canonical): at most one wrap
            const V c = _mm256_loadu_si256((const V*)(rc + 12 * r + 4 * k)), v = _mm256_add_epi64(s[k], c);
            s[k] = _mm256_add_epi64(v, _mm256_and_si256(below(v, c), eps));
        }
        if (r < 4 || r >= 26) {
            for (int k = 0; k < 3; ++k) s[k] = pow7(s[k]);
        } else {
            const u64 x = (u64)_mm256_extract_epi64(s[0], 0), x2 = mul1(x, x), x3 = mul1(x2, x), x4 = mul1(x2, x2);
            s[0] = _mm256_insert_epi64(s[0], (long long)mul1(x3, x4), 0);
        }
        for (int k = 0; k < 3; ++k) {
            const V l = _mm256_and_si256(s[k], m), h = _mm256_srli_epi64(s[k], 32);
            _mm256_store_si256((V*)(lo + 4 * k), l); _mm256_store_si256((V*)(lo + 12 + 4 * k), l);
            _mm256_store_si256((V*)(hi + 4 * k), h); _mm256_store_si256((V*)(hi + 12 + 4 * k), h);
        }
        V al[3], ah[3];
        for (int k = 0; k < 3; ++k) al[k] = ah[k] = _mm256_setzero_si256();
        for (int i = 0; i < 12; ++i) {
            const V c = bc(CIRC[i]);
            for (int k = 0; k < 3; ++k) {
                al[k] = _mm256_add_epi64(al[k], _mm256_mul_epu32(_mm256_loadu_si256((const V*)(lo + i + 4 * k)), c));
                ah[k] = _mm256_add_epi64(ah[k], _mm256_mul_epu32(_mm256_loadu_si256((const V*)(hi + i + 4 * k)), c));
            }
        }
        const V d = _mm256_set_epi64x(0, 0, 0, 8);     // + 8 x[0] on word 0
        al[0] = _mm256_add_epi64(al[0], _mm256_mul_epu32(_mm256_load_si256((const V*)lo), d));
        ah[0] = _mm256_add_epi64(ah[0], _mm256_mul_epu32(_mm256_load_si256((const V*)hi), d));
        for (int k = 0; k < 3; ++k) {                  // al + ah 2^32 as (hi:lo), then fold
            const V lo64 = _mm256_add_epi64(_mm256_slli_epi64(ah[k], 32), al[k]);
            const V carry = _mm256_srli_epi64(below(lo64, al[k]), 63);
            s[k] = fold(_mm256_add_epi64(_mm256_srli_epi64(ah[k], 32), carry), lo64);
        }
    }
    alignas(32) u64 out[12];
    for (int k = 0; k < 3; ++k) _mm256_store_si256((V*)(out + 4 * k), s[k]);
    for (int i = 0; i < 12; ++i) st[i] = out[i] >= P ? out[i] - P : out[i];
}

}  // namespace zkb
