// Poseidon round constants for the Goldilocks width-12 permutation, regenerated the way qp-plonky2 does
// (ChaCha8 stream of rand's seed_from_u64(0), rejection-sampled into [0, p)). Host only.
#pragma once
#include <mutex>
#include "field.cuh"

namespace zkb {
namespace detail {
inline u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
inline void chacha8_block(const u32 key[8], u64 counter, u32 out[16]) {
    u32 s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; ++i) s[4 + i] = key[i];
    s[12] = (u32)counter; s[13] = (u32)(counter >> 32); s[14] = 0; s[15] = 0;
    u32 x[16];
    for (int i = 0; i < 16; ++i) x[i] = s[i];
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < 4; ++r) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}
}  // namespace detail

inline const u64* host_round_constants() {
    static u64 rc[360];
    static std::once_flag once;
    std::call_once(once, [] {
        u64 state = 0;
        u32 key[8];
        for (int i = 0; i < 8; ++i) {   // PCG32 seed expansion
            state = state * 6364136223846793005ULL + 11634580027462260723ULL;
            u32 xs = (u32)(((state >> 18) ^ state) >> 27), rot = (u32)(state >> 59);
            key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
        u32 blk[16];
        int pos = 16, n = 0;
        u64 ctr = 0;
        auto next = [&]() { if (pos == 16) { detail::chacha8_block(key, ctr++, blk); pos = 0; } return blk[pos++]; };
        while (n < 360) {
            u64 lo = next(), hi = next();
            unsigned __int128 m = (unsigned __int128)(lo | (hi << 32)) * GL_P;   // uniform sample in [0, p)
            if ((u64)m <= GL_P - 1) rc[n++] = (u64)(m >> 64);
        }
    });
    return rc;
}


// limb form used by poseidon_permute_limbs: rc3[36 r + 3 j + k], r = 0..30 (row 30 = the constant 0), BIASED:
// limbs (22, 22, 20 bits) of (value - (2^32 - 1)) mod p, plus 2^20 in limb 2 — the same field element, since
// 2^20 * 2^44 = 2^64 = 2^32 - 1 (mod p), but every word's top limb is then >= 2^20 and its carry never negative.
constexpr int RC3_WORDS = 3 * 12 * 31;
inline const u32* host_round_constant_limbs() {
    static u32 rc3[RC3_WORDS];
    static std::once_flag once;
    std::call_once(once, [] {
        const u64* rc = host_round_constants();
        for (int i = 0; i < 12 * 31; ++i) {
            const u64 v = i < 360 ? rc[i] : 0;
            const u64 w = gl_sub(v, GL_EPS);
            rc3[3 * i] = (u32)(w & 0x3FFFFF);
            rc3[3 * i + 1] = (u32)((w >> 22) & 0x3FFFFF);
            rc3[3 * i + 2] = (u32)(w >> 44) + (1u << 20);
        }
    });
    return rc3;
}

}  // namespace zkb
