// Host-side Poseidon sponge + duplex challenger for the Fiat-Shamir transcript
// (qp-plonky2 1.1.1 `iop::challenger::Challenger<F, PoseidonHash>`). The transcript is
// strictly serial (~100 permutations per proof), so it runs on the host between kernel launches.
#pragma once
#include <chrono>
#include <vector>
#include "field.cuh"
#include "kernels.h"

namespace zkb {

// ---- fast host permutation. The transcript is on the critical path of every proof (~130 permutations between
// kernel launches), so it gets the same algebra as the device code: branch-free lazy reduction, and the MDS layer as the
// shift/add circulant of poseidon.cuh applied to the 32-bit halves of the state (sums stay below 2^41 in u64 lanes).
inline u64 h_mul_lazy(u64 a, u64 b) {                       // any u64 in, result in [0, 2^64) with the same residue
    unsigned __int128 x = (unsigned __int128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    t0 -= (lo < hh) ? GL_EPS : 0;                           // borrowed: subtract 2^64 mod p
    u64 t1 = (hl << 32) - hl, r = t0 + t1;
    r += (r < t1) ? GL_EPS : 0;
    return r;
}
inline u64 h_sbox7(u64 x) {
    u64 x2 = h_mul_lazy(x, x), x4 = h_mul_lazy(x2, x2), x3 = h_mul_lazy(x2, x);
    return h_mul_lazy(x3, x4);
}
// out[r] = sum_i x[(i+r) % 12] C[i] + 8 x[0] [r == 0], arithmetic mod 2^64 on values < 2^32 (true sums < 2^41)
inline void h_mds_half(u64* x) {
    u64 P[3], M[3], R[3], I[3];
    for (int b = 0; b < 3; ++b) {
        u64 t0 = x[b] + x[6 + b], t1 = x[3 + b] + x[9 + b];
        P[b] = t0 + t1; M[b] = t0 - t1; R[b] = x[b] - x[6 + b]; I[b] = x[3 + b] - x[9 + b];
    }
    const u64 x0 = x[0], T = P[0] + P[1] + P[2];
    const u64 G[3] = {T + P[2], T + P[0], T + P[1]};
    const u64 F[3] = {8 * M[2] - M[0] - 2 * M[1], 0 - M[1] - 2 * M[2] - 8 * M[0], 2 * M[0] - M[2] - 8 * M[1]};
    u64 Ar[3], Ai[3], Br[3], Bi[3], Dr[3], Di[3];
    for (int b = 0; b < 3; ++b) {
        Ar[b] = 2 * R[b] - I[b];  Ai[b] = R[b] + 2 * I[b];
        Br[b] = R[b] - 16 * I[b]; Bi[b] = 16 * R[b] + I[b];
        Dr[b] = R[b] + 4 * I[b];  Di[b] = I[b] - 4 * R[b];
    }
    const u64 Re[3] = {Ar[0] + Br[1] + Dr[2], Ar[1] + Br[2] + Di[0], Ar[2] + Bi[0] + Di[1]};
    const u64 Im[3] = {Ai[0] + Bi[1] + Di[2], Ai[1] + Bi[2] - Dr[0], Ai[2] - Br[0] - Dr[1]};
    for (int b = 0; b < 3; ++b) {
        u64 U = 16 * G[b] + F[b], V = 16 * G[b] - F[b];
        x[b] = U + Re[b]; x[6 + b] = U - Re[b]; x[3 + b] = V + Im[b]; x[9 + b] = V - Im[b];
    }
    x[0] += 8 * x0;
}
// host_poseidon_avx2.cpp (compiled with -mavx2, used when the CPU has it): the same permutation, ~2.5x faster
void h_poseidon_permute_avx2(u64* s);
const u64* host_round_constants_ptr();
inline void h_poseidon_permute_scalar(u64* s) {
    const u64* rc = host_round_constants();
    for (int r = 0; r < 30; ++r) {
        u64 t[12];
        for (int i = 0; i < 12; ++i) {                      // lazy add of a canonical constant: at most one wrap
            u64 v = s[i] + rc[12 * r + i];
            t[i] = v + ((v < s[i]) ? GL_EPS : 0);
        }
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; ++i) t[i] = h_sbox7(t[i]);
        } else {
            t[0] = h_sbox7(t[0]);
        }
        u64 lo[12], hi[12];
        for (int i = 0; i < 12; ++i) { lo[i] = t[i] & GL_EPS; hi[i] = t[i] >> 32; }
        h_mds_half(lo);
        h_mds_half(hi);
        for (int o = 0; o < 12; ++o) {                      // lo + hi 2^32 with hi = hh 2^32 + hl: hh 2^64 = hh (2^32 - 1)
            u64 hl = hi[o] & GL_EPS, hh = hi[o] >> 32;
            u64 v = lo[o] + ((hh << 32) - hh);              // < 2^42
            u64 w = v + (hl << 32);
            s[o] = w + ((w < v) ? GL_EPS : 0);
        }
    }
    for (int i = 0; i < 12; ++i) s[i] = gl_canon(s[i]);
}
#if defined(ZKB_HOST_AVX2)
inline void h_poseidon_permute(u64* s) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) h_poseidon_permute_avx2(s);
    else h_poseidon_permute_scalar(s);
}
#else
inline void h_poseidon_permute(u64* s) { h_poseidon_permute_scalar(s); }
#endif
inline void h_hash_no_pad(const u64* v, size_t len, u64 out[4]) {
    u64 s[12] = {0};
    for (size_t off = 0; off < len; off += 8) {
        size_t m = len - off < 8 ? len - off : 8;
        for (size_t i = 0; i < m; ++i) s[i] = v[off + i];
        h_poseidon_permute(s);
    }
    for (int i = 0; i < 4; ++i) out[i] = s[i];
}
inline void h_hash_pad(const u64* v, size_t len, u64 out[4]) {
    std::vector<u64> p(v, v + len);
    p.push_back(1);
    while ((p.size() + 1) % 8 != 0) p.push_back(0);
    p.push_back(1);
    h_hash_no_pad(p.data(), p.size(), out);
}

struct Challenger {
    u64 sponge[12] = {0};
    u64 in_buf[8];
    int in_len = 0;
    u64 out_buf[8];
    int out_len = 0;
    unsigned permutations = 0;      // host permutations so far (bench reports them: they sit on the proof's critical path)
    double seconds = 0;
    void duplex() {
        for (int i = 0; i < in_len; ++i) sponge[i] = in_buf[i];
        in_len = 0;
        const auto t0 = std::chrono::steady_clock::now();
        h_poseidon_permute(sponge);
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ++permutations;
        for (int i = 0; i < 8; ++i) out_buf[i] = sponge[i];
        out_len = 8;
    }
    void observe(u64 x) {
        out_len = 0;
        in_buf[in_len++] = x;
        if (in_len == 8) duplex();
    }
    void observe_many(const u64* v, size_t n) { for (size_t i = 0; i < n; ++i) observe(v[i]); }
    void observe_ext(ext2 e) { observe(e.a); observe(e.b); }
    u64 get() {
        if (in_len > 0 || out_len == 0) duplex();
        return out_buf[--out_len];
    }
    ext2 get_ext() { u64 a = get(); u64 b = get(); return e_make(a, b); }
};

}  // namespace zkb
