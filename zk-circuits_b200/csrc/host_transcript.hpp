// Host-side Poseidon sponge + duplex challenger for the Fiat-Shamir transcript
// (qp-plonky2 1.1.1 `iop::challenger::Challenger<F, PoseidonHash>`). The transcript is
// strictly serial (~100 permutations per proof), so it runs on the host between kernel launches.
#pragma once
#include <vector>
#include "field.cuh"
#include "kernels.h"

namespace zkb {

inline u64 h_sbox7(u64 x) {
    u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}
inline void h_poseidon_permute(u64* s) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const u64* rc = host_round_constants();
    for (int r = 0; r < 30; ++r) {
        for (int i = 0; i < 12; ++i) s[i] = gl_add(s[i], rc[12 * r + i]);
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; ++i) s[i] = h_sbox7(s[i]);
        } else {
            s[0] = h_sbox7(s[0]);
        }
        u64 out[12];
        for (int o = 0; o < 12; ++o) {
            unsigned __int128 acc = o == 0 ? (unsigned __int128)s[0] * 8 : 0;
            for (int i = 0; i < 12; ++i) acc += (unsigned __int128)s[(i + o) % 12] * C[i];
            out[o] = gl_canon(gl_reduce128_lazy((u64)acc, (u64)(acc >> 64)));
        }
        for (int o = 0; o < 12; ++o) s[o] = out[o];
    }
}
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
    void duplex() {
        for (int i = 0; i < in_len; ++i) sponge[i] = in_buf[i];
        in_len = 0;
        h_poseidon_permute(sponge);
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
