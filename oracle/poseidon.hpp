// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// CPU restatement of the width-12 Goldilocks Poseidon permutation, the overwrite-mode sponge
// (hash_no_pad / hash_or_noop / two_to_one / hash_pad), the Merkle tree with cap and the duplex
// Fiat-Shamir challenger of qp-plonky2 1.1.1 (pinned in /root/reference/Cargo.lock:489-492).
// Reference call sites that pin the behaviour:
//   /root/reference/wormhole/circuit/src/unspendable_account.rs:54-56 (hash_no_pad twice)
//   /root/reference/wormhole/circuit/src/nullifier.rs:64-65
//   /root/reference/wormhole/tests/src/circuit/unspendable_account_tests.rs:12-27 (5 KATs)
//   /root/reference/wormhole/tests/src/prover/prover_tests.rs:31-41 (nullifier KAT)
//   /root/reference/wormhole/tests/test-helpers/src/lib.rs:68-80 (7-node storage-proof chain)
// Algorithm statement: SURVEY.md Appendix A.2-A.4. The permutation uses the *naive* round form
// (add constants, S-box, dense MDS) — deliberately different from the product's kernels.
#pragma once
#include "goldilocks.hpp"
#include <cstring>

namespace orc {

constexpr int SPONGE_WIDTH = 12;
constexpr int SPONGE_RATE = 8;
constexpr int HALF_N_FULL_ROUNDS = 4;
constexpr int N_PARTIAL_ROUNDS = 22;
constexpr int N_ROUNDS = 2 * HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS;
constexpr u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
constexpr u64 MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// The 360 round constants, regenerated from ChaCha8(seed_from_u64(0)) on first use (A.2).
const u64* poseidon_round_constants();

template <class Ops>
inline typename Ops::T sbox7(typename Ops::T x) {
    auto x2 = Ops::mul(x, x);
    auto x4 = Ops::mul(x2, x2);
    auto x3 = Ops::mul(x2, x);
    return Ops::mul(x3, x4);
}

template <class Ops>
inline void mds_layer(typename Ops::T* st) {
    typename Ops::T out[12];
    for (int r = 0; r < 12; ++r) {
        auto acc = Ops::zero();
        for (int i = 0; i < 12; ++i) acc = Ops::add(acc, Ops::mulc(st[(i + r) % 12], MDS_CIRC[i]));
        acc = Ops::add(acc, Ops::mulc(st[r], MDS_DIAG[r]));
        out[r] = acc;
    }
    for (int r = 0; r < 12; ++r) st[r] = out[r];
}

// Base-field MDS layer with one reduction per output (the 13 products of a 64-bit state word and
// a constant <= 41 sum to < 2^74, which fits a u128 accumulator).
template <>
inline void mds_layer<BaseOps>(u64* st) {
    u64 out[12];
    for (int r = 0; r < 12; ++r) {
        u128 acc = (u128)st[r] * MDS_DIAG[r];
        for (int i = 0; i < 12; ++i) acc += (u128)st[(i + r) % 12] * MDS_CIRC[i];
        out[r] = reduce128(acc);
    }
    for (int r = 0; r < 12; ++r) st[r] = out[r];
}

void poseidon_permute_fast(u64* st);      // poseidon_fast.hpp: the CPU-baseline arm's optimised form of the same function
extern bool g_fast_poseidon;

inline void poseidon_permute_naive(u64* st) {
    const u64* rc = poseidon_round_constants();
    for (int r = 0; r < N_ROUNDS; ++r) {
        for (int i = 0; i < 12; ++i) st[i] = fadd(st[i], rc[12 * r + i]);
        bool full = r < HALF_N_FULL_ROUNDS || r >= HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS;
        if (full) {
            for (int i = 0; i < 12; ++i) st[i] = sbox7<BaseOps>(st[i]);
        } else {
            st[0] = sbox7<BaseOps>(st[0]);
        }
        mds_layer<BaseOps>(st);
    }
}

inline void poseidon_permute(u64* st) {
    if (g_fast_poseidon) poseidon_permute_fast(st);
    else poseidon_permute_naive(st);
}

using Digest = std::array<u64, 4>;

inline Digest hash_no_pad(const u64* v, size_t len) {
    u64 st[12] = {0};
    for (size_t off = 0; off < len; off += SPONGE_RATE) {
        size_t m = len - off < (size_t)SPONGE_RATE ? len - off : SPONGE_RATE;
        for (size_t i = 0; i < m; ++i) st[i] = v[off + i];   // overwrite mode
        poseidon_permute(st);
    }
    return {st[0], st[1], st[2], st[3]};
}
inline Digest hash_no_pad(const std::vector<u64>& v) { return hash_no_pad(v.data(), v.size()); }

inline Digest hash_or_noop(const u64* v, size_t len) {
    if (len <= 4) {
        Digest d = {0, 0, 0, 0};
        for (size_t i = 0; i < len; ++i) d[i] = v[i];
        return d;
    }
    return hash_no_pad(v, len);
}

inline Digest two_to_one(const Digest& l, const Digest& r) {
    u64 st[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    poseidon_permute(st);
    return {st[0], st[1], st[2], st[3]};
}

inline Digest hash_pad(const std::vector<u64>& v) {
    std::vector<u64> p(v);
    p.push_back(1);
    while ((p.size() + 1) % SPONGE_RATE != 0) p.push_back(0);
    p.push_back(1);
    return hash_no_pad(p);
}

// ---- Merkle tree with cap (A.3) ----
struct MerkleTree {
    size_t num_leaves = 0;
    size_t leaf_width = 0;
    unsigned cap_height = 0;
    std::vector<u64> leaves;                 // row-major [num_leaves][leaf_width]
    std::vector<std::vector<Digest>> levels; // levels[0] = leaf digests, levels[k] has num_leaves>>k
    std::vector<Digest> cap;

    const u64* leaf(size_t i) const { return leaves.data() + i * leaf_width; }
    // siblings bottom-up, length log2(num_leaves) - cap_height
    std::vector<Digest> prove(size_t i) const {
        std::vector<Digest> path;
        unsigned lg = log2_strict(num_leaves);
        for (unsigned k = 0; k + cap_height < lg; ++k) {
            path.push_back(levels[k][i ^ 1]);
            i >>= 1;
        }
        return path;
    }
};
MerkleTree merkle_build(std::vector<u64> leaves, size_t num_leaves, size_t leaf_width, unsigned cap_height);
bool merkle_verify(const u64* leaf, size_t leaf_width, size_t index, const std::vector<Digest>& cap,
                   const std::vector<Digest>& path);

// ---- Challenger (A.4) ----
struct Challenger {
    u64 sponge[12] = {0};
    std::vector<u64> in_buf, out_buf;
    void duplex() {
        for (size_t i = 0; i < in_buf.size(); ++i) sponge[i] = in_buf[i];
        in_buf.clear();
        poseidon_permute(sponge);
        out_buf.assign(sponge, sponge + SPONGE_RATE);
    }
    void observe(u64 x) {
        out_buf.clear();
        in_buf.push_back(x);
        if (in_buf.size() == (size_t)SPONGE_RATE) duplex();
    }
    void observe_digest(const Digest& d) { for (u64 x : d) observe(x); }
    void observe_cap(const std::vector<Digest>& cap) { for (auto& d : cap) observe_digest(d); }
    void observe_ext(E2 e) { observe(e.a); observe(e.b); }
    u64 get() {
        if (!in_buf.empty() || out_buf.empty()) duplex();
        u64 x = out_buf.back();
        out_buf.pop_back();
        return x;
    }
    E2 get_ext() { u64 a = get(); u64 b = get(); return E2(a, b); }
};

}  // namespace orc
