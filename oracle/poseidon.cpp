// ORACLE — TEST INFRASTRUCTURE ONLY (see poseidon.hpp header).
#include "poseidon.hpp"
#include "poseidon_fast.hpp"
#include "ntt.hpp"
#include <map>
#include <mutex>

namespace orc {

bool g_fast_poseidon = false;
void poseidon_permute_fast(u64* st) { poseidon_permute_fast_impl(st); }
const std::vector<u64>& fastntt::twiddles(unsigned lg) {
    static std::mutex mu;
    static std::map<unsigned, std::vector<u64>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(lg);
    if (it != cache.end()) return it->second;
    std::vector<u64> t(lg ? (size_t(1) << (lg - 1)) : 1);
    const u64 w = root_of_unity(lg);
    t[0] = 1;
    for (size_t k = 1; k < t.size(); ++k) t[k] = fmul(t[k - 1], w);
    return cache.emplace(lg, std::move(t)).first->second;
}

// ---- round-constant regeneration: ChaCha8 keystream keyed by rand_core's
// seed_from_u64(0) PCG32 expansion; u64 draws mapped to [0,p) by the widening-multiply rule.
static inline u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
static void chacha8_block(const u32 key[8], u64 counter, u32 out[16]) {
    static const u32 sigma[4] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    u32 s[16];
    for (int i = 0; i < 4; ++i) s[i] = sigma[i];
    for (int i = 0; i < 8; ++i) s[4 + i] = key[i];
    s[12] = (u32)counter;
    s[13] = (u32)(counter >> 32);
    s[14] = 0;
    s[15] = 0;
    u32 x[16];
    for (int i = 0; i < 16; ++i) x[i] = s[i];
#define QR(a, b, c, d)                                   \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 16); \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 12); \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 8);  \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 7);
    for (int r = 0; r < 4; ++r) {  // 8 rounds = 4 double rounds
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

static u64 g_rc[12 * N_ROUNDS];
static std::once_flag g_rc_once;

static void gen_round_constants() {
    // rand_core::SeedableRng::seed_from_u64(0): PCG32 fills the 32-byte seed.
    u64 state = 0;
    u32 key[8];
    for (int i = 0; i < 8; ++i) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        u32 xs = (u32)(((state >> 18) ^ state) >> 27);
        u32 rot = (u32)(state >> 59);
        key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    u64 counter = 0;
    u32 blk[16];
    int pos = 16;
    auto next_u32 = [&]() {
        if (pos == 16) { chacha8_block(key, counter++, blk); pos = 0; }
        return blk[pos++];
    };
    int n = 0;
    while (n < 12 * N_ROUNDS) {
        u64 lo = next_u32();
        u64 hi = next_u32();
        u64 v = lo | (hi << 32);
        u128 m = (u128)v * P;
        if ((u64)m <= P - 1) g_rc[n++] = (u64)(m >> 64);
    }
}

const u64* poseidon_round_constants() {
    std::call_once(g_rc_once, gen_round_constants);
    return g_rc;
}

// ---- Merkle ----
MerkleTree merkle_build(std::vector<u64> leaves, size_t num_leaves, size_t leaf_width, unsigned cap_height) {
    MerkleTree t;
    t.num_leaves = num_leaves;
    t.leaf_width = leaf_width;
    t.cap_height = cap_height;
    t.leaves = std::move(leaves);
    unsigned lg = log2_strict(num_leaves);
    if (cap_height > lg) throw std::runtime_error("cap_height > log2(num_leaves)");
    t.levels.resize(lg - cap_height + 1);
    t.levels[0].resize(num_leaves);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)num_leaves; ++i) t.levels[0][i] = hash_or_noop(t.leaf(i), leaf_width);
    for (unsigned k = 1; k <= lg - cap_height; ++k) {
        size_t m = num_leaves >> k;
        t.levels[k].resize(m);
#pragma omp parallel for schedule(static) if (m >= 64)
        for (long i = 0; i < (long)m; ++i) t.levels[k][i] = two_to_one(t.levels[k - 1][2 * i], t.levels[k - 1][2 * i + 1]);
    }
    t.cap = t.levels[lg - cap_height];
    return t;
}

bool merkle_verify(const u64* leaf, size_t leaf_width, size_t index, const std::vector<Digest>& cap,
                   const std::vector<Digest>& path) {
    Digest cur = hash_or_noop(leaf, leaf_width);
    for (auto& sib : path) {
        size_t bit = index & 1;
        index >>= 1;
        cur = bit ? two_to_one(sib, cur) : two_to_one(cur, sib);
    }
    if (index >= cap.size()) return false;
    return cur == cap[index];
}

}  // namespace orc
