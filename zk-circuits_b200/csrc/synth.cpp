// Synthetic wormhole-/voting-shaped circuit + witness generator (host code, part of the product's
// tooling): stands in for the Rust side of the boundary — `CircuitBuilder::build_prover`
// (/root/reference/wormhole/circuit/src/circuit.rs:98-108) and witness generation
// (/root/reference/wormhole/prover/src/lib.rs:209-225) — which cannot run without a Rust toolchain.
// It emits what zkb_circuit_create()/zkb_prove() consume: CommonCircuitData bytes, constants/sigma
// values, the wires matrix and public inputs, for the reference circuit's configuration, gate set,
// selector grouping (decoded from wormhole/bench-data/common.bin) and row mix
// (SURVEY.md App. C.1), including the zk blinding rows of upstream `CircuitBuilder::blind`.
// Not on the proving hot path; used by bench.py, smoke() and the tests to obtain workloads.
#include "synth.hpp"
#include "host_transcript.hpp"
#include <algorithm>
#include <numeric>
#include <stdexcept>

namespace zkb {

namespace {
inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
struct Rng {
    u64 s;
    u64 next() { s += 0x9e3779b97f4a7c15ULL; return mix64(s); }
    u64 felt() { return gl_canon(next()); }
    u64 below(u64 m) { return next() % m; }
};
// row kinds; a circuit's gate table maps them to gate indices
enum Kind { G_NOOP = 0, G_CONST, G_PI, G_BASESUM, G_ARITH, G_POSEIDON,
            G_ARITH_EXT, G_MUL_EXT, G_REDUCING, G_REDUCING_EXT, G_RANDOM_ACCESS, G_EXP, G_COSET, G_MDS, NUM_KINDS };
// one entry of the serialized gate list: DefaultGateSerializer tag, parameters, constraint degree
struct GateRow { Kind kind; u32 tag; std::vector<u64> params; unsigned degree; };
// sorted by (degree, id) as upstream's CircuitBuilder::build leaves them
std::vector<GateRow> gate_table(bool recursion, const std::vector<u64>& bary_weights) {
    if (!recursion)
        return {{G_NOOP, 9, {}, 0}, {G_CONST, 3, {2}, 1}, {G_PI, 12, {}, 1}, {G_BASESUM, 2, {63}, 2}, {G_ARITH, 0, {20}, 3},
                {G_POSEIDON, 11, {}, 7}};
    std::vector<u64> coset = {4, 6, 16};      // subgroup_bits, degree, weights.len, weights...
    coset.insert(coset.end(), bary_weights.begin(), bary_weights.end());
    return {{G_NOOP, 9, {}, 0}, {G_CONST, 3, {2}, 1}, {G_MDS, 10, {}, 1}, {G_PI, 12, {}, 1}, {G_BASESUM, 2, {63}, 2},
            {G_REDUCING_EXT, 14, {32}, 2}, {G_REDUCING, 15, {43}, 2}, {G_ARITH_EXT, 1, {10}, 3}, {G_ARITH, 0, {20}, 3},
            {G_MUL_EXT, 8, {13}, 3}, {G_EXP, 5, {66}, 4}, {G_RANDOM_ACCESS, 13, {4, 4, 2}, 5}, {G_COSET, 4, coset, 6},
            {G_POSEIDON, 11, {}, 7}};
}
constexpr int W_SWAP = 24, W_DELTA = 25, W_FULL0 = 29, W_PARTIAL = 65, W_FULL1 = 87;
constexpr u64 UNUSED_SEL = 0xFFFFFFFFULL;
constexpr u64 NUM_QUERY_ROUNDS = 28, RATE_BITS = 3, CAP_HEIGHT = 4, ARITY_BITS = 4, FINAL_POLY_BITS = 5;

struct Row {
    int gate = G_NOOP;
    u64 consts[2] = {0, 0};
    std::vector<u64> w;
    Row() : w(135, 0) {}
};
struct Cell { u32 row, col; };
struct Pooled { u64 value; Cell cell; };

void h_mds_layer(u64* s) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 o[12];
    for (int r = 0; r < 12; ++r) {
        unsigned __int128 acc = r == 0 ? (unsigned __int128)s[0] * 8 : 0;
        for (int i = 0; i < 12; ++i) acc += (unsigned __int128)s[(i + r) % 12] * C[i];
        o[r] = gl_canon(gl_reduce128_lazy((u64)acc, (u64)(acc >> 64)));
    }
    for (int r = 0; r < 12; ++r) s[r] = o[r];
}
// FriReductionStrategy::ConstantArityBits(4, 5)
std::vector<u64> arity_schedule(u64 degree_bits) {
    std::vector<u64> v;
    while (degree_bits > FINAL_POLY_BITS && degree_bits + RATE_BITS - ARITY_BITS >= CAP_HEIGHT) {
        v.push_back(ARITY_BITS);
        degree_bits -= ARITY_BITS;
    }
    return v;
}
struct Bytes {
    std::vector<uint8_t> b;
    void w8(uint8_t v) { b.push_back(v); }
    void w32(u32 v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void w64(u64 v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void vec(const std::vector<u64>& v) { w64(v.size()); for (u64 x : v) w64(x); }
};
}  // namespace

// CommonCircuitData::to_bytes layout
static std::vector<uint8_t> serialize_common(const SynthCircuit& s, const std::vector<u64>& k_is, bool zk,
                                             const std::vector<GateRow>& gates, const std::vector<u64>& selector_indices,
                                             const std::vector<std::pair<u64, u64>>& groups, u64 num_gate_constraints) {
    Bytes o;
    o.w64(135); o.w64(80); o.w64(2); o.w64(100); o.w64(2); o.w64(8);
    o.w8(1); o.w8(zk ? 1 : 0);
    for (int rep = 0; rep < 2; ++rep) {     // FriConfig, then again inside FriParams
        o.w64(RATE_BITS); o.w64(CAP_HEIGHT); o.w64(NUM_QUERY_ROUNDS); o.w32(16);
        o.w8(1); o.w64(ARITY_BITS); o.w64(FINAL_POLY_BITS);
    }
    o.vec(s.reduction_arity_bits);
    o.w64(s.degree_bits);
    o.w8(zk ? 1 : 0);                       // hiding
    o.vec(selector_indices);
    o.w64(groups.size());
    for (auto& g : groups) { o.w64(g.first); o.w64(g.second); }
    o.w64(8); o.w64(num_gate_constraints); o.w64(groups.size() + 2); o.w64(s.public_inputs.size());
    o.vec(k_is);
    o.w64(9); o.w64(0); o.w64(0); o.w64(0);
    o.w64(gates.size());
    for (auto& g : gates) {
        o.w32(g.tag);
        for (u64 p : g.params) o.w64(p);
    }
    return o.b;
}

SynthCircuit make_synth_circuit(const SynthSpec& spec) {
    Rng rng{spec.seed * 0x243f6a8885a308d3ULL + 0x13198a2e03707344ULL};
    SynthCircuit out;
    std::vector<u64> k_is(80);
    k_is[0] = 1;
    for (int j = 1; j < 80; ++j) k_is[j] = gl_mul(k_is[j - 1], GL_GEN);

    // gate table, selector groups by upstream's greedy rule (extend a group while size + degree < 9)
    const bool recursion = spec.recursion();
    std::vector<u64> bary(16);
    {
        u64 xs[16], w16 = gl_root_of_unity(4);
        xs[0] = 1;
        for (int i = 1; i < 16; ++i) xs[i] = gl_mul(xs[i - 1], w16);
        for (int i = 0; i < 16; ++i) {
            u64 d = 1;
            for (int j = 0; j < 16; ++j) if (j != i) d = gl_mul(d, gl_sub(xs[i], xs[j]));
            bary[i] = gl_inv(d);
        }
    }
    const std::vector<GateRow> gates = gate_table(recursion, bary);
    int gate_index[NUM_KINDS];
    std::fill(gate_index, gate_index + NUM_KINDS, -1);
    for (size_t i = 0; i < gates.size(); ++i) gate_index[gates[i].kind] = (int)i;
    std::vector<u64> selector_indices(gates.size());
    std::vector<std::pair<u64, u64>> groups;
    for (size_t start = 0; start < gates.size();) {
        size_t size = 0;
        while (start + size < gates.size() && size + gates[start + size].degree < 9) ++size;
        for (size_t i = start; i < start + size; ++i) selector_indices[i] = groups.size();
        groups.push_back({start, start + size});
        start += size;
    }
    const size_t nsel = groups.size();

    std::vector<Row> rows;
    std::vector<std::pair<Cell, Cell>> copies;
    std::vector<Pooled> pool, bools;
    auto connect = [&](Cell a, Cell b) { copies.push_back({a, b}); };
    auto new_row = [&](int gate) { rows.emplace_back(); rows.back().gate = gate; return (u32)(rows.size() - 1); };
    // routed input: copy from the pool (3/4) or fresh
    auto take_input = [&](u32 row, u32 col) {
        u64 v;
        if (!pool.empty() && rng.below(4) != 0) {
            const Pooled& p = pool[rng.below(pool.size())];
            v = p.value;
            connect({row, col}, p.cell);
        } else {
            v = rng.felt();
        }
        rows[row].w[col] = v;
        return v;
    };
    auto take_ext = [&](u32 row, u32 col) {     // an F_{p^2} value on two consecutive routed wires
        u64 a = take_input(row, col);
        u64 b = take_input(row, col + 1);
        return e_make(a, b);
    };
    auto put_ext = [&](u32 row, u32 col, ext2 v, bool to_pool) {
        rows[row].w[col] = v.a;
        rows[row].w[col + 1] = v.b;
        if (to_pool) { pool.push_back({v.a, {row, col}}); pool.push_back({v.b, {row, col + 1}}); }
    };
    auto copy_from = [&](u32 row, u32 col, const Pooled& p) {
        rows[row].w[col] = p.value;
        connect({row, col}, p.cell);
    };

    // row 0: public-input gate; row 1: constants 0 and 1
    u32 pi_row = new_row(G_PI);
    u32 c01 = new_row(G_CONST);
    rows[c01].consts[0] = 0; rows[c01].consts[1] = 1;
    rows[c01].w[0] = 0; rows[c01].w[1] = 1;
    Pooled ZERO{0, {c01, 0}}, ONE{1, {c01, 1}};
    bools.push_back(ZERO);
    bools.push_back(ONE);

    const u64* rc = host_round_constants();
    // Fills a Poseidon row whose input wires 0..11 and swap wire are already set; returns nothing.
    auto fill_poseidon = [&](u32 r) {
        std::vector<u64>& w = rows[r].w;
        u64 swap = w[W_SWAP];
        u64 st[12];
        for (int i = 0; i < 4; ++i) {
            u64 d = gl_mul(swap, gl_sub(w[i + 4], w[i]));
            w[W_DELTA + i] = d;
            st[i] = gl_add(w[i], d);
            st[i + 4] = gl_sub(w[i + 4], d);
        }
        for (int i = 8; i < 12; ++i) st[i] = w[i];
        int round = 0;
        for (int k = 0; k < 4; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = gl_add(st[i], rc[12 * round + i]);
            if (k != 0) for (int i = 0; i < 12; ++i) w[W_FULL0 + 12 * (k - 1) + i] = st[i];
            for (int i = 0; i < 12; ++i) st[i] = h_sbox7(st[i]);
            h_mds_layer(st);
        }
        for (int k = 0; k < 22; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = gl_add(st[i], rc[12 * round + i]);
            w[W_PARTIAL + k] = st[0];
            st[0] = h_sbox7(st[0]);
            h_mds_layer(st);
        }
        for (int k = 0; k < 4; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = gl_add(st[i], rc[12 * round + i]);
            for (int i = 0; i < 12; ++i) w[W_FULL1 + 12 * k + i] = st[i];
            for (int i = 0; i < 12; ++i) st[i] = h_sbox7(st[i]);
            h_mds_layer(st);
        }
        for (int i = 0; i < 12; ++i) w[12 + i] = st[i];
    };

    // public inputs hashed in-circuit by a sponge of Poseidon rows, digest wired to the PI gate
    out.public_inputs.resize(spec.num_public_inputs);
    for (auto& p : out.public_inputs) p = (rng.below(3) == 0) ? rng.felt() : rng.below(u64(1) << 32);
    size_t poseidon_used = 0;
    {
        size_t npi = spec.num_public_inputs;
        size_t nchunks = (npi + 7) / 8;      // no public inputs: hash_no_pad([]) is the zero digest, no permutation
        u32 prev = 0;
        u64 state[12] = {0};
        for (size_t k = 0; k < nchunks; ++k) {
            u32 r = new_row(G_POSEIDON);
            ++poseidon_used;
            size_t off = 8 * k;
            size_t m = npi > off ? std::min<size_t>(8, npi - off) : 0;
            for (size_t i = 0; i < 12; ++i) {
                if (i < m) rows[r].w[i] = out.public_inputs[off + i];
                else if (k == 0) copy_from(r, (u32)i, ZERO);
                else copy_from(r, (u32)i, Pooled{state[i], {prev, (u32)(12 + i)}});
            }
            copy_from(r, W_SWAP, ZERO);
            fill_poseidon(r);
            for (int i = 0; i < 12; ++i) state[i] = rows[r].w[12 + i];
            prev = r;
        }
        u64 h[4]; h_hash_no_pad(out.public_inputs.data(), out.public_inputs.size(), h);
        for (int i = 0; i < 4; ++i) {
            if (state[i] != h[i]) throw std::runtime_error("synthetic PI sponge mismatch");
            if (nchunks == 0) copy_from(pi_row, (u32)i, ZERO);
            else copy_from(pi_row, (u32)i, Pooled{state[i], {prev, (u32)(12 + i)}});
        }
    }

    // remaining rows in a deterministic shuffle of gate types
    std::vector<int> todo;
    for (size_t i = 1; i < spec.n_const; ++i) todo.push_back(G_CONST);
    for (size_t i = 0; i < spec.n_base_sum; ++i) todo.push_back(G_BASESUM);
    for (size_t i = 0; i < spec.n_arith; ++i) todo.push_back(G_ARITH);
    {
        const std::pair<Kind, size_t> extra[] = {{G_ARITH_EXT, spec.n_arith_ext}, {G_MUL_EXT, spec.n_mul_ext}, {G_REDUCING, spec.n_reducing},
                                                 {G_REDUCING_EXT, spec.n_reducing_ext}, {G_RANDOM_ACCESS, spec.n_random_access},
                                                 {G_EXP, spec.n_exp}, {G_COSET, spec.n_coset}, {G_MDS, spec.n_mds}};
        for (auto& e : extra) todo.insert(todo.end(), e.second, (int)e.first);
    }
    // Poseidon rows come in sponge chains (storage-proof-like, 24 rows) and single compressions
    size_t pos_left = spec.n_poseidon > poseidon_used ? spec.n_poseidon - poseidon_used : 0;
    const int CHAIN = -1;
    while (pos_left > 0) {
        if (pos_left >= 24 && rng.below(4) != 0) { todo.push_back(CHAIN); pos_left -= 24; }
        else { todo.push_back(G_POSEIDON); pos_left -= 1; }
    }
    for (size_t i = todo.size(); i > 1; --i) std::swap(todo[i - 1], todo[rng.below(i)]);

    std::vector<Pooled> sums;
    for (int t : todo) {
        if (t == G_CONST) {
            u32 r = new_row(G_CONST);
            for (int k = 0; k < 2; ++k) {
                u64 v = rng.below(2) ? rng.felt() : rng.below(256);
                rows[r].consts[k] = v;
                rows[r].w[k] = v;
                pool.push_back({v, {r, (u32)k}});
            }
        } else if (t == G_BASESUM) {
            u32 r = new_row(G_BASESUM);
            u64 v;
            if (!sums.empty() && rng.below(4) == 0) {
                const Pooled p = sums[rng.below(sums.size())];
                v = p.value;
                copy_from(r, 0, p);
            } else {
                v = rng.below(2) ? rng.below(u64(1) << 32) : (rng.next() >> 1);
                rows[r].w[0] = v;
            }
            for (int k = 0; k < 63; ++k) rows[r].w[1 + k] = (v >> k) & 1;
            sums.push_back({v, {r, 0}});
            if (sums.size() > 64) sums.erase(sums.begin());
            pool.push_back({v, {r, 0}});
            for (int k = 0; k < 3; ++k) {
                u32 b = (u32)rng.below(63);
                bools.push_back({rows[r].w[1 + b], {r, 1 + b}});
            }
        } else if (t == G_ARITH) {
            u32 r = new_row(G_ARITH);
            u64 sel = rng.below(3);
            rows[r].consts[0] = sel == 0 ? 1 : (sel == 1 ? 1 : rng.felt());
            rows[r].consts[1] = sel == 0 ? 1 : (sel == 1 ? 0 : rng.felt());
            for (u32 i = 0; i < 20; ++i) {
                u64 m0 = take_input(r, 4 * i), m1 = take_input(r, 4 * i + 1), ad = take_input(r, 4 * i + 2);
                u64 out = gl_add(gl_mul(rows[r].consts[0], gl_mul(m0, m1)), gl_mul(rows[r].consts[1], ad));
                rows[r].w[4 * i + 3] = out;
                pool.push_back({out, {r, 4 * i + 3}});
            }
        } else if (t == G_ARITH_EXT || t == G_MUL_EXT) {
            // out = c0 a b (+ c1 addend) in F_{p^2}; 10 ops of 8 wires, or 13 ops of 6 wires without the addend
            const u32 stride = t == G_ARITH_EXT ? 8 : 6, ops = t == G_ARITH_EXT ? 10 : 13;
            u32 r = new_row(t);
            rows[r].consts[0] = rng.below(2) ? 1 : rng.felt();
            if (t == G_ARITH_EXT) rows[r].consts[1] = rng.below(2) ? 1 : rng.felt();
            for (u32 i = 0; i < ops; ++i) {
                ext2 a = take_ext(r, stride * i);
                ext2 b = take_ext(r, stride * i + 2);
                ext2 o = e_mul_base(e_mul(a, b), rows[r].consts[0]);
                if (t == G_ARITH_EXT) o = e_add(o, e_mul_base(take_ext(r, stride * i + 4), rows[r].consts[1]));
                put_ext(r, stride * i + stride - 2, o, true);
            }
        } else if (t == G_REDUCING || t == G_REDUCING_EXT) {
            // Horner chain acc <- acc alpha + coeff over 43 base / 32 extension coefficients; the last value lands on wires 0..1
            const u32 nc = t == G_REDUCING ? 43 : 32, cw = t == G_REDUCING ? 1 : 2, accs_at = 6 + cw * nc;
            u32 r = new_row(t);
            ext2 alpha = take_ext(r, 2);
            ext2 acc = take_ext(r, 4);
            for (u32 i = 0; i < nc; ++i) {
                ext2 coeff = cw == 2 ? take_ext(r, 6 + 2 * i) : e_from(take_input(r, 6 + i));
                acc = e_add(e_mul(acc, alpha), coeff);
                if (i + 1 == nc) put_ext(r, 0, acc, true);
                else put_ext(r, accs_at + 2 * i, acc, false);
            }
        } else if (t == G_RANDOM_ACCESS) {
            // four 16-way selections (index bits on the unrouted wires 74..89) and two constants on wires 72, 73
            u32 r = new_row(t);
            for (u32 cpy = 0; cpy < 4; ++cpy) {
                const u32 at = 18 * cpy;
                const u64 idx = rng.below(16);
                rows[r].w[at] = idx;
                for (u32 i = 0; i < 16; ++i) take_input(r, at + 2 + i);
                rows[r].w[at + 1] = rows[r].w[at + 2 + idx];
                pool.push_back({rows[r].w[at + 1], {r, at + 1}});
                for (u32 i = 0; i < 4; ++i) rows[r].w[74 + 4 * cpy + i] = (idx >> i) & 1;
            }
            for (u32 k = 0; k < 2; ++k) {
                const u64 v = rng.below(2) ? rng.felt() : rng.below(256);
                rows[r].consts[k] = v;
                rows[r].w[72 + k] = v;
                pool.push_back({v, {r, 72 + k}});
            }
        } else if (t == G_EXP) {
            // square-and-multiply over 66 exponent bits (most significant first), every partial power on its own wire
            u32 r = new_row(t);
            const u64 base = take_input(r, 0);
            for (u32 i = 0; i < 66; ++i) rows[r].w[1 + i] = rng.below(2);
            u64 power = 1;
            for (u32 i = 0; i < 66; ++i) {
                if (i) power = gl_mul(power, power);
                if (rows[r].w[66 - i]) power = gl_mul(power, base);
                rows[r].w[68 + i] = power;
            }
            rows[r].w[67] = power;
            pool.push_back({power, {r, 67}});
            for (u32 k = 0; k < 2; ++k) { u32 b = (u32)rng.below(66); bools.push_back({rows[r].w[1 + b], {r, 1 + b}}); }
        } else if (t == G_COSET) {
            // 16 F_{p^2} values on shift * <w_16>, interpolated at an F_{p^2} point: barycentric sums in three runs of
            // 6 + 5 + 5 points, the two intermediate (sum, product) pairs on wires 37..44, point / shift on 45..46
            u32 r = new_row(t);
            u64 shift;
            do shift = rng.felt(); while (shift == 0);
            rows[r].w[0] = shift;
            for (u32 i = 0; i < 32; ++i) take_input(r, 1 + i);
            const ext2 z = e_mul_base(take_ext(r, 33), gl_inv(shift));
            put_ext(r, 45, z, false);
            ext2 sum = e_from(0), prod = e_from(1);
            u64 x = 1;
            const u64 w16 = gl_root_of_unity(4);
            for (u32 k = 0; k < 16; ++k, x = gl_mul(x, w16)) {
                if (k == 6 || k == 11) {
                    const u32 i = k == 6 ? 0 : 1;
                    put_ext(r, 37 + 2 * i, sum, false);
                    put_ext(r, 41 + 2 * i, prod, false);
                }
                const ext2 term = e_sub(z, e_from(x));
                const ext2 wv = e_mul_base(e_make(rows[r].w[1 + 2 * k], rows[r].w[2 + 2 * k]), bary[k]);
                sum = e_add(e_mul(sum, term), e_mul(wv, prod));
                prod = e_mul(prod, term);
            }
            put_ext(r, 35, sum, true);
        } else if (t == G_MDS) {
            // Poseidon's MDS matrix on 12 F_{p^2} values: it acts on the two components separately
            u32 r = new_row(t);
            u64 lo[12], hi[12];
            for (u32 i = 0; i < 12; ++i) { ext2 v = take_ext(r, 2 * i); lo[i] = v.a; hi[i] = v.b; }
            h_mds_layer(lo);
            h_mds_layer(hi);
            for (u32 k = 0; k < 12; ++k) put_ext(r, 24 + 2 * k, e_make(lo[k], hi[k]), true);
        } else if (t == G_POSEIDON) {
            u32 r = new_row(G_POSEIDON);
            for (u32 i = 0; i < 8; ++i) take_input(r, i);
            for (u32 i = 8; i < 12; ++i) copy_from(r, i, ZERO);
            copy_from(r, W_SWAP, bools[rng.below(bools.size())]);
            fill_poseidon(r);
            for (u32 i = 0; i < 4; ++i) pool.push_back({rows[r].w[12 + i], {r, 12 + i}});
        } else {  // sponge chain of 24 permutations
            u32 prev = 0;
            for (int k = 0; k < 24; ++k) {
                u32 r = new_row(G_POSEIDON);
                for (u32 i = 0; i < 8; ++i) take_input(r, i);
                for (u32 i = 8; i < 12; ++i) {
                    if (k == 0) copy_from(r, i, ZERO);
                    else copy_from(r, i, Pooled{rows[prev].w[12 + i], {prev, 12 + i}});
                }
                copy_from(r, W_SWAP, ZERO);
                fill_poseidon(r);
                prev = r;
            }
            for (u32 i = 0; i < 4; ++i) pool.push_back({rows[prev].w[12 + i], {prev, 12 + i}});
        }
        if (pool.size() > 4096) pool.erase(pool.begin(), pool.begin() + 1024);
    }

    // zk blinding rows (upstream CircuitBuilder::blind / blinding_counts)
    if (spec.zk) {
        size_t num_gates = rows.size();
        unsigned est_bits = 0;
        while ((size_t(1) << est_bits) < num_gates) ++est_bits;
        size_t regular = 0, zop = 0;
        for (;; ++est_bits) {
            auto arities = arity_schedule(est_bits);
            u64 fold = 0, asum = 0;
            for (u64 a : arities) { fold += (u64(1) << a) - 1; asum += a; }
            u64 final_coeffs = (u64(1) << est_bits) >> asum;
            u64 fri_open = NUM_QUERY_ROUNDS * (1 + 2 * fold + 2 * final_coeffs);
            regular = 2 + fri_open;
            zop = 4 + fri_open;
            if (num_gates + regular + 2 * zop <= (size_t(1) << est_bits)) break;
        }
        for (size_t i = 0; i < regular; ++i) {
            u32 r = new_row(G_NOOP);
            for (auto& x : rows[r].w) x = rng.felt();
        }
        for (size_t i = 0; i < zop; ++i) {
            u32 r1 = new_row(G_NOOP), r2 = new_row(G_NOOP);
            for (u32 j = 0; j < 80; ++j) {
                u64 v = rng.felt();
                rows[r1].w[j] = v;
                rows[r2].w[j] = v;
                connect({r1, j}, {r2, j});
            }
        }
    }
    unsigned db = 0;
    while ((size_t(1) << db) < rows.size() || db < spec.min_degree_bits) ++db;
    while (rows.size() < (size_t(1) << db)) new_row(G_NOOP);
    size_t n = rows.size();
    out.degree_bits = db;
    out.reduction_arity_bits = arity_schedule(db);

    // wires, constants
    out.wires.assign(135, std::vector<u64>(n));
    out.const_sigma_values.assign(nsel + 2 + 80, std::vector<u64>(n));
    for (size_t r = 0; r < n; ++r) {
        for (int j = 0; j < 135; ++j) out.wires[j][r] = rows[r].w[j];
        const int g = gate_index[rows[r].gate];
        if (g < 0) throw std::runtime_error("row kind outside the circuit's gate set");
        for (size_t k = 0; k < nsel; ++k) out.const_sigma_values[k][r] = selector_indices[g] == k ? (u64)g : UNUSED_SEL;
        out.const_sigma_values[nsel][r] = rows[r].consts[0];
        out.const_sigma_values[nsel + 1][r] = rows[r].consts[1];
    }
    // sigma from the copy-constraint classes (cycle through each class)
    std::vector<u32> parent(80 * n);
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto id = [&](Cell cl) { return (u32)(cl.col * n + cl.row); };
    for (auto& cp : copies) {
        u32 a = find(id(cp.first)), b = find(id(cp.second));
        if (a != b) parent[a] = b;
    }
    std::vector<u32> order(80 * n);
    std::iota(order.begin(), order.end(), 0u);
    std::vector<u32> rootv(80 * n);
    for (u32 i = 0; i < 80 * n; ++i) rootv[i] = find(i);
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return rootv[a] < rootv[b]; });
    std::vector<u32> sigma(80 * n);
    for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && rootv[order[j]] == rootv[order[i]]) ++j;
        for (size_t k = i; k < j; ++k) sigma[order[k]] = order[k + 1 < j ? k + 1 : i];
        i = j;
    }
    u64 w = gl_root_of_unity(db);
    std::vector<u64> subgroup(n);
    subgroup[0] = 1;
    for (size_t i = 1; i < n; ++i) subgroup[i] = gl_mul(subgroup[i - 1], w);
    for (u32 col = 0; col < 80; ++col)
        for (size_t r = 0; r < n; ++r) {
            u32 t = sigma[col * n + r];
            out.const_sigma_values[nsel + 2 + col][r] = gl_mul(k_is[t / n], subgroup[t % n]);
        }
    out.common = serialize_common(out, k_is, spec.zk, gates, selector_indices, groups, 123);
    return out;
}


}  // namespace zkb
