// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// Synthetic "wormhole-shaped" / "voting-shaped" circuit + witness generator (SURVEY.md §7.2 step 5,
// §8d: no real witness is obtainable without the Rust toolchain). It plays the role that
// `CircuitBuilder::build_prover` + witness generation play in the reference
// (/root/reference/wormhole/circuit/src/circuit.rs:98-108, /root/reference/wormhole/prover/src/lib.rs:209-225):
// it emits the inputs the prove boundary receives — CommonCircuitData, the constants/sigma
// polynomials, the full wires matrix and the public inputs — for a circuit with the same
// configuration, gate set, selector grouping and row mix as the reference's wormhole circuit
// (gate set and selector layout decoded from wormhole/bench-data/common.bin, SURVEY.md B.1;
// row mix from SURVEY.md App. C.1; zk blinding rows per upstream `CircuitBuilder::blind`).
#include "prover.hpp"
#include "gates.hpp"
#include <numeric>
#include <algorithm>

namespace orc {

namespace {
struct Rng {
    u64 s;
    u64 next() { s += 0x9e3779b97f4a7c15ULL; return splitmix64_finalize(s); }
    u64 felt() { return from_u64(next()); }
    u64 below(u64 m) { return next() % m; }
};
// row kinds (indices into the per-circuit gate table are looked up through `gate_index`)
enum Kind { G_NOOP = 0, G_CONST, G_PI, G_BASESUM, G_ARITH, G_POSEIDON,
            G_ARITH_EXT, G_MUL_EXT, G_REDUCING, G_REDUCING_EXT, G_RANDOM_ACCESS, G_EXP, G_COSET, G_MDS, NUM_KINDS };

struct Row {
    int gate = G_NOOP;
    u64 consts[2] = {0, 0};
    std::vector<u64> w;
    Row() : w(135, 0) {}
};
struct Cell { u32 row, col; };
struct Pooled { u64 value; Cell cell; };
}  // namespace

SynthCircuit make_synth_circuit(const SynthSpec& spec) {
    Rng rng{spec.seed * 0x243f6a8885a308d3ULL + 0x13198a2e03707344ULL};
    SynthCircuit sc;
    CommonData& c = sc.common;
    c.num_wires = 135; c.num_routed_wires = 80; c.num_constants_cfg = 2; c.security_bits = 100;
    c.num_challenges = 2; c.max_quotient_degree_factor = 8;
    c.use_base_arithmetic_gate = true; c.zero_knowledge = spec.zk;
    c.fri_config = FriConfig{3, 4, 28, 16, 1, {4, 5}};
    c.hiding = spec.zk;
    // gate table sorted by (degree, id) as upstream's CircuitBuilder::build does, and selector groups by its greedy rule
    // (selector_polynomials: extend a group while size + degree < max_quotient_degree_factor + 1)
    const bool recursion = spec.n_arith_ext + spec.n_mul_ext + spec.n_reducing + spec.n_reducing_ext + spec.n_random_access +
                           spec.n_exp + spec.n_coset + spec.n_mds > 0;
    int gate_index[NUM_KINDS];
    std::fill(gate_index, gate_index + NUM_KINDS, -1);
    if (!recursion) {
        c.gates = {{GATE_NOOP, 0}, {GATE_CONSTANT, 2}, {GATE_PUBLIC_INPUT, 0}, {GATE_BASE_SUM_2, 63}, {GATE_ARITHMETIC, 20}, {GATE_POSEIDON, 0}};
        const Kind kinds[] = {G_NOOP, G_CONST, G_PI, G_BASESUM, G_ARITH, G_POSEIDON};
        for (int i = 0; i < 6; ++i) gate_index[kinds[i]] = i;
    } else {
        // the set `verify_proof` instantiates under standard_recursion_config (135 wires, 80 routed, D = 2)
        Gate coset(GATE_COSET_INTERP, 4, 6);      // arity 16, with_max_degree(4, 8): 2 intermediates, degree 6
        {
            u64 gen = root_of_unity(4), xi = 1;
            std::vector<u64> xs(16);
            for (auto& x : xs) { x = xi; xi = fmul(xi, gen); }
            for (int i = 0; i < 16; ++i) {
                u64 d = 1;
                for (int j = 0; j < 16; ++j) if (j != i) d = fmul(d, fsub(xs[i], xs[j]));
                coset.weights.push_back(finv(d));
            }
        }
        struct KG { Kind k; Gate g; };
        const KG table[] = {
            {G_NOOP, {GATE_NOOP, 0}}, {G_CONST, {GATE_CONSTANT, 2}}, {G_MDS, {GATE_POSEIDON_MDS, 0}}, {G_PI, {GATE_PUBLIC_INPUT, 0}},
            {G_BASESUM, {GATE_BASE_SUM_2, 63}}, {G_REDUCING_EXT, {GATE_REDUCING_EXT, 32}}, {G_REDUCING, {GATE_REDUCING, 43}},
            {G_ARITH_EXT, {GATE_ARITHMETIC_EXT, 10}}, {G_ARITH, {GATE_ARITHMETIC, 20}}, {G_MUL_EXT, {GATE_MUL_EXT, 13}},
            {G_EXP, {GATE_EXPONENTIATION, 66}}, {G_RANDOM_ACCESS, {GATE_RANDOM_ACCESS, 4, 4, 2}}, {G_COSET, coset},
            {G_POSEIDON, {GATE_POSEIDON, 0}}};
        for (const KG& e : table) { gate_index[e.k] = (int)c.gates.size(); c.gates.push_back(e.g); }
    }
    c.selector_indices.assign(c.gates.size(), 0);
    for (size_t start = 0; start < c.gates.size();) {
        size_t size = 0;
        while (start + size < c.gates.size() && size + c.gates[start + size].degree() < c.max_quotient_degree_factor + 1) ++size;
        if (size == 0) throw std::runtime_error("gate degree too high for the quotient degree factor");
        for (size_t i = start; i < start + size; ++i) c.selector_indices[i] = c.groups.size();
        c.groups.push_back({start, start + size});
        start += size;
    }
    const size_t nsel = c.groups.size();
    c.quotient_degree_factor = 8; c.num_gate_constraints = 0; c.num_constants = nsel + 2;
    for (auto& g : c.gates) c.num_gate_constraints = std::max<u64>(c.num_gate_constraints, g.num_constraints());
    c.num_public_inputs = spec.num_public_inputs;
    c.k_is.resize(80);
    c.k_is[0] = 1;
    for (int j = 1; j < 80; ++j) c.k_is[j] = fmul(c.k_is[j - 1], GEN);
    c.num_partial_products = 9;

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
    auto take_ext = [&](u32 row, u32 col) {     // two routed wires = one F_{p^2} value (sequenced: the RNG stream is part of the spec)
        u64 a = take_input(row, col);
        u64 b = take_input(row, col + 1);
        return E2(a, b);
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

    const u64* rc = poseidon_round_constants();
    // Fills a Poseidon row whose input wires 0..11 and swap wire are already set; returns nothing.
    auto fill_poseidon = [&](u32 r) {
        std::vector<u64>& w = rows[r].w;
        u64 swap = w[PG_WIRE_SWAP];
        u64 st[12];
        for (int i = 0; i < 4; ++i) {
            u64 d = fmul(swap, fsub(w[i + 4], w[i]));
            w[PG_START_DELTA + i] = d;
            st[i] = fadd(w[i], d);
            st[i + 4] = fsub(w[i + 4], d);
        }
        for (int i = 8; i < 12; ++i) st[i] = w[i];
        int round = 0;
        for (int k = 0; k < HALF_N_FULL_ROUNDS; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = fadd(st[i], rc[12 * round + i]);
            if (k != 0) for (int i = 0; i < 12; ++i) w[PG_START_FULL_0 + 12 * (k - 1) + i] = st[i];
            for (int i = 0; i < 12; ++i) st[i] = sbox7<BaseOps>(st[i]);
            mds_layer<BaseOps>(st);
        }
        for (int k = 0; k < N_PARTIAL_ROUNDS; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = fadd(st[i], rc[12 * round + i]);
            w[PG_START_PARTIAL + k] = st[0];
            st[0] = sbox7<BaseOps>(st[0]);
            mds_layer<BaseOps>(st);
        }
        for (int k = 0; k < HALF_N_FULL_ROUNDS; ++k, ++round) {
            for (int i = 0; i < 12; ++i) st[i] = fadd(st[i], rc[12 * round + i]);
            for (int i = 0; i < 12; ++i) w[PG_START_FULL_1 + 12 * k + i] = st[i];
            for (int i = 0; i < 12; ++i) st[i] = sbox7<BaseOps>(st[i]);
            mds_layer<BaseOps>(st);
        }
        for (int i = 0; i < 12; ++i) w[12 + i] = st[i];
    };

    // public inputs hashed in-circuit by a sponge of Poseidon rows, digest wired to the PI gate
    sc.public_inputs.resize(spec.num_public_inputs);
    for (auto& p : sc.public_inputs) p = (rng.below(3) == 0) ? rng.felt() : rng.below(u64(1) << 32);
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
                if (i < m) rows[r].w[i] = sc.public_inputs[off + i];
                else if (k == 0) copy_from(r, (u32)i, ZERO);
                else copy_from(r, (u32)i, Pooled{state[i], {prev, (u32)(12 + i)}});
            }
            copy_from(r, PG_WIRE_SWAP, ZERO);
            fill_poseidon(r);
            for (int i = 0; i < 12; ++i) state[i] = rows[r].w[12 + i];
            prev = r;
        }
        Digest h = hash_no_pad(sc.public_inputs);
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
    for (size_t i = 0; i < spec.n_arith_ext; ++i) todo.push_back(G_ARITH_EXT);
    for (size_t i = 0; i < spec.n_mul_ext; ++i) todo.push_back(G_MUL_EXT);
    for (size_t i = 0; i < spec.n_reducing; ++i) todo.push_back(G_REDUCING);
    for (size_t i = 0; i < spec.n_reducing_ext; ++i) todo.push_back(G_REDUCING_EXT);
    for (size_t i = 0; i < spec.n_random_access; ++i) todo.push_back(G_RANDOM_ACCESS);
    for (size_t i = 0; i < spec.n_exp; ++i) todo.push_back(G_EXP);
    for (size_t i = 0; i < spec.n_coset; ++i) todo.push_back(G_COSET);
    for (size_t i = 0; i < spec.n_mds; ++i) todo.push_back(G_MDS);
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
                u64 out = fadd(fmul(rows[r].consts[0], fmul(m0, m1)), fmul(rows[r].consts[1], ad));
                rows[r].w[4 * i + 3] = out;
                pool.push_back({out, {r, 4 * i + 3}});
            }
        } else if (t == G_ARITH_EXT || t == G_MUL_EXT) {
            // F_{p^2} multiply-add rows of the recursive verifier (reduce_with_powers, opening combination)
            const bool addend = t == G_ARITH_EXT;
            const u32 per = addend ? 8 : 6, ops = addend ? 10 : 13;
            u32 r = new_row(t);
            rows[r].consts[0] = rng.below(2) ? 1 : rng.felt();
            if (addend) rows[r].consts[1] = rng.below(2) ? 1 : rng.felt();
            for (u32 i = 0; i < ops; ++i) {
                E2 a = take_ext(r, per * i);
                E2 b = take_ext(r, per * i + 2);
                E2 o = emul_base(a * b, rows[r].consts[0]);
                if (addend) o = o + emul_base(take_ext(r, per * i + 4), rows[r].consts[1]);
                rows[r].w[per * i + per - 2] = o.a;
                rows[r].w[per * i + per - 1] = o.b;
                pool.push_back({o.a, {r, per * i + per - 2}});
                pool.push_back({o.b, {r, per * i + per - 1}});
            }
        } else if (t == G_REDUCING || t == G_REDUCING_EXT) {
            // acc <- acc * alpha + coeff chains (FRI batch reduction inside the recursive verifier)
            const bool ext = t == G_REDUCING_EXT;
            const u32 nc = ext ? 32 : 43, start_accs = 6 + (ext ? 2 * nc : nc);
            u32 r = new_row(t);
            E2 alpha = take_ext(r, 2);
            E2 acc = take_ext(r, 4);
            for (u32 i = 0; i < nc; ++i) {
                E2 coeff = ext ? take_ext(r, 6 + 2 * i) : E2(take_input(r, 6 + i));
                acc = acc * alpha + coeff;
                u32 at = i + 1 == nc ? 0 : start_accs + 2 * i;
                rows[r].w[at] = acc.a;
                rows[r].w[at + 1] = acc.b;
            }
            pool.push_back({rows[r].w[0], {r, 0}});
            pool.push_back({rows[r].w[1], {r, 1}});
        } else if (t == G_RANDOM_ACCESS) {
            // 4 copies of a 16-way lookup (Merkle cap / coset selection) + 2 extra constants
            u32 r = new_row(t);
            for (u32 cpy = 0; cpy < 4; ++cpy) {
                const u32 base = 18 * cpy;
                u64 idx = rng.below(16);
                rows[r].w[base] = idx;
                for (u32 i = 0; i < 16; ++i) take_input(r, base + 2 + i);
                rows[r].w[base + 1] = rows[r].w[base + 2 + idx];
                pool.push_back({rows[r].w[base + 1], {r, base + 1}});
                for (u32 i = 0; i < 4; ++i) rows[r].w[74 + 4 * cpy + i] = (idx >> i) & 1;
            }
            for (u32 k = 0; k < 2; ++k) {
                u64 v = rng.below(2) ? rng.felt() : rng.below(256);
                rows[r].consts[k] = v;
                rows[r].w[72 + k] = v;
                pool.push_back({v, {r, 72 + k}});
            }
        } else if (t == G_EXP) {
            // base^(66 power bits), square-and-multiply with intermediate wires
            u32 r = new_row(t);
            u64 base = take_input(r, 0), cur = 1;
            for (u32 i = 0; i < 66; ++i) rows[r].w[1 + i] = rng.below(2);
            for (u32 i = 0; i < 66; ++i) {
                u64 prev = i == 0 ? 1 : fsqr(cur);
                cur = rows[r].w[1 + (65 - i)] ? fmul(prev, base) : prev;
                rows[r].w[68 + i] = cur;
            }
            rows[r].w[67] = cur;
            pool.push_back({cur, {r, 67}});
            for (u32 k = 0; k < 2; ++k) { u32 b = (u32)rng.below(66); bools.push_back({rows[r].w[1 + b], {r, 1 + b}}); }
        } else if (t == G_COSET) {
            // barycentric interpolation of 16 F_{p^2} values over a coset shift * <w_16> at an F_{p^2} point (FRI fold check)
            u32 r = new_row(t);
            const Gate& g = c.gates[gate_index[G_COSET]];
            u64 shift;
            do shift = rng.felt(); while (shift == 0);
            rows[r].w[0] = shift;
            for (u32 i = 0; i < 32; ++i) take_input(r, 1 + i);
            E2 point = take_ext(r, 33);
            E2 sh = emul_base(point, finv(shift));
            rows[r].w[45] = sh.a;
            rows[r].w[46] = sh.b;
            using A = Alg<BaseOps>;
            A shifted{sh.a, sh.b}, eval = A::zero(), prod = A::one();
            const u64* w = rows[r].w.data();
            partial_interpolate<BaseOps>(g, w, 0, 6, shifted, eval, prod);
            for (u32 i = 0; i < 2; ++i) {
                rows[r].w[37 + 2 * i] = eval.a; rows[r].w[38 + 2 * i] = eval.b;
                rows[r].w[41 + 2 * i] = prod.a; rows[r].w[42 + 2 * i] = prod.b;
                partial_interpolate<BaseOps>(g, rows[r].w.data(), 6 + 5 * i, 11 + 5 * i, shifted, eval, prod);
            }
            rows[r].w[35] = eval.a;
            rows[r].w[36] = eval.b;
            pool.push_back({eval.a, {r, 35}});
            pool.push_back({eval.b, {r, 36}});
        } else if (t == G_MDS) {
            // the Poseidon MDS layer applied to 12 F_{p^2} values (recursive Poseidon gate evaluation)
            u32 r = new_row(t);
            E2 in[12];
            for (u32 i = 0; i < 12; ++i) in[i] = take_ext(r, 2 * i);
            for (u32 k = 0; k < 12; ++k) {
                E2 acc = emul_base(in[k], MDS_DIAG[k]);
                for (u32 i = 0; i < 12; ++i) acc = acc + emul_base(in[(i + k) % 12], MDS_CIRC[i]);
                rows[r].w[24 + 2 * k] = acc.a;
                rows[r].w[25 + 2 * k] = acc.b;
                pool.push_back({acc.a, {r, 24 + 2 * k}});
                pool.push_back({acc.b, {r, 25 + 2 * k}});
            }
        } else if (t == G_POSEIDON) {
            u32 r = new_row(G_POSEIDON);
            for (u32 i = 0; i < 8; ++i) take_input(r, i);
            for (u32 i = 8; i < 12; ++i) copy_from(r, i, ZERO);
            copy_from(r, PG_WIRE_SWAP, bools[rng.below(bools.size())]);
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
                copy_from(r, PG_WIRE_SWAP, ZERO);
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
            auto arities = fri_reduction_arity_bits(c.fri_config, est_bits);
            u64 fold = 0, asum = 0;
            for (u64 a : arities) { fold += (u64(1) << a) - 1; asum += a; }
            u64 final_coeffs = (u64(1) << est_bits) >> asum;
            u64 fri_open = c.fri_config.num_query_rounds * (1 + 2 * fold + 2 * final_coeffs);
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
    c.degree_bits = db;
    c.reduction_arity_bits = fri_reduction_arity_bits(c.fri_config, db);

    // wires, constants
    sc.wires.assign(135, std::vector<u64>(n));
    sc.const_sigma_values.assign(nsel + 2 + 80, std::vector<u64>(n));
    for (size_t r = 0; r < n; ++r) {
        for (int j = 0; j < 135; ++j) sc.wires[j][r] = rows[r].w[j];
        const int g = gate_index[rows[r].gate];
        if (g < 0) throw std::runtime_error("row kind outside the circuit's gate set");
        for (size_t k = 0; k < nsel; ++k) sc.const_sigma_values[k][r] = c.selector_indices[g] == k ? (u64)g : UNUSED_SELECTOR;
        sc.const_sigma_values[nsel][r] = rows[r].consts[0];
        sc.const_sigma_values[nsel + 1][r] = rows[r].consts[1];
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
    u64 w = root_of_unity(db);
    std::vector<u64> subgroup(n);
    subgroup[0] = 1;
    for (size_t i = 1; i < n; ++i) subgroup[i] = fmul(subgroup[i - 1], w);
    for (u32 col = 0; col < 80; ++col)
        for (size_t r = 0; r < n; ++r) {
            u32 t = sigma[col * n + r];
            sc.const_sigma_values[nsel + 2 + col][r] = fmul(c.k_is[t / n], subgroup[t % n]);
        }
    return sc;
}

std::string check_witness(const SynthCircuit& sc) {
    const CommonData& c = sc.common;
    size_t n = c.degree();
    Digest pi_hash = hash_no_pad(sc.public_inputs);
    std::vector<u64> out, w(135);
    for (size_t r = 0; r < n; ++r) {
        const size_t nsel = c.num_selectors();
        size_t g = c.gates.size();
        for (size_t k = 0; k < nsel; ++k)
            if (sc.const_sigma_values[k][r] != UNUSED_SELECTOR) g = sc.const_sigma_values[k][r];
        if (g >= c.gates.size()) return "bad selector at row " + std::to_string(r);
        for (int j = 0; j < 135; ++j) w[j] = sc.wires[j][r];
        u64 consts[2] = {sc.const_sigma_values[nsel][r], sc.const_sigma_values[nsel + 1][r]};
        eval_gate_unfiltered<BaseOps>(c.gates[g], consts, w.data(), pi_hash, out);
        for (size_t k = 0; k < out.size(); ++k)
            if (out[k] != 0) return "gate constraint " + std::to_string(k) + " fails at row " + std::to_string(r);
    }
    // copy constraints: wire value at a cell equals the value at sigma(cell); recover sigma from its value
    u64 w_n = root_of_unity((unsigned)c.degree_bits);
    std::vector<u64> subgroup(n);
    subgroup[0] = 1;
    for (size_t i = 1; i < n; ++i) subgroup[i] = fmul(subgroup[i - 1], w_n);
    // product check (the grand product must telescope to 1) with fixed pseudo-random beta/gamma
    u64 beta = 0x123456789abcdefULL, gamma = 0xfedcba987654321ULL, num = 1, den = 1;
    for (size_t r = 0; r < n; ++r)
        for (size_t j = 0; j < 80; ++j) {
            num = fmul(num, fadd(fadd(sc.wires[j][r], fmul(beta, fmul(c.k_is[j], subgroup[r]))), gamma));
            den = fmul(den, fadd(fadd(sc.wires[j][r], fmul(beta, sc.const_sigma_values[c.num_constants + j][r])), gamma));
        }
    if (num != den) return "permutation grand product does not telescope";
    return "";
}

}  // namespace orc
