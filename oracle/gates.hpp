// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// Gate constraint polynomials and the Plonk vanishing polynomial, written once over a generic
// field (base field for the prover's quotient evaluation, F_{p^2} for the verifier's check at zeta).
// Restates qp-plonky2 1.1.1 `plonk::vanishing_poly::{eval_vanishing_poly, eval_vanishing_poly_base_batch}`,
// `gates::{noop,constant,public_input,base_sum,arithmetic_base,poseidon}` and `gates::selectors`
// (reached from /root/reference/wormhole/prover/src/lib.rs:234-236 and
// /root/reference/wormhole/verifier/src/lib.rs:156-157). Formulas: SURVEY.md A.5 step 2, App. C.1.
// Pinned by: the vanishing identity of wormhole/bench-data/proof.bin (all six gates of the wormhole / voting set).
//
// The RECURSION gate set (ArithmeticExtension, MulExtension, Reducing, ReducingExtension, RandomAccess, Exponentiation,
// CosetInterpolation, PoseidonMds — instantiated by `verify_proof` at /root/reference/wormhole/aggregator/src/circuits/tree.rs:119)
// restates upstream Plonky2's `gates::{arithmetic_extension, multiplication_extension, reducing, reducing_extension,
// random_access, exponentiation, coset_interpolation, poseidon_mds}` from their published definitions.
// PARITY UNPINNED for these eight: the reference tree holds no recursion-circuit common data / verifier key, so their
// wire layouts and constraint order cannot be checked against a reference fixture here; prover and verifier are
// self-consistent (proofs verify), nothing more is claimed.
#pragma once
#include "circuit.hpp"

namespace orc {

constexpr u64 UNUSED_SELECTOR = 0xFFFFFFFFULL;

// Poseidon gate wire layout (C.1)
constexpr int PG_WIRE_SWAP = 24;
constexpr int PG_START_DELTA = 25;
constexpr int PG_START_FULL_0 = 29;
constexpr int PG_START_PARTIAL = 65;
constexpr int PG_START_FULL_1 = 87;

// Extension algebra over T: pairs (a, b) = a + b X with X^2 = 7 and coefficients in T. For T = F_p this is F_{p^2}
// itself (the prover's base-field evaluation); for T = F_{p^2} it is upstream's ExtensionAlgebra (the verifier's evaluation).
template <class Ops>
struct Alg {
    using T = typename Ops::T;
    T a, b;
    static Alg zero() { return {Ops::zero(), Ops::zero()}; }
    static Alg one() { return {Ops::one(), Ops::zero()}; }
    static Alg from_base(T x) { return {x, Ops::zero()}; }
    static Alg read(const T* w, size_t at) { return {w[at], w[at + 1]}; }
    Alg operator+(const Alg& o) const { return {Ops::add(a, o.a), Ops::add(b, o.b)}; }
    Alg operator-(const Alg& o) const { return {Ops::sub(a, o.a), Ops::sub(b, o.b)}; }
    Alg operator*(const Alg& o) const {
        return {Ops::add(Ops::mul(a, o.a), Ops::mulc(Ops::mul(b, o.b), EXT_W)), Ops::add(Ops::mul(a, o.b), Ops::mul(b, o.a))};
    }
    Alg scale(T s) const { return {Ops::mul(a, s), Ops::mul(b, s)}; }
    Alg scalec(u64 c) const { return {Ops::mulc(a, c), Ops::mulc(b, c)}; }
    void push(std::vector<T>& out) const { out.push_back(a); out.push_back(b); }
};

// CosetInterpolation: one chunk of the barycentric recurrence over points x_k = w^k of the size-2^bits subgroup
template <class Ops>
void partial_interpolate(const Gate& g, const typename Ops::T* w, size_t lo, size_t hi, const Alg<Ops>& point,
                         Alg<Ops>& eval, Alg<Ops>& prod) {
    u64 gen = root_of_unity((unsigned)g.param), x = fpow(gen, lo);
    for (size_t k = lo; k < hi; ++k, x = fmul(x, gen)) {
        Alg<Ops> val = Alg<Ops>::read(w, 1 + 2 * k);
        Alg<Ops> term = point - Alg<Ops>::from_base(Ops::from(x));
        eval = eval * term + val.scalec(g.weights[k]) * prod;
        prod = prod * term;
    }
}

template <class Ops>
void eval_gate_unfiltered(const Gate& g, const typename Ops::T* consts, const typename Ops::T* w,
                          const Digest& pi_hash, std::vector<typename Ops::T>& out) {
    using T = typename Ops::T;
    out.clear();
    switch (g.tag) {
        case GATE_NOOP: break;
        case GATE_CONSTANT:
            for (u64 i = 0; i < g.param; ++i) out.push_back(Ops::sub(consts[i], w[i]));
            break;
        case GATE_PUBLIC_INPUT:
            for (int i = 0; i < 4; ++i) out.push_back(Ops::sub(w[i], Ops::from(pi_hash[i])));
            break;
        case GATE_BASE_SUM_2: {
            T sum = Ops::zero();
            for (u64 k = g.param; k-- > 0;) sum = Ops::add(Ops::mulc(sum, 2), w[1 + k]);
            out.push_back(Ops::sub(sum, w[0]));
            for (u64 k = 0; k < g.param; ++k) out.push_back(Ops::mul(w[1 + k], Ops::sub(w[1 + k], Ops::one())));
            break;
        }
        case GATE_ARITHMETIC:
            for (u64 i = 0; i < g.param; ++i) {
                T prod = Ops::mul(Ops::mul(w[4 * i], w[4 * i + 1]), consts[0]);
                T add = Ops::mul(w[4 * i + 2], consts[1]);
                out.push_back(Ops::sub(w[4 * i + 3], Ops::add(prod, add)));
            }
            break;
        case GATE_POSEIDON: {
            const u64* rc = poseidon_round_constants();
            T swap = w[PG_WIRE_SWAP];
            out.push_back(Ops::mul(swap, Ops::sub(swap, Ops::one())));
            for (int i = 0; i < 4; ++i)
                out.push_back(Ops::sub(Ops::mul(swap, Ops::sub(w[i + 4], w[i])), w[PG_START_DELTA + i]));
            T st[12];
            for (int i = 0; i < 4; ++i) {
                st[i] = Ops::add(w[i], w[PG_START_DELTA + i]);
                st[i + 4] = Ops::sub(w[i + 4], w[PG_START_DELTA + i]);
            }
            for (int i = 8; i < 12; ++i) st[i] = w[i];
            int round = 0;
            for (int r = 0; r < HALF_N_FULL_ROUNDS; ++r, ++round) {
                for (int i = 0; i < 12; ++i) st[i] = Ops::add(st[i], Ops::from(rc[12 * round + i]));
                if (r != 0) {
                    for (int i = 0; i < 12; ++i) {
                        T sin = w[PG_START_FULL_0 + 12 * (r - 1) + i];
                        out.push_back(Ops::sub(st[i], sin));
                        st[i] = sin;
                    }
                }
                for (int i = 0; i < 12; ++i) st[i] = sbox7<Ops>(st[i]);
                mds_layer<Ops>(st);
            }
            for (int r = 0; r < N_PARTIAL_ROUNDS; ++r, ++round) {
                for (int i = 0; i < 12; ++i) st[i] = Ops::add(st[i], Ops::from(rc[12 * round + i]));
                T sin = w[PG_START_PARTIAL + r];
                out.push_back(Ops::sub(st[0], sin));
                st[0] = sbox7<Ops>(sin);
                mds_layer<Ops>(st);
            }
            for (int r = 0; r < HALF_N_FULL_ROUNDS; ++r, ++round) {
                for (int i = 0; i < 12; ++i) st[i] = Ops::add(st[i], Ops::from(rc[12 * round + i]));
                for (int i = 0; i < 12; ++i) {
                    T sin = w[PG_START_FULL_1 + 12 * r + i];
                    out.push_back(Ops::sub(st[i], sin));
                    st[i] = sin;
                }
                for (int i = 0; i < 12; ++i) st[i] = sbox7<Ops>(st[i]);
                mds_layer<Ops>(st);
            }
            for (int i = 0; i < 12; ++i) out.push_back(Ops::sub(st[i], w[12 + i]));
            break;
        }
        case GATE_ARITHMETIC_EXT:      // wires per op: a[2] b[2] addend[2] out[2]; out - (c0 a b + c1 addend)
            for (u64 i = 0; i < g.param; ++i) {
                using A = Alg<Ops>;
                A m0 = A::read(w, 8 * i), m1 = A::read(w, 8 * i + 2), ad = A::read(w, 8 * i + 4), o = A::read(w, 8 * i + 6);
                (o - ((m0 * m1).scale(consts[0]) + ad.scale(consts[1]))).push(out);
            }
            break;
        case GATE_MUL_EXT:             // wires per op: a[2] b[2] out[2]; out - c0 a b
            for (u64 i = 0; i < g.param; ++i) {
                using A = Alg<Ops>;
                A m0 = A::read(w, 6 * i), m1 = A::read(w, 6 * i + 2), o = A::read(w, 6 * i + 4);
                (o - (m0 * m1).scale(consts[0])).push(out);
            }
            break;
        case GATE_REDUCING:            // out[0..2] alpha[2..4] old_acc[4..6] coeffs[6 .. 6+n) accs (n-1 pairs); last acc = out
        case GATE_REDUCING_EXT: {      // same with extension coefficients (pairs) at 6 + 2i
            using A = Alg<Ops>;
            const bool ext = g.tag == GATE_REDUCING_EXT;
            const size_t nc = g.param, start_accs = 6 + (ext ? 2 * nc : nc);
            A alpha = A::read(w, 2), acc = A::read(w, 4);
            for (size_t i = 0; i < nc; ++i) {
                A coeff = ext ? A::read(w, 6 + 2 * i) : A::from_base(w[6 + i]);
                A next = i + 1 == nc ? A::read(w, 0) : A::read(w, start_accs + 2 * i);
                ((acc * alpha + coeff) - next).push(out);      // upstream: acc * alpha + coeff - accs[i]
                acc = next;
            }
            break;
        }
        case GATE_RANDOM_ACCESS: {     // per copy: index, claimed, 2^bits items (routed); extra constants; then the bits
            const size_t bits = g.param, vec = size_t(1) << bits, copies = g.p2, extra = g.p3;
            const size_t routed = (2 + vec) * copies + extra;
            for (size_t cpy = 0; cpy < copies; ++cpy) {
                const size_t base = (2 + vec) * cpy;
                std::vector<T> items(w + base + 2, w + base + 2 + vec), b(bits);
                for (size_t i = 0; i < bits; ++i) b[i] = w[routed + cpy * bits + i];
                for (size_t i = 0; i < bits; ++i) out.push_back(Ops::mul(b[i], Ops::sub(b[i], Ops::one())));
                T idx = Ops::zero();
                for (size_t i = bits; i-- > 0;) idx = Ops::add(Ops::add(idx, idx), b[i]);
                out.push_back(Ops::sub(idx, w[base]));
                for (size_t i = 0; i < bits; ++i) {
                    std::vector<T> nx;
                    for (size_t k = 0; k + 1 < items.size(); k += 2)
                        nx.push_back(Ops::add(items[k], Ops::mul(b[i], Ops::sub(items[k + 1], items[k]))));
                    items.swap(nx);
                }
                out.push_back(Ops::sub(items[0], w[base + 1]));
            }
            for (size_t i = 0; i < extra; ++i) out.push_back(Ops::sub(consts[i], w[(2 + vec) * copies + i]));
            break;
        }
        case GATE_EXPONENTIATION: {    // base 0, power bits 1..n (LE), output n+1, intermediates n+2..
            const size_t nb = g.param;
            T base = w[0];
            for (size_t i = 0; i < nb; ++i) {
                T prev = i == 0 ? Ops::one() : Ops::mul(w[nb + 2 + i - 1], w[nb + 2 + i - 1]);
                T bit = w[1 + (nb - 1 - i)];
                T factor = Ops::add(Ops::mul(bit, base), Ops::sub(Ops::one(), bit));
                out.push_back(Ops::sub(Ops::mul(prev, factor), w[nb + 2 + i]));
            }
            out.push_back(Ops::sub(w[nb + 1], w[nb + 2 + nb - 1]));
            break;
        }
        case GATE_COSET_INTERP: {      // shift 0; values 1..; point; value; intermediates (eval_i, then prod_i); shifted point
            using A = Alg<Ops>;
            const size_t np = size_t(1) << g.param, deg = g.p2, ni = coset_interp_num_intermediates(g);
            const size_t at_point = 1 + 2 * np, at_value = at_point + 2, at_inter = at_value + 2, at_shifted = at_inter + 4 * ni;
            A point = A::read(w, at_point), shifted = A::read(w, at_shifted);
            (point - shifted.scale(w[0])).push(out);
            A eval = A::zero(), prod = A::one();
            partial_interpolate<Ops>(g, w, 0, deg, shifted, eval, prod);
            for (size_t i = 0; i < ni; ++i) {
                A ie = A::read(w, at_inter + 2 * i), ip = A::read(w, at_inter + 2 * (ni + i));
                (ie - eval).push(out);
                (ip - prod).push(out);
                size_t lo = 1 + (deg - 1) * (i + 1), hi = std::min(lo + deg - 1, np);
                eval = ie; prod = ip;
                partial_interpolate<Ops>(g, w, lo, hi, shifted, eval, prod);
            }
            (A::read(w, at_value) - eval).push(out);
            break;
        }
        case GATE_POSEIDON_MDS: {      // inputs 12 pairs at 2i, outputs 12 pairs at 24 + 2i
            using A = Alg<Ops>;
            for (int r = 0; r < 12; ++r) {
                A acc = A::zero();
                for (int i = 0; i < 12; ++i) acc = acc + A::read(w, 2 * ((i + r) % 12)).scalec(MDS_CIRC[i]);
                acc = acc + A::read(w, 2 * r).scalec(MDS_DIAG[r]);
                (A::read(w, 24 + 2 * r) - acc).push(out);
            }
            break;
        }
        default: throw std::runtime_error("unsupported gate");
    }
}

template <class Ops>
typename Ops::T compute_filter(size_t row, std::pair<u64, u64> group, typename Ops::T s, bool many_selectors) {
    using T = typename Ops::T;
    T f = Ops::one();
    for (u64 i = group.first; i < group.second; ++i)
        if (i != row) f = Ops::mul(f, Ops::sub(Ops::from(i), s));
    if (many_selectors) f = Ops::mul(f, Ops::sub(Ops::from(UNUSED_SELECTOR), s));
    return f;
}

// Returns, per challenge c, sum_k terms_k * alpha_c^k with terms = [L0(x)(Z_c-1)]_c ‖ [pp checks]_c ‖ gate constraints.
// `x` is the evaluation point; constants has c.num_constants entries (selectors first).
template <class Ops>
std::vector<typename Ops::T> eval_vanishing(const CommonData& c, typename Ops::T x, typename Ops::T l0_x,
                                            const typename Ops::T* constants, const typename Ops::T* sigmas,
                                            const typename Ops::T* wires, const typename Ops::T* zs,
                                            const typename Ops::T* zs_next, const typename Ops::T* pps,
                                            const Digest& pi_hash, const u64* betas, const u64* gammas,
                                            const u64* alphas) {
    using T = typename Ops::T;
    size_t nch = c.num_challenges, npp = c.num_partial_products, chunk = c.quotient_degree_factor;
    // scratch reused across calls (the prover calls this once per LDE point from every OpenMP thread)
    static thread_local std::vector<T> z1_terms, pp_terms, accs, constraints, gc, terms;
    z1_terms.clear(); pp_terms.clear();
    for (size_t ch = 0; ch < nch; ++ch) {
        z1_terms.push_back(Ops::mul(l0_x, Ops::sub(zs[ch], Ops::one())));
        accs.clear();
        accs.push_back(zs[ch]);
        for (size_t k = 0; k < npp; ++k) accs.push_back(pps[ch * npp + k]);
        accs.push_back(zs_next[ch]);
        size_t nchunks = (c.num_routed_wires + chunk - 1) / chunk;
        for (size_t k = 0; k < nchunks; ++k) {
            T num = Ops::one(), den = Ops::one();
            for (size_t j = k * chunk; j < (k + 1) * chunk && j < c.num_routed_wires; ++j) {
                T s_id = Ops::mulc(x, c.k_is[j]);
                T nj = Ops::add(Ops::add(wires[j], Ops::mulc(s_id, betas[ch])), Ops::from(gammas[ch]));
                T dj = Ops::add(Ops::add(wires[j], Ops::mulc(sigmas[j], betas[ch])), Ops::from(gammas[ch]));
                num = Ops::mul(num, nj);
                den = Ops::mul(den, dj);
            }
            pp_terms.push_back(Ops::sub(Ops::mul(accs[k], num), Ops::mul(accs[k + 1], den)));
        }
    }
    constraints.assign(c.num_gate_constraints, Ops::zero());
    size_t nsel = c.num_selectors();
    for (size_t gi = 0; gi < c.gates.size(); ++gi) {
        u64 sel = c.selector_indices[gi];
        T filter = compute_filter<Ops>(gi, c.groups[sel], constants[sel], nsel > 1);
        eval_gate_unfiltered<Ops>(c.gates[gi], constants + nsel, wires, pi_hash, gc);
        for (size_t k = 0; k < gc.size(); ++k) constraints[k] = Ops::add(constraints[k], Ops::mul(filter, gc[k]));
    }
    terms.clear();
    terms.insert(terms.end(), z1_terms.begin(), z1_terms.end());
    terms.insert(terms.end(), pp_terms.begin(), pp_terms.end());
    terms.insert(terms.end(), constraints.begin(), constraints.end());
    std::vector<T> res;
    for (size_t ch = 0; ch < nch; ++ch) {
        T acc = Ops::zero();
        for (size_t k = terms.size(); k-- > 0;) acc = Ops::add(Ops::mulc(acc, alphas[ch]), terms[k]);
        res.push_back(acc);
    }
    return res;
}

}  // namespace orc
