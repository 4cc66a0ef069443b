// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// Gate constraint polynomials and the Plonk vanishing polynomial, written once over a generic
// field (base field for the prover's quotient evaluation, F_{p^2} for the verifier's check at zeta).
// Restates qp-plonky2 1.1.1 `plonk::vanishing_poly::{eval_vanishing_poly, eval_vanishing_poly_base_batch}`,
// `gates::{noop,constant,public_input,base_sum,arithmetic_base,poseidon}` and `gates::selectors`
// (reached from /root/reference/wormhole/prover/src/lib.rs:234-236 and
// /root/reference/wormhole/verifier/src/lib.rs:156-157). Formulas: SURVEY.md A.5 step 2, App. C.1.
// Pinned by: the vanishing identity of wormhole/bench-data/proof.bin (all six gates).
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
    std::vector<T> z1_terms, pp_terms;
    for (size_t ch = 0; ch < nch; ++ch) {
        z1_terms.push_back(Ops::mul(l0_x, Ops::sub(zs[ch], Ops::one())));
        std::vector<T> accs;
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
    std::vector<T> constraints(c.num_gate_constraints, Ops::zero());
    std::vector<T> gc;
    size_t nsel = c.num_selectors();
    for (size_t gi = 0; gi < c.gates.size(); ++gi) {
        u64 sel = c.selector_indices[gi];
        T filter = compute_filter<Ops>(gi, c.groups[sel], constants[sel], nsel > 1);
        eval_gate_unfiltered<Ops>(c.gates[gi], constants + nsel, wires, pi_hash, gc);
        for (size_t k = 0; k < gc.size(); ++k) constraints[k] = Ops::add(constraints[k], Ops::mul(filter, gc[k]));
    }
    std::vector<T> terms;
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
