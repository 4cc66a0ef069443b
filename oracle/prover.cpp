// ORACLE — TEST INFRASTRUCTURE ONLY (see prover.hpp header for scope, citations and parity status).
#include "prover.hpp"
#include "gates.hpp"
#include "vec_ops.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace orc {

PolyBatch batch_from_coeffs(std::vector<std::vector<u64>> coeffs, unsigned rate_bits, bool blinding,
                            unsigned cap_height, const u64* salts, u64 seed, unsigned batch_id) {
    PolyBatch b;
    b.ncols = coeffs.size();
    b.n = coeffs.at(0).size();
    b.rate_bits = rate_bits;
    b.blinding = blinding;
    size_t N = b.n << rate_bits;
    unsigned lg = log2_strict(N);
    size_t width = b.ncols + (blinding ? 4 : 0);
    std::vector<u64> leaves(N * width);
#pragma omp parallel for schedule(dynamic)
    for (long c = 0; c < (long)b.ncols; ++c) {
        if (coeffs[c].size() != b.n) continue;
        if (g_fast_poseidon) {       // CPU-baseline arm: 2^rate_bits coset transforms of size n, already in leaf order
            std::vector<u64> v(N);
            lde_coset_leaf_order_fast(coeffs[c], rate_bits, GEN, v.data());
            for (size_t l = 0; l < N; ++l) leaves[l * width + c] = v[l];
            continue;
        }
        std::vector<u64> v = lde_coset<u64>(coeffs[c], rate_bits);
        for (size_t i = 0; i < N; ++i) leaves[reverse_bits(i, lg) * width + c] = v[i];
    }
    for (auto& c : coeffs)
        if (c.size() != b.n) throw std::runtime_error("ragged polynomial batch");
    if (blinding) {
        for (unsigned s = 0; s < 4; ++s)
            for (size_t l = 0; l < N; ++l)
                leaves[l * width + b.ncols + s] = salts ? salts[(size_t)s * N + l] : salt_value(seed, batch_id, s, l);
    }
    b.coeffs = std::move(coeffs);
    b.tree = merkle_build(std::move(leaves), N, width, cap_height);
    return b;
}

PolyBatch batch_from_values(std::vector<std::vector<u64>> values, unsigned rate_bits, bool blinding,
                            unsigned cap_height, const u64* salts, u64 seed, unsigned batch_id) {
#pragma omp parallel for schedule(dynamic)
    for (long c = 0; c < (long)values.size(); ++c) ifft(values[c]);
    return batch_from_coeffs(std::move(values), rate_bits, blinding, cap_height, salts, seed, batch_id);
}

CircuitData circuit_from_values(const CommonData& c, std::vector<std::vector<u64>> csv) {
    CircuitData cd;
    cd.common = c;
    if (csv.size() != c.num_constants + c.num_routed_wires) throw std::runtime_error("bad constants/sigmas column count");
    cd.sigma_values.assign(csv.begin() + c.num_constants, csv.end());
    cd.constants_sigmas = batch_from_values(std::move(csv), (unsigned)c.fri_config.rate_bits, false,
                                            (unsigned)c.fri_config.cap_height, nullptr, 0, 0);
    cd.vo.constants_sigmas_cap = cd.constants_sigmas.tree.cap;
    cd.vo.circuit_digest = compute_circuit_digest(cd.vo.constants_sigmas_cap, c.degree_bits);
    return cd;
}

std::vector<std::vector<u64>> partial_products_and_zs(const CircuitData& cd, const std::vector<std::vector<u64>>& wires,
                                                      const std::vector<u64>& betas, const std::vector<u64>& gammas) {
    const CommonData& c = cd.common;
    size_t n = c.degree(), nr = c.num_routed_wires, chunk = c.quotient_degree_factor, npp = c.num_partial_products;
    size_t nchunks = (nr + chunk - 1) / chunk;
    if (nchunks != npp + 1) throw std::runtime_error("partial product count mismatch");
    size_t nch = c.num_challenges;
    std::vector<std::vector<u64>> out(nch * (1 + npp), std::vector<u64>(n));
    u64 w = root_of_unity((unsigned)c.degree_bits);
    std::vector<u64> subgroup(n);
    subgroup[0] = 1;
    for (size_t i = 1; i < n; ++i) subgroup[i] = fmul(subgroup[i - 1], w);
    for (size_t ch = 0; ch < nch; ++ch) {
        u64 beta = betas[ch], gamma = gammas[ch];
        std::vector<u64> chunk_prod(n * nchunks);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)n; ++i) {
            std::vector<u64> num(nr), den(nr), pref(nr);
            for (size_t j = 0; j < nr; ++j) {
                u64 wv = wires[j][i];
                num[j] = fadd(fadd(wv, fmul(beta, fmul(c.k_is[j], subgroup[i]))), gamma);
                den[j] = fadd(fadd(wv, fmul(beta, cd.sigma_values[j][i])), gamma);
            }
            // batch inversion of den
            u64 acc = 1;
            for (size_t j = 0; j < nr; ++j) { pref[j] = acc; acc = fmul(acc, den[j]); }
            u64 inv = finv(acc);
            for (size_t j = nr; j-- > 0;) { u64 di = fmul(inv, pref[j]); inv = fmul(inv, den[j]); den[j] = di; }
            for (size_t k = 0; k < nchunks; ++k) {
                u64 p = 1;
                for (size_t j = k * chunk; j < std::min(nr, (k + 1) * chunk); ++j) p = fmul(p, fmul(num[j], den[j]));
                chunk_prod[i * nchunks + k] = p;
            }
        }
        u64 z = 1;
        for (size_t i = 0; i < n; ++i) {
            out[ch][i] = z;
            u64 acc = z;
            for (size_t k = 0; k < nchunks; ++k) {
                acc = fmul(acc, chunk_prod[i * nchunks + k]);
                if (k < npp) out[nch + ch * npp + k][i] = acc;
            }
            z = acc;
        }
    }
    return out;
}

std::vector<std::vector<u64>> compute_quotient_chunks(const CircuitData& cd, const PolyBatch& wires_b, const PolyBatch& zs_b,
                                                      const Digest& pi_hash, const std::vector<u64>& betas,
                                                      const std::vector<u64>& gammas, const std::vector<u64>& alphas) {
    const CommonData& c = cd.common;
    size_t n = c.degree(), nch = c.num_challenges, qdf = c.quotient_degree_factor;
    unsigned rate_bits = (unsigned)c.fri_config.rate_bits;
    if ((size_t(1) << rate_bits) != qdf) throw std::runtime_error("oracle supports quotient_degree_factor == 2^rate_bits only");
    size_t N = n << rate_bits;
    unsigned lgN = log2_strict(N);
    u64 wN = root_of_unity(lgN);
    std::vector<u64> xs(N);
    xs[0] = GEN;
    for (size_t i = 1; i < N; ++i) xs[i] = fmul(xs[i - 1], wN);
    // Z_H on the coset: g^n * w_rate^(i mod rate) - 1
    u64 g_n = fpow(GEN, n);
    u64 w_rate = root_of_unity(rate_bits);
    std::vector<u64> zh(qdf), zh_inv(qdf);
    for (size_t j = 0; j < qdf; ++j) { zh[j] = fsub(fmul(g_n, fpow(w_rate, j)), 1); zh_inv[j] = finv(zh[j]); }
    u64 n_f = from_u64(n);
    size_t next_step = qdf;
    std::vector<std::vector<u64>> q(nch, std::vector<u64>(N));
    size_t ncs = c.num_constants;
#if defined(__AVX2__)
    if (g_fast_poseidon && N % 4 == 0) {
        // CPU-baseline arm: the same generic constraint code (gates.hpp) instantiated over four lanes (vec_ops.hpp), four LDE
        // points per call; rows are gathered into lane vectors first
        const size_t nzs = nch * (1 + c.num_partial_products);
#pragma omp parallel
        {
            std::vector<V4> vcs(ncs + c.num_routed_wires), vw(c.num_wires), vz(nzs), vzn(nch);
#pragma omp for schedule(static)
            for (long i4 = 0; i4 < (long)(N / 4); ++i4) {
                const size_t i = (size_t)i4 * 4;
                const u64 *cs[4], *wr[4], *zr[4], *zn[4];
                u64 l0[4];
                for (int k = 0; k < 4; ++k) {
                    cs[k] = cd.constants_sigmas.lde_row(i + k);
                    wr[k] = wires_b.lde_row(i + k);
                    zr[k] = zs_b.lde_row(i + k);
                    zn[k] = zs_b.lde_row((i + k + next_step) % N);
                    l0[k] = fmul(zh[(i + k) % qdf], finv(fmul(n_f, fsub(xs[i + k], 1))));
                }
                for (size_t j = 0; j < vcs.size(); ++j) vcs[j] = VecOps::lanes(cs[0][j], cs[1][j], cs[2][j], cs[3][j]);
                for (size_t j = 0; j < vw.size(); ++j) vw[j] = VecOps::lanes(wr[0][j], wr[1][j], wr[2][j], wr[3][j]);
                for (size_t j = 0; j < vz.size(); ++j) vz[j] = VecOps::lanes(zr[0][j], zr[1][j], zr[2][j], zr[3][j]);
                for (size_t j = 0; j < nch; ++j) vzn[j] = VecOps::lanes(zn[0][j], zn[1][j], zn[2][j], zn[3][j]);
                auto van = eval_vanishing<VecOps>(c, VecOps::lanes(xs[i], xs[i + 1], xs[i + 2], xs[i + 3]),
                                                  VecOps::lanes(l0[0], l0[1], l0[2], l0[3]), vcs.data(), vcs.data() + ncs, vw.data(),
                                                  vz.data(), vzn.data(), vz.data() + nch, pi_hash, betas.data(), gammas.data(),
                                                  alphas.data());
                for (size_t ch = 0; ch < nch; ++ch) {
                    u64 out[4];
                    VecOps::store(van[ch], out);
                    for (int k = 0; k < 4; ++k) q[ch][i + k] = fmul(out[k], zh_inv[(i + k) % qdf]);
                }
            }
        }
    } else
#endif
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) {
        u64 x = xs[i];
        const u64* cs = cd.constants_sigmas.lde_row(i);
        const u64* wr = wires_b.lde_row(i);
        const u64* zr = zs_b.lde_row(i);
        const u64* zn = zs_b.lde_row((i + next_step) % N);
        u64 l0 = fmul(zh[i % qdf], finv(fmul(n_f, fsub(x, 1))));
        auto van = eval_vanishing<BaseOps>(c, x, l0, cs, cs + ncs, wr, zr, zn, zr + nch, pi_hash, betas.data(),
                                           gammas.data(), alphas.data());
        for (size_t ch = 0; ch < nch; ++ch) q[ch][i] = fmul(van[ch], zh_inv[i % qdf]);
    }
    std::vector<std::vector<u64>> chunks;
    for (size_t ch = 0; ch < nch; ++ch) {
        coset_ifft(q[ch], GEN);
        for (size_t m = 0; m < qdf; ++m) chunks.emplace_back(q[ch].begin() + m * n, q[ch].begin() + (m + 1) * n);
    }
    return chunks;
}

static E2 eval_poly_ext(const std::vector<u64>& coeffs, E2 z) {
    E2 acc;
    for (size_t k = coeffs.size(); k-- > 0;) acc = acc * z + E2(coeffs[k]);
    return acc;
}

Proof prove(const CircuitData& cd, const std::vector<std::vector<u64>>& wires, const std::vector<u64>& public_inputs,
            const u64* salts, u64 salt_seed, ProveTrace* trace) {
    const CommonData& c = cd.common;
    size_t n = c.degree(), nch = c.num_challenges;
    unsigned rate_bits = (unsigned)c.fri_config.rate_bits, cap_h = (unsigned)c.fri_config.cap_height;
    size_t N = n << rate_bits;
    bool zk = c.zero_knowledge;
    if (wires.size() != c.num_wires) throw std::runtime_error("wrong wire column count");
    if (public_inputs.size() != c.num_public_inputs) throw std::runtime_error("wrong public input count");
    // ORC_TIMING=1: wall-clock per stage on stderr (where the CPU baseline spends its time)
    const bool timing = std::getenv("ORC_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[orc prove] %-22s %9.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    Proof pr;
    pr.public_inputs = public_inputs;
    Digest pi_hash = hash_no_pad(public_inputs);
    auto salt_ptr = [&](unsigned b) { return salts ? salts + (size_t)b * 4 * N : nullptr; };

    PolyBatch wires_b = batch_from_values(wires, rate_bits, zk, cap_h, salt_ptr(0), salt_seed, 0);
    pr.wires_cap = wires_b.tree.cap;
    lap("wires commit");
    Challenger chal;
    chal.observe_digest(cd.vo.circuit_digest);
    chal.observe_digest(pi_hash);
    chal.observe_cap(pr.wires_cap);
    std::vector<u64> betas, gammas, alphas;
    for (size_t i = 0; i < nch; ++i) betas.push_back(chal.get());
    for (size_t i = 0; i < nch; ++i) gammas.push_back(chal.get());

    auto zs_pp = partial_products_and_zs(cd, wires, betas, gammas);
    lap("partial products");
    if (trace) trace->zs_pp_values = zs_pp;
    PolyBatch zs_b = batch_from_values(zs_pp, rate_bits, zk, cap_h, salt_ptr(1), salt_seed, 1);
    pr.zs_pp_cap = zs_b.tree.cap;
    lap("zs commit");
    chal.observe_cap(pr.zs_pp_cap);
    for (size_t i = 0; i < nch; ++i) alphas.push_back(chal.get());

    auto chunks = compute_quotient_chunks(cd, wires_b, zs_b, pi_hash, betas, gammas, alphas);
    lap("quotient");
    if (trace) trace->quotient_chunks = chunks;
    PolyBatch quot_b = batch_from_coeffs(chunks, rate_bits, zk, cap_h, salt_ptr(2), salt_seed, 2);
    pr.quotient_cap = quot_b.tree.cap;
    lap("quotient commit");
    chal.observe_cap(pr.quotient_cap);
    E2 zeta = chal.get_ext();
    if (epow2k(zeta, (unsigned)c.degree_bits) == E2(1)) throw std::runtime_error("Opening point is in the subgroup.");
    E2 g_h(root_of_unity((unsigned)c.degree_bits));
    E2 zeta_next = g_h * zeta;

    const PolyBatch* oracles[4] = {&cd.constants_sigmas, &wires_b, &zs_b, &quot_b};
    auto eval_batch = [&](const PolyBatch& b, E2 z) {
        std::vector<E2> r(b.ncols);
#pragma omp parallel for schedule(dynamic)
        for (long j = 0; j < (long)b.ncols; ++j) r[j] = eval_poly_ext(b.coeffs[j], z);
        return r;
    };
    auto cs_eval = eval_batch(cd.constants_sigmas, zeta);
    auto zs_eval = eval_batch(zs_b, zeta);
    auto zs_next_eval = eval_batch(zs_b, zeta_next);
    OpeningSet& o = pr.openings;
    o.constants.assign(cs_eval.begin(), cs_eval.begin() + c.num_constants);
    o.plonk_sigmas.assign(cs_eval.begin() + c.num_constants, cs_eval.end());
    o.wires = eval_batch(wires_b, zeta);
    o.plonk_zs.assign(zs_eval.begin(), zs_eval.begin() + nch);
    o.plonk_zs_next.assign(zs_next_eval.begin(), zs_next_eval.begin() + nch);
    o.partial_products.assign(zs_eval.begin() + nch, zs_eval.end());
    o.quotient_polys = eval_batch(quot_b, zeta);
    for (auto* v : {&o.constants, &o.plonk_sigmas, &o.wires, &o.plonk_zs, &o.partial_products, &o.quotient_polys})
        for (E2 e : *v) chal.observe_ext(e);
    for (E2 e : o.plonk_zs_next) chal.observe_ext(e);

    lap("openings");
    // ---- prove_openings ----
    E2 alpha = chal.get_ext();
    std::vector<E2> final_poly(n);
    {
        // batch 0: all polys at zeta. comp[k] = sum_j alpha^j c_j[k]: independent per coefficient index k
        std::vector<E2> comp(n);
        std::vector<E2> apow;
        std::vector<const std::vector<u64>*> cols;
        E2 ap(1);
        for (int t = 0; t < 4; ++t)
            for (size_t j = 0; j < oracles[t]->ncols; ++j) {
                cols.push_back(&oracles[t]->coeffs[j]);
                apow.push_back(ap);
                ap = ap * alpha;
            }
#pragma omp parallel for schedule(static)
        for (long k = 0; k < (long)n; ++k) {
            E2 acc;
            for (size_t j = 0; j < cols.size(); ++j) acc = acc + emul_base(apow[j], (*cols[j])[k]);
            comp[k] = acc;
        }
        auto divide_by_linear = [&](const std::vector<E2>& p, E2 z) {
            std::vector<E2> qv(n);  // n-1 coefficients + a zero pad
            E2 acc;
            for (size_t k = n; k-- > 1;) { acc = acc * z + p[k]; qv[k - 1] = acc; }
            return qv;
        };
        std::vector<E2> q0 = divide_by_linear(comp, zeta);
        // batch 1: Zs at g*zeta
        std::vector<E2> comp1(n);
        ap = E2(1);
        for (size_t j = 0; j < nch; ++j) {
            const auto& cf = zs_b.coeffs[j];
            for (size_t k = 0; k < n; ++k) comp1[k] = comp1[k] + emul_base(ap, cf[k]);
            ap = ap * alpha;
        }
        std::vector<E2> q1 = divide_by_linear(comp1, zeta_next);
        E2 shift = epow(alpha, nch);
        for (size_t k = 0; k < n; ++k) final_poly[k] = q0[k] * shift + q1[k];
    }
    if (trace) trace->final_poly_pre_fri = final_poly;

    lap("fri combine");
    // ---- FRI commit phase ----
    std::vector<E2> coeffs(final_poly);
    coeffs.resize(N);
    std::vector<E2> values(coeffs);
    coset_fft<E2>(values, GEN);
    u64 shift = GEN;
    std::vector<MerkleTree> fri_trees;
    std::vector<E2> fri_betas;
    for (u64 ab : c.reduction_arity_bits) {
        size_t arity = size_t(1) << ab;
        size_t m = values.size();
        unsigned lg = log2_strict(m);
        std::vector<E2> rev(m);
        for (size_t i = 0; i < m; ++i) rev[reverse_bits(i, lg)] = values[i];
        if (trace) trace->fri_layer_values.push_back(rev);
        std::vector<u64> leaves(2 * m);
        for (size_t i = 0; i < m; ++i) { leaves[2 * i] = rev[i].a; leaves[2 * i + 1] = rev[i].b; }
        fri_trees.push_back(merkle_build(std::move(leaves), m / arity, 2 * arity, cap_h));
        pr.commit_phase_caps.push_back(fri_trees.back().cap);
        chal.observe_cap(fri_trees.back().cap);
        E2 beta = chal.get_ext();
        fri_betas.push_back(beta);
        std::vector<E2> folded(coeffs.size() / arity);
        for (size_t k = 0; k < folded.size(); ++k) {
            E2 acc;
            for (size_t i = arity; i-- > 0;) acc = acc * beta + coeffs[k * arity + i];
            folded[k] = acc;
        }
        coeffs = std::move(folded);
        shift = fpow(shift, arity);
        values = coeffs;
        coset_fft<E2>(values, shift);
    }
    coeffs.resize(coeffs.size() >> rate_bits);
    pr.final_poly = coeffs;
    for (E2 e : pr.final_poly) chal.observe_ext(e);

    lap("fri commit phase");
    // ---- PoW (MIN rule) ----
    {
        u64 st0[12];
        for (int i = 0; i < 12; ++i) st0[i] = chal.sponge[i];
        size_t pos = chal.in_buf.size();
        for (size_t i = 0; i < pos; ++i) st0[i] = chal.in_buf[i];
        u64 found = 0;
        bool ok = false;
        const u64 CH = 1 << 14;
        for (u64 base = 0; !ok; base += CH) {
            u64 best = ~u64(0);
#pragma omp parallel for schedule(static) reduction(min : best)
            for (long k = 0; k < (long)CH; ++k) {
                u64 cand = base + k;
                u64 st[12];
                for (int i = 0; i < 12; ++i) st[i] = st0[i];
                st[pos] = cand;
                poseidon_permute(st);
                u64 resp = st[7];
                unsigned lz = resp == 0 ? 64 : (unsigned)__builtin_clzll(resp);
                if (lz >= c.fri_config.proof_of_work_bits && cand < best) best = cand;
            }
            if (best != ~u64(0)) { found = best; ok = true; }
        }
        pr.pow_witness = found;
        chal.observe(found);
        u64 resp = chal.get();
        unsigned lz = resp == 0 ? 64 : (unsigned)__builtin_clzll(resp);
        if (lz < c.fri_config.proof_of_work_bits) throw std::runtime_error("PoW recheck failed");
    }

    lap("pow");
    // ---- query rounds ----
    std::vector<size_t> qidx;
    for (u64 i = 0; i < c.fri_config.num_query_rounds; ++i) qidx.push_back((size_t)(chal.get() % N));
    for (size_t x0 : qidx) {
        FriQueryRound qr;
        size_t x = x0;
        for (int t = 0; t < 4; ++t) {
            const MerkleTree& tr = oracles[t]->tree;
            qr.initial[t].evals.assign(tr.leaf(x), tr.leaf(x) + tr.leaf_width);
            qr.initial[t].path = tr.prove(x);
        }
        for (size_t i = 0; i < fri_trees.size(); ++i) {
            unsigned ab = (unsigned)c.reduction_arity_bits[i];
            size_t ci = x >> ab;
            FriQueryStep st;
            const u64* lf = fri_trees[i].leaf(ci);
            for (size_t k = 0; k < (size_t(1) << ab); ++k) st.evals.push_back(E2(lf[2 * k], lf[2 * k + 1]));
            st.path = fri_trees[i].prove(ci);
            qr.steps.push_back(std::move(st));
            x = ci;
        }
        pr.query_rounds.push_back(std::move(qr));
    }
    if (trace) {
        trace->betas = betas; trace->gammas = gammas; trace->alphas = alphas;
        trace->zeta = zeta; trace->fri_alpha = alpha; trace->fri_betas = fri_betas;
        trace->query_indices = qidx;
    }
    return pr;
}

}  // namespace orc
