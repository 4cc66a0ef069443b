// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
// Flat C entry points over the oracle so tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs can drive it through ctypes. All matrices are column-major [col][row] u64 unless noted.
#include "prover.hpp"
#include "vec_ops.hpp"
#include <cstring>
#include <chrono>
#include <omp.h>

using namespace orc;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) { g_err = e.what(); return -1; }
std::vector<std::vector<u64>> cols_from(const u64* p, size_t ncols, size_t n) {
    std::vector<std::vector<u64>> v(ncols);
    for (size_t c = 0; c < ncols; ++c) v[c].assign(p + c * n, p + (c + 1) * n);
    return v;
}
void cols_to(const std::vector<std::vector<u64>>& v, u64* out) {
    for (size_t c = 0; c < v.size(); ++c) std::memcpy(out + c * v[c].size(), v[c].data(), v[c].size() * 8);
}
}  // namespace

struct OrcCircuit {
    CircuitData cd;
    std::vector<u8> common_bytes;
};
struct OrcSynth {
    SynthCircuit sc;
    std::vector<u8> common_bytes;
};
struct OrcTrace {
    ProveTrace t;
};

extern "C" {
#pragma GCC visibility push(default)   // built with -fvisibility=hidden: only these entry points are exported

const char* orc_last_error() { return g_err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }
// 1: the optimised host forms (CPU-baseline arm); 0: the readable restatement (default: what the parity tests use)
void orc_set_fast(int on) { g_fast_poseidon = on != 0; }
int orc_get_fast() { return g_fast_poseidon ? 1 : 0; }

// ---- field / Poseidon ----
void orc_round_constants(u64* out360) { std::memcpy(out360, poseidon_round_constants(), 360 * 8); }
void orc_poseidon_permute(u64* states, size_t count) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)count; ++i) poseidon_permute(states + 12 * i);
}
void orc_hash_no_pad(const u64* v, size_t len, u64* out4) {
    Digest d = hash_no_pad(v, len);
    std::memcpy(out4, d.data(), 32);
}
void orc_hash_pad(const u64* v, size_t len, u64* out4) {
    Digest d = hash_pad(std::vector<u64>(v, v + len));
    std::memcpy(out4, d.data(), 32);
}
void orc_two_to_one(const u64* l, const u64* r, u64* out4) {
    Digest a, b;
    std::memcpy(a.data(), l, 32);
    std::memcpy(b.data(), r, 32);
    Digest d = two_to_one(a, b);
    std::memcpy(out4, d.data(), 32);
}
u64 orc_fmul(u64 a, u64 b) { return fmul(a, b); }
u64 orc_finv(u64 a) { return finv(a); }
u64 orc_root_of_unity(unsigned k) { return root_of_unity(k); }

// ---- NTT / LDE: values [ncols][n] -> coeffs [ncols][n] and LDE leaves in the reference's leaf
// order: lde_out [ncols][n<<rate_bits] with lde_out[c][l] = P_c(g * w^bitrev(l))  (A.1, A.6)
int orc_lde_batch(const u64* values, size_t ncols, size_t n, unsigned rate_bits, int from_coeffs,
                  u64* coeffs_out, u64* lde_out) {
    try {
        size_t N = n << rate_bits;
        unsigned lg = log2_strict(N);
#pragma omp parallel for schedule(dynamic)
        for (long c = 0; c < (long)ncols; ++c) {
            std::vector<u64> v(values + c * n, values + (c + 1) * n);
            if (!from_coeffs) ifft(v);
            if (coeffs_out) std::memcpy(coeffs_out + c * n, v.data(), n * 8);
            if (lde_out) {
                std::vector<u64> e = lde_coset<u64>(v, rate_bits);
                for (size_t i = 0; i < N; ++i) lde_out[c * N + reverse_bits(i, lg)] = e[i];
            }
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// forward/inverse plain NTT on ext or base columns (natural order), for unit tests
int orc_ntt(u64* data, size_t ncols, size_t n, int inverse) {
    try {
        for (size_t c = 0; c < ncols; ++c) {
            std::vector<u64> v(data + c * n, data + (c + 1) * n);
            if (inverse) ifft(v); else fft(v);
            std::memcpy(data + c * n, v.data(), n * 8);
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- Merkle: leaves column-major [width][num_leaves]; digests_out = all levels concatenated
// (level 0 first: num_leaves digests, then num_leaves/2 ... down to the cap level), cap_out = 2^cap_height digests
int orc_merkle_commit(const u64* leaves_colmajor, size_t width, size_t num_leaves, unsigned cap_height,
                      u64* digests_out, u64* cap_out) {
    try {
        std::vector<u64> rows(num_leaves * width);
        for (size_t c = 0; c < width; ++c)
            for (size_t l = 0; l < num_leaves; ++l) rows[l * width + c] = leaves_colmajor[c * num_leaves + l];
        MerkleTree t = merkle_build(std::move(rows), num_leaves, width, cap_height);
        if (digests_out) {
            size_t off = 0;
            for (auto& lv : t.levels) { std::memcpy(digests_out + off, lv.data(), lv.size() * 32); off += lv.size() * 4; }
        }
        if (cap_out) std::memcpy(cap_out, t.cap.data(), t.cap.size() * 32);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- verifier ----
// verifier_only: 2^cap_height*4 cap felts then 4 digest felts (the VerifierOnlyCircuitData payload)
int orc_verify(const u8* common, size_t common_len, const u64* cap, size_t cap_len, const u64* digest,
               const u8* proof, size_t proof_len) {
    try {
        CommonData c = parse_common(common, common_len);
        VerifierOnly vo;
        for (size_t i = 0; i < cap_len; ++i) vo.constants_sigmas_cap.push_back({cap[4 * i], cap[4 * i + 1], cap[4 * i + 2], cap[4 * i + 3]});
        vo.circuit_digest = {digest[0], digest[1], digest[2], digest[3]};
        Proof pr = parse_proof(proof, proof_len, c);
        std::string r = verify_proof(c, vo, pr);
        if (!r.empty()) { g_err = r; return 1; }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// verify against a VerifierCircuitData blob (verifier-only ‖ common), e.g. bench-data/verifier.bin
int orc_verify_with_verifier_bin(const u8* vbin, size_t vlen, const u8* proof, size_t proof_len) {
    try {
        size_t used = 0;
        VerifierOnly vo = parse_verifier_only(vbin, vlen, &used);
        CommonData c = parse_common(vbin + used, vlen - used);
        Proof pr = parse_proof(proof, proof_len, c);
        std::string r = verify_proof(c, vo, pr);
        if (!r.empty()) { g_err = r; return 1; }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// parse + re-serialize (round trip) helpers for the wire formats
long orc_common_roundtrip(const u8* common, size_t len, u8* out, size_t cap) {
    try {
        auto b = write_common(parse_common(common, len));
        if (b.size() > cap) { g_err = "buffer too small"; return -1; }
        std::memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (const std::exception& e) { return fail(e); }
}
long orc_proof_roundtrip(const u8* common, size_t clen, const u8* proof, size_t plen, u8* out, size_t cap) {
    try {
        CommonData c = parse_common(common, clen);
        auto b = write_proof(parse_proof(proof, plen, c));
        if (b.size() > cap) { g_err = "buffer too small"; return -1; }
        std::memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (const std::exception& e) { return fail(e); }
}
// info[0..]: degree_bits, num_wires, num_routed, num_constants, num_challenges, num_partial_products,
// quotient_degree_factor, zk, rate_bits, cap_height, num_query_rounds, pow_bits, num_public_inputs, num_gates, n_arity
int orc_common_info(const u8* common, size_t len, u64* info, u64* arity_bits_out) {
    try {
        CommonData c = parse_common(common, len);
        u64 v[] = {c.degree_bits, c.num_wires, c.num_routed_wires, c.num_constants, c.num_challenges,
                   c.num_partial_products, c.quotient_degree_factor, (u64)c.zero_knowledge, c.fri_config.rate_bits,
                   c.fri_config.cap_height, c.fri_config.num_query_rounds, c.fri_config.proof_of_work_bits,
                   c.num_public_inputs, c.gates.size(), c.reduction_arity_bits.size()};
        std::memcpy(info, v, sizeof(v));
        if (arity_bits_out) for (size_t i = 0; i < c.reduction_arity_bits.size(); ++i) arity_bits_out[i] = c.reduction_arity_bits[i];
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// Fiat-Shamir replay. out: betas[nch] gammas[nch] alphas[nch] zeta[2] fri_alpha[2] pow_response[1] then query indices
int orc_challenges(const u8* vbin, size_t vlen, const u8* proof, size_t proof_len, u64* out, size_t out_cap) {
    try {
        size_t used = 0;
        VerifierOnly vo = parse_verifier_only(vbin, vlen, &used);
        CommonData c = parse_common(vbin + used, vlen - used);
        Proof pr = parse_proof(proof, proof_len, c);
        Challenges ch = derive_challenges(c, vo, pr);
        std::vector<u64> v;
        v.insert(v.end(), ch.betas.begin(), ch.betas.end());
        v.insert(v.end(), ch.gammas.begin(), ch.gammas.end());
        v.insert(v.end(), ch.alphas.begin(), ch.alphas.end());
        v.push_back(ch.zeta.a); v.push_back(ch.zeta.b);
        v.push_back(ch.fri_alpha.a); v.push_back(ch.fri_alpha.b);
        v.push_back(ch.pow_response);
        for (size_t q : ch.query_indices) v.push_back(q);
        if (v.size() > out_cap) { g_err = "buffer too small"; return -1; }
        std::memcpy(out, v.data(), v.size() * 8);
        return (int)v.size();
    } catch (const std::exception& e) { return fail(e); }
}

// ---- synthetic circuits ----
OrcSynth* orc_synth_make(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith,
                         size_t n_const, size_t num_public_inputs, u64 seed) {
    try {
        SynthSpec sp;
        sp.min_degree_bits = min_degree_bits; sp.zk = zk != 0;
        sp.n_poseidon = n_poseidon; sp.n_base_sum = n_base_sum; sp.n_arith = n_arith; sp.n_const = n_const;
        sp.num_public_inputs = num_public_inputs; sp.seed = seed;
        auto* s = new OrcSynth{make_synth_circuit(sp), {}};
        s->common_bytes = write_common(s->sc.common);
        return s;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
// recursion-shaped circuit: counts[8] = rows of ArithmeticExtension, MulExtension, Reducing, ReducingExtension,
// RandomAccess, Exponentiation, CosetInterpolation, PoseidonMds
OrcSynth* orc_synth_make_recursion(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith, size_t n_const,
                                   size_t num_public_inputs, u64 seed, const size_t* counts) {
    try {
        SynthSpec sp;
        sp.min_degree_bits = min_degree_bits; sp.zk = zk != 0;
        sp.n_poseidon = n_poseidon; sp.n_base_sum = n_base_sum; sp.n_arith = n_arith; sp.n_const = n_const;
        sp.num_public_inputs = num_public_inputs; sp.seed = seed;
        sp.n_arith_ext = counts[0]; sp.n_mul_ext = counts[1]; sp.n_reducing = counts[2]; sp.n_reducing_ext = counts[3];
        sp.n_random_access = counts[4]; sp.n_exp = counts[5]; sp.n_coset = counts[6]; sp.n_mds = counts[7];
        auto* s = new OrcSynth{make_synth_circuit(sp), {}};
        s->common_bytes = write_common(s->sc.common);
        return s;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
void orc_synth_free(OrcSynth* s) { delete s; }
size_t orc_synth_common_len(const OrcSynth* s) { return s->common_bytes.size(); }
void orc_synth_common(const OrcSynth* s, u8* out) { std::memcpy(out, s->common_bytes.data(), s->common_bytes.size()); }
size_t orc_synth_degree(const OrcSynth* s) { return s->sc.common.degree(); }
void orc_synth_const_sigma_values(const OrcSynth* s, u64* out) { cols_to(s->sc.const_sigma_values, out); }
void orc_synth_wires(const OrcSynth* s, u64* out) { cols_to(s->sc.wires, out); }
void orc_synth_public_inputs(const OrcSynth* s, u64* out) { std::memcpy(out, s->sc.public_inputs.data(), s->sc.public_inputs.size() * 8); }
int orc_synth_check(const OrcSynth* s) {
    std::string r = check_witness(s->sc);
    if (!r.empty()) { g_err = r; return 1; }
    return 0;
}

// ---- circuit context + prover ----
OrcCircuit* orc_circuit_create(const u8* common, size_t len, const u64* const_sigma_values) {
    try {
        CommonData c = parse_common(common, len);
        auto* oc = new OrcCircuit{circuit_from_values(c, cols_from(const_sigma_values, c.num_constants + c.num_routed_wires, c.degree())),
                                  std::vector<u8>(common, common + len)};
        return oc;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
void orc_circuit_free(OrcCircuit* c) { delete c; }
void orc_circuit_cap(const OrcCircuit* c, u64* out) { std::memcpy(out, c->cd.vo.constants_sigmas_cap.data(), c->cd.vo.constants_sigmas_cap.size() * 32); }
void orc_circuit_digest(const OrcCircuit* c, u64* out4) { std::memcpy(out4, c->cd.vo.circuit_digest.data(), 32); }
void orc_circuit_const_sigma_coeffs(const OrcCircuit* c, u64* out) { cols_to(c->cd.constants_sigmas.coeffs, out); }

OrcTrace* orc_trace_new() { return new OrcTrace(); }
void orc_trace_free(OrcTrace* t) { delete t; }
// which: 0 zs_pp_values [20][n], 1 quotient_chunks [16][n], 2 final_poly_pre_fri [n][2] (interleaved),
//        3+i: FRI layer i committed values (bit-reversed order) [m][2] interleaved
long orc_trace_get(const OrcTrace* t, int which, u64* out, size_t cap_words) {
    std::vector<u64> v;
    if (which == 0) for (auto& c : t->t.zs_pp_values) v.insert(v.end(), c.begin(), c.end());
    else if (which == 1) for (auto& c : t->t.quotient_chunks) v.insert(v.end(), c.begin(), c.end());
    else if (which == 2) for (E2 e : t->t.final_poly_pre_fri) { v.push_back(e.a); v.push_back(e.b); }
    else if (which >= 3 && (size_t)(which - 3) < t->t.fri_layer_values.size())
        for (E2 e : t->t.fri_layer_values[which - 3]) { v.push_back(e.a); v.push_back(e.b); }
    else return -1;
    if (out) {
        if (v.size() > cap_words) return -1;
        std::memcpy(out, v.data(), v.size() * 8);
    }
    return (long)v.size();
}
// challenges: betas[2] gammas[2] alphas[2] zeta[2] fri_alpha[2] then fri_betas (2 each)
long orc_trace_challenges(const OrcTrace* t, u64* out, size_t cap_words) {
    std::vector<u64> v;
    v.insert(v.end(), t->t.betas.begin(), t->t.betas.end());
    v.insert(v.end(), t->t.gammas.begin(), t->t.gammas.end());
    v.insert(v.end(), t->t.alphas.begin(), t->t.alphas.end());
    v.push_back(t->t.zeta.a); v.push_back(t->t.zeta.b);
    v.push_back(t->t.fri_alpha.a); v.push_back(t->t.fri_alpha.b);
    for (E2 e : t->t.fri_betas) { v.push_back(e.a); v.push_back(e.b); }
    if (v.size() > cap_words) return -1;
    std::memcpy(out, v.data(), v.size() * 8);
    return (long)v.size();
}

// wires [num_wires][n]; salts NULL or [3][4][8n]; returns proof length or <0
long orc_prove(const OrcCircuit* c, const u64* wires, const u64* public_inputs, size_t n_pi, const u64* salts,
               u64 salt_seed, u8* proof_out, size_t cap, OrcTrace* trace) {
    try {
        const CommonData& cm = c->cd.common;
        Proof pr = prove(c->cd, cols_from(wires, cm.num_wires, cm.degree()), std::vector<u64>(public_inputs, public_inputs + n_pi),
                         salts, salt_seed, trace ? &trace->t : nullptr);
        auto b = write_proof(pr);
        if (b.size() > cap) { g_err = "buffer too small"; return -2; }
        std::memcpy(proof_out, b.data(), b.size());
        return (long)b.size();
    } catch (const std::exception& e) { return fail(e); }
}

// stage-level entry points for parity tests
int orc_partial_products(const OrcCircuit* c, const u64* wires, const u64* betas, const u64* gammas, u64* out) {
    try {
        const CommonData& cm = c->cd.common;
        auto r = partial_products_and_zs(c->cd, cols_from(wires, cm.num_wires, cm.degree()),
                                         std::vector<u64>(betas, betas + cm.num_challenges), std::vector<u64>(gammas, gammas + cm.num_challenges));
        cols_to(r, out);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// wires [num_wires][n] and zs_pp [num_zs_pp][n] as VALUES over H; unsalted; out: [nch*qdf][n] coefficient chunks
int orc_quotient(const OrcCircuit* c, const u64* wires, const u64* zs_pp, const u64* public_inputs, size_t n_pi,
                 const u64* betas, const u64* gammas, const u64* alphas, u64* out) {
    try {
        const CommonData& cm = c->cd.common;
        unsigned rb = (unsigned)cm.fri_config.rate_bits, ch = (unsigned)cm.fri_config.cap_height;
        PolyBatch wb = batch_from_values(cols_from(wires, cm.num_wires, cm.degree()), rb, false, ch, nullptr, 0, 0);
        PolyBatch zb = batch_from_values(cols_from(zs_pp, cm.num_zs_pp(), cm.degree()), rb, false, ch, nullptr, 0, 1);
        Digest pih = hash_no_pad(public_inputs, n_pi);
        size_t nch = cm.num_challenges;
        auto r = compute_quotient_chunks(c->cd, wb, zb, pih, std::vector<u64>(betas, betas + nch),
                                         std::vector<u64>(gammas, gammas + nch), std::vector<u64>(alphas, alphas + nch));
        cols_to(r, out);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

u64 orc_salt_value(u64 seed, unsigned batch, unsigned s, u64 leaf) { return salt_value(seed, batch, s, leaf); }

// the 4-lane field operations of the CPU-baseline arm's gate evaluator (vec_ops.hpp), for a direct check against big-integer
// arithmetic on corner values: a, b [count] canonical (count a multiple of 4) -> add, sub, mul, mds-free sbox input products
int orc_vecops_check(const u64* a, const u64* b, size_t count, u64 c, u64* add, u64* sub, u64* mul, u64* mulc) {
#if defined(__AVX2__)
    if (count % 4) { g_err = "count must be a multiple of 4"; return -1; }
    for (size_t i = 0; i < count; i += 4) {
        V4 x = VecOps::lanes(a[i], a[i + 1], a[i + 2], a[i + 3]), y = VecOps::lanes(b[i], b[i + 1], b[i + 2], b[i + 3]);
        VecOps::store(VecOps::add(x, y), add + i);
        VecOps::store(VecOps::sub(x, y), sub + i);
        VecOps::store(VecOps::mul(x, y), mul + i);
        VecOps::store(VecOps::mulc(x, c), mulc + i);
    }
    return 0;
#else
    g_err = "built without AVX2";
    return -1;
#endif
}
// mds_layer<VecOps> on four states [4][12] (row = one state) against mds_layer<BaseOps>
int orc_vecops_mds(const u64* states, u64* out) {
#if defined(__AVX2__)
    V4 st[12];
    for (int j = 0; j < 12; ++j) st[j] = VecOps::lanes(states[j], states[12 + j], states[24 + j], states[36 + j]);
    mds_layer<VecOps>(st);
    for (int j = 0; j < 12; ++j) {
        u64 l[4];
        VecOps::store(st[j], l);
        for (int k = 0; k < 4; ++k) out[12 * k + j] = l[k];
    }
    return 0;
#else
    g_err = "built without AVX2";
    return -1;
#endif
}

#pragma GCC visibility pop
}  // extern "C"
