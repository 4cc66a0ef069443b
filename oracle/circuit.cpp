// ORACLE — TEST INFRASTRUCTURE ONLY (see circuit.hpp header).
#include "circuit.hpp"

namespace orc {

size_t Gate::num_constraints() const {
    switch (tag) {
        case GATE_NOOP: return 0;
        case GATE_CONSTANT: return param;
        case GATE_PUBLIC_INPUT: return 4;
        case GATE_BASE_SUM_2: return 1 + param;
        case GATE_ARITHMETIC: return param;
        case GATE_POSEIDON: return 123;
        // recursion gate set (D = 2)
        case GATE_ARITHMETIC_EXT: case GATE_MUL_EXT: return 2 * param;
        case GATE_REDUCING: case GATE_REDUCING_EXT: return 2 * param;
        case GATE_RANDOM_ACCESS: return p2 * (param + 2) + p3;
        case GATE_EXPONENTIATION: return param + 1;
        case GATE_COSET_INTERP: return 2 + 2 * 2 * coset_interp_num_intermediates(*this) + 2;
        case GATE_POSEIDON_MDS: return 24;
        default: throw std::runtime_error("unsupported gate tag " + std::to_string(tag));
    }
}
size_t coset_interp_num_intermediates(const Gate& g) {
    return ((size_t(1) << g.param) - 2) / (g.p2 - 1);
}
unsigned Gate::degree() const {
    switch (tag) {
        case GATE_NOOP: return 0;
        case GATE_CONSTANT: return 1;
        case GATE_PUBLIC_INPUT: return 1;
        case GATE_BASE_SUM_2: return 2;
        case GATE_ARITHMETIC: return 3;
        case GATE_POSEIDON: return 7;
        case GATE_ARITHMETIC_EXT: case GATE_MUL_EXT: return 3;
        case GATE_REDUCING: case GATE_REDUCING_EXT: return 2;
        case GATE_RANDOM_ACCESS: return (unsigned)param + 1;
        case GATE_EXPONENTIATION: return 4;
        case GATE_COSET_INTERP: return (unsigned)p2;
        case GATE_POSEIDON_MDS: return 1;
        default: throw std::runtime_error("unsupported gate tag " + std::to_string(tag));
    }
}
size_t Gate::num_constants() const {
    switch (tag) {
        case GATE_CONSTANT: return param;
        case GATE_ARITHMETIC: case GATE_ARITHMETIC_EXT: return 2;
        case GATE_MUL_EXT: return 1;
        case GATE_RANDOM_ACCESS: return p3;
        default: return 0;
    }
}

namespace {
struct Reader {
    const u8* p;
    size_t len, pos = 0;
    void need(size_t n) {
        if (pos + n > len) throw std::runtime_error("truncated buffer");
    }
    u8 r8() { need(1); return p[pos++]; }
    u32 r32() {
        need(4);
        u32 v = 0;
        for (int i = 0; i < 4; ++i) v |= (u32)p[pos + i] << (8 * i);
        pos += 4;
        return v;
    }
    u64 r64() {
        need(8);
        u64 v = 0;
        for (int i = 0; i < 8; ++i) v |= (u64)p[pos + i] << (8 * i);
        pos += 8;
        return v;
    }
    u64 felt() {
        u64 v = r64();
        if (v >= P) throw std::runtime_error("non-canonical field element");
        return v;
    }
    E2 ext() { u64 a = felt(); u64 b = felt(); return E2(a, b); }
    Digest digest() { Digest d; for (auto& x : d) x = felt(); return d; }
    bool boolean() {
        u8 b = r8();
        if (b > 1) throw std::runtime_error("bad bool");
        return b;
    }
    std::vector<u64> usize_vec() {
        u64 n = r64();
        if (n > len) throw std::runtime_error("bad vec len");
        std::vector<u64> v(n);
        for (auto& x : v) x = r64();
        return v;
    }
};
struct Writer {
    std::vector<u8> b;
    void w8(u8 v) { b.push_back(v); }
    void w32(u32 v) { for (int i = 0; i < 4; ++i) b.push_back((u8)(v >> (8 * i))); }
    void w64(u64 v) { for (int i = 0; i < 8; ++i) b.push_back((u8)(v >> (8 * i))); }
    void ext(E2 e) { w64(e.a); w64(e.b); }
    void digest(const Digest& d) { for (u64 x : d) w64(x); }
    void usize_vec(const std::vector<u64>& v) { w64(v.size()); for (u64 x : v) w64(x); }
};

FriConfig read_fri_config(Reader& r) {
    FriConfig f;
    f.rate_bits = r.r64();
    f.cap_height = r.r64();
    f.num_query_rounds = r.r64();
    f.proof_of_work_bits = r.r32();
    f.strategy_tag = r.r8();
    if (f.strategy_tag == 0) {
        f.strategy_args = r.usize_vec();
    } else if (f.strategy_tag == 1) {
        f.strategy_args = {r.r64(), r.r64()};
    } else if (f.strategy_tag == 2) {
        u8 some = r.r8();
        f.strategy_args.clear();
        if (some) f.strategy_args.push_back(r.r64());
    } else {
        throw std::runtime_error("bad FRI strategy tag");
    }
    return f;
}
void write_fri_config(Writer& w, const FriConfig& f) {
    w.w64(f.rate_bits);
    w.w64(f.cap_height);
    w.w64(f.num_query_rounds);
    w.w32(f.proof_of_work_bits);
    w.w8(f.strategy_tag);
    if (f.strategy_tag == 0) {
        w.usize_vec(f.strategy_args);
    } else if (f.strategy_tag == 1) {
        w.w64(f.strategy_args.at(0));
        w.w64(f.strategy_args.at(1));
    } else {
        w.w8(f.strategy_args.empty() ? 0 : 1);
        if (!f.strategy_args.empty()) w.w64(f.strategy_args[0]);
    }
}
}  // namespace

CommonData parse_common(const u8* p, size_t len, size_t* consumed) {
    Reader r{p, len};
    CommonData c;
    c.num_wires = r.r64();
    c.num_routed_wires = r.r64();
    c.num_constants_cfg = r.r64();
    c.security_bits = r.r64();
    c.num_challenges = r.r64();
    c.max_quotient_degree_factor = r.r64();
    c.use_base_arithmetic_gate = r.boolean();
    c.zero_knowledge = r.boolean();
    c.fri_config = read_fri_config(r);
    FriConfig again = read_fri_config(r);  // FriParams embeds the config a second time
    (void)again;
    c.reduction_arity_bits = r.usize_vec();
    c.degree_bits = r.r64();
    c.hiding = r.boolean();
    c.selector_indices = r.usize_vec();
    u64 ng = r.r64();
    if (ng > len) throw std::runtime_error("bad groups len");
    for (u64 i = 0; i < ng; ++i) {
        u64 a = r.r64(), b = r.r64();
        c.groups.push_back({a, b});
    }
    c.quotient_degree_factor = r.r64();
    c.num_gate_constraints = r.r64();
    c.num_constants = r.r64();
    c.num_public_inputs = r.r64();
    u64 nk = r.r64();
    if (nk > len) throw std::runtime_error("bad k_is len");
    for (u64 i = 0; i < nk; ++i) c.k_is.push_back(r.felt());
    c.num_partial_products = r.r64();
    c.num_lookup_polys = r.r64();
    c.num_lookup_selectors = r.r64();
    u64 nluts = r.r64();
    if (nluts != 0 || c.num_lookup_polys != 0) throw std::runtime_error("lookup tables unsupported");
    u64 ngates = r.r64();
    if (ngates > len) throw std::runtime_error("bad gates len");
    for (u64 i = 0; i < ngates; ++i) {
        Gate g;
        g.tag = r.r32();
        switch (g.tag) {
            case GATE_NOOP: case GATE_PUBLIC_INPUT: case GATE_POSEIDON: break;
            case GATE_CONSTANT: case GATE_BASE_SUM_2: case GATE_ARITHMETIC: g.param = r.r64(); break;
            case GATE_POSEIDON_MDS: break;
            case GATE_ARITHMETIC_EXT: case GATE_MUL_EXT: case GATE_REDUCING: case GATE_REDUCING_EXT:
            case GATE_EXPONENTIATION: g.param = r.r64(); break;
            case GATE_RANDOM_ACCESS: g.param = r.r64(); g.p2 = r.r64(); g.p3 = r.r64(); break;
            case GATE_COSET_INTERP: {
                g.param = r.r64(); g.p2 = r.r64();
                u64 nw = r.r64();
                if (g.param > 8 || nw != (u64(1) << g.param) || g.p2 < 2) throw std::runtime_error("bad CosetInterpolationGate");
                for (u64 k = 0; k < nw; ++k) g.weights.push_back(r.felt());
                break;
            }
            default: throw std::runtime_error("unsupported gate tag " + std::to_string(g.tag));
        }
        c.gates.push_back(g);
    }
    if (c.selector_indices.size() != c.gates.size()) throw std::runtime_error("selector/gate count mismatch");
    if (consumed) *consumed = r.pos;
    else if (r.pos != len) throw std::runtime_error("trailing bytes after CommonCircuitData");
    return c;
}

std::vector<u8> write_common(const CommonData& c) {
    Writer w;
    w.w64(c.num_wires); w.w64(c.num_routed_wires); w.w64(c.num_constants_cfg); w.w64(c.security_bits);
    w.w64(c.num_challenges); w.w64(c.max_quotient_degree_factor);
    w.w8(c.use_base_arithmetic_gate); w.w8(c.zero_knowledge);
    write_fri_config(w, c.fri_config);
    write_fri_config(w, c.fri_config);
    w.usize_vec(c.reduction_arity_bits);
    w.w64(c.degree_bits);
    w.w8(c.hiding);
    w.usize_vec(c.selector_indices);
    w.w64(c.groups.size());
    for (auto& g : c.groups) { w.w64(g.first); w.w64(g.second); }
    w.w64(c.quotient_degree_factor); w.w64(c.num_gate_constraints); w.w64(c.num_constants); w.w64(c.num_public_inputs);
    w.usize_vec(c.k_is);
    w.w64(c.num_partial_products); w.w64(c.num_lookup_polys); w.w64(c.num_lookup_selectors);
    w.w64(0);  // luts.len
    w.w64(c.gates.size());
    for (auto& g : c.gates) {
        w.w32(g.tag);
        switch (g.tag) {
            case GATE_CONSTANT: case GATE_BASE_SUM_2: case GATE_ARITHMETIC: case GATE_ARITHMETIC_EXT: case GATE_MUL_EXT:
            case GATE_REDUCING: case GATE_REDUCING_EXT: case GATE_EXPONENTIATION: w.w64(g.param); break;
            case GATE_RANDOM_ACCESS: w.w64(g.param); w.w64(g.p2); w.w64(g.p3); break;
            case GATE_COSET_INTERP: w.w64(g.param); w.w64(g.p2); w.usize_vec(g.weights); break;
            default: break;
        }
    }
    return w.b;
}

VerifierOnly parse_verifier_only(const u8* p, size_t len, size_t* consumed) {
    Reader r{p, len};
    VerifierOnly v;
    u64 cap_height = r.r64();
    if (cap_height > 20) throw std::runtime_error("bad cap height");
    for (u64 i = 0; i < (u64(1) << cap_height); ++i) v.constants_sigmas_cap.push_back(r.digest());
    v.circuit_digest = r.digest();
    if (consumed) *consumed = r.pos;
    return v;
}

Proof parse_proof(const u8* p, size_t len, const CommonData& c) {
    Reader r{p, len};
    Proof pr;
    size_t cap_n = size_t(1) << c.fri_config.cap_height;
    auto cap = [&]() { std::vector<Digest> v(cap_n); for (auto& d : v) d = r.digest(); return v; };
    auto exts = [&](size_t n) { std::vector<E2> v(n); for (auto& e : v) e = r.ext(); return v; };
    auto path = [&]() { u8 n = r.r8(); std::vector<Digest> v(n); for (auto& d : v) d = r.digest(); return v; };
    pr.wires_cap = cap();
    pr.zs_pp_cap = cap();
    pr.quotient_cap = cap();
    pr.openings.constants = exts(c.num_constants);
    pr.openings.plonk_sigmas = exts(c.num_routed_wires);
    pr.openings.wires = exts(c.num_wires);
    pr.openings.plonk_zs = exts(c.num_challenges);
    pr.openings.plonk_zs_next = exts(c.num_challenges);
    pr.openings.partial_products = exts(c.num_challenges * c.num_partial_products);
    pr.openings.quotient_polys = exts(c.num_quotient_polys());
    for (size_t i = 0; i < c.reduction_arity_bits.size(); ++i) pr.commit_phase_caps.push_back(cap());
    size_t salt = c.hiding ? 4 : 0;
    size_t widths[4] = {c.num_constants + c.num_routed_wires, c.num_wires + salt, c.num_zs_pp() + salt,
                        c.num_quotient_polys() + salt};
    for (u64 q = 0; q < c.fri_config.num_query_rounds; ++q) {
        FriQueryRound qr;
        for (int t = 0; t < 4; ++t) {
            qr.initial[t].evals.resize(widths[t]);
            for (auto& x : qr.initial[t].evals) x = r.felt();
            qr.initial[t].path = path();
        }
        for (u64 ab : c.reduction_arity_bits) {
            FriQueryStep st;
            st.evals = exts(size_t(1) << ab);
            st.path = path();
            qr.steps.push_back(std::move(st));
        }
        pr.query_rounds.push_back(std::move(qr));
    }
    pr.final_poly = exts(c.final_poly_len());
    pr.pow_witness = r.felt();
    u64 npi = r.r64();
    if (npi > len) throw std::runtime_error("bad public input count");
    for (u64 i = 0; i < npi; ++i) pr.public_inputs.push_back(r.felt());
    if (r.pos != len) throw std::runtime_error("trailing bytes after proof");
    return pr;
}

std::vector<u8> write_proof(const Proof& pr) {
    Writer w;
    auto cap = [&](const std::vector<Digest>& c) { for (auto& d : c) w.digest(d); };
    auto exts = [&](const std::vector<E2>& v) { for (auto& e : v) w.ext(e); };
    auto path = [&](const std::vector<Digest>& v) { w.w8((u8)v.size()); for (auto& d : v) w.digest(d); };
    cap(pr.wires_cap); cap(pr.zs_pp_cap); cap(pr.quotient_cap);
    exts(pr.openings.constants); exts(pr.openings.plonk_sigmas); exts(pr.openings.wires);
    exts(pr.openings.plonk_zs); exts(pr.openings.plonk_zs_next); exts(pr.openings.partial_products);
    exts(pr.openings.quotient_polys);
    for (auto& c : pr.commit_phase_caps) cap(c);
    for (auto& qr : pr.query_rounds) {
        for (int t = 0; t < 4; ++t) {
            for (u64 x : qr.initial[t].evals) w.w64(x);
            path(qr.initial[t].path);
        }
        for (auto& st : qr.steps) { exts(st.evals); path(st.path); }
    }
    exts(pr.final_poly);
    w.w64(pr.pow_witness);
    w.w64(pr.public_inputs.size());
    for (u64 x : pr.public_inputs) w.w64(x);
    return w.b;
}

std::vector<u64> fri_reduction_arity_bits(const FriConfig& cfg, u64 degree_bits) {
    std::vector<u64> out;
    if (cfg.strategy_tag == 0) return cfg.strategy_args;
    if (cfg.strategy_tag != 1) throw std::runtime_error("unsupported FRI reduction strategy");
    u64 arity_bits = cfg.strategy_args.at(0), final_poly_bits = cfg.strategy_args.at(1);
    while (degree_bits > final_poly_bits && degree_bits + cfg.rate_bits - arity_bits >= cfg.cap_height) {
        out.push_back(arity_bits);
        degree_bits -= arity_bits;
    }
    return out;
}

Digest compute_circuit_digest(const std::vector<Digest>& cap, u64 degree_bits) {
    std::vector<u64> parts;
    for (auto& d : cap) for (u64 x : d) parts.push_back(x);
    Digest ds = hash_pad({});
    for (u64 x : ds) parts.push_back(x);
    parts.push_back(degree_bits);
    return hash_no_pad(parts);
}

}  // namespace orc
