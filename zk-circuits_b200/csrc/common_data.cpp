#include "common_data.hpp"

namespace zkb {

namespace {
struct Cursor {
    const uint8_t* p;
    size_t len, pos = 0;
    void need(size_t n) {
        if (len - pos < n) throw ParseError("CommonCircuitData: truncated at byte " + std::to_string(pos));
    }
    uint8_t u8() { need(1); return p[pos++]; }
    u32 le32() { need(4); u32 v = 0; for (int i = 0; i < 4; ++i) v |= (u32)p[pos + i] << (8 * i); pos += 4; return v; }
    u64 le64() { need(8); u64 v = 0; for (int i = 0; i < 8; ++i) v |= (u64)p[pos + i] << (8 * i); pos += 8; return v; }
    bool flag() { uint8_t b = u8(); if (b > 1) throw ParseError("CommonCircuitData: bad bool"); return b != 0; }
    u64 count(size_t elem_bytes) {
        u64 n = le64();
        if (n > (len - pos) / (elem_bytes ? elem_bytes : 1)) throw ParseError("CommonCircuitData: bad length prefix");
        return n;
    }
    std::vector<u64> vec64() { u64 n = count(8); std::vector<u64> v(n); for (auto& x : v) x = le64(); return v; }
};

struct FriCfg { u64 rate_bits, cap_height, num_query_rounds; u32 pow_bits; };
FriCfg read_fri_config(Cursor& c) {
    FriCfg f;
    f.rate_bits = c.le64();
    f.cap_height = c.le64();
    f.num_query_rounds = c.le64();
    f.pow_bits = c.le32();
    uint8_t tag = c.u8();           // reduction strategy: only needed to skip its payload
    if (tag == 0) (void)c.vec64();
    else if (tag == 1) { (void)c.le64(); (void)c.le64(); }
    else if (tag == 2) { if (c.flag()) (void)c.le64(); }
    else throw ParseError("CommonCircuitData: bad FRI reduction strategy tag");
    return f;
}
}  // namespace

CommonData parse_common_data(const uint8_t* p, size_t len) {
    if (!p) throw ParseError("CommonCircuitData: null buffer");
    Cursor c{p, len};
    CommonData d;
    d.num_wires = c.le64();
    d.num_routed_wires = c.le64();
    d.num_constants_cfg = c.le64();
    d.security_bits = c.le64();
    d.num_challenges = c.le64();
    d.max_quotient_degree_factor = c.le64();
    d.use_base_arithmetic_gate = c.flag();
    d.zero_knowledge = c.flag();
    FriCfg f = read_fri_config(c);
    (void)read_fri_config(c);   // FriParams repeats the FriConfig
    d.rate_bits = f.rate_bits;
    d.cap_height = f.cap_height;
    d.num_query_rounds = f.num_query_rounds;
    d.proof_of_work_bits = f.pow_bits;
    d.reduction_arity_bits = c.vec64();
    d.degree_bits = c.le64();
    d.hiding = c.flag();
    d.selector_indices = c.vec64();
    u64 ngroups = c.count(16);
    for (u64 i = 0; i < ngroups; ++i) { u64 a = c.le64(), b = c.le64(); d.groups.push_back({a, b}); }
    d.quotient_degree_factor = c.le64();
    d.num_gate_constraints = c.le64();
    d.num_constants = c.le64();
    d.num_public_inputs = c.le64();
    d.k_is = c.vec64();
    for (u64 k : d.k_is) if (k >= GL_P) throw ParseError("CommonCircuitData: non-canonical k_i");
    d.num_partial_products = c.le64();
    u64 num_lookup_polys = c.le64(), num_lookup_selectors = c.le64(), num_luts = c.le64();
    if (num_lookup_polys || num_lookup_selectors || num_luts) throw UnsupportedError("lookup tables are not supported");
    u64 ngates = c.count(4);
    for (u64 i = 0; i < ngates; ++i) {
        GateInfo g;
        g.tag = c.le32();
        switch (g.tag) {
            case GT_NOOP: case GT_PUBLIC_INPUT: case GT_POSEIDON: case GT_POSEIDON_MDS: break;
            case GT_CONSTANT: case GT_BASE_SUM: case GT_ARITHMETIC: case GT_ARITHMETIC_EXT: case GT_MUL_EXT: case GT_REDUCING:
            case GT_REDUCING_EXT: case GT_EXPONENTIATION: g.param = c.le64(); break;
            case GT_RANDOM_ACCESS: g.param = c.le64(); g.p2 = c.le64(); g.p3 = c.le64(); break;
            case GT_COSET_INTERP: {
                g.param = c.le64(); g.p2 = c.le64();
                g.weights = c.vec64();
                if (g.param == 0 || g.param > 4) throw UnsupportedError("CosetInterpolationGate: subgroup_bits must be 1..4");
                if (g.p2 < 2 || g.weights.size() != (size_t(1) << g.param)) throw ParseError("CosetInterpolationGate: bad degree / weights");
                for (u64 w : g.weights) if (w >= GL_P) throw ParseError("CosetInterpolationGate: non-canonical weight");
                break;
            }
            default: throw UnsupportedError("gate tag " + std::to_string(g.tag) + " is outside the implemented gate set");
        }
        d.gates.push_back(g);
    }
    if (c.pos != len) throw ParseError("CommonCircuitData: trailing bytes");

    // sanity / supported envelope
    if (d.degree_bits == 0 || d.degree_bits > 24) throw UnsupportedError("degree_bits out of range");
    if (d.rate_bits == 0 || d.rate_bits > 4) throw UnsupportedError("rate_bits out of range");
    if (d.num_challenges == 0 || d.num_challenges > 2) throw UnsupportedError("num_challenges must be 1 or 2");
    if (d.quotient_degree_factor != (u64(1) << d.rate_bits))
        throw UnsupportedError("quotient_degree_factor must equal 2^rate_bits");
    if (d.selector_indices.size() != d.gates.size()) throw ParseError("selector / gate count mismatch");
    for (u64 s : d.selector_indices) if (s >= d.groups.size()) throw ParseError("selector index out of range");
    if (d.gates.size() > 32 || d.num_routed_wires > 128 || d.k_is.size() != d.num_routed_wires)
        throw UnsupportedError("circuit exceeds the supported gate / routed-wire count");
    if (d.num_routed_wires > d.num_wires || d.num_wires > 1024) throw ParseError("bad wire counts");
    if ((d.num_routed_wires + d.quotient_degree_factor - 1) / d.quotient_degree_factor != d.num_partial_products + 1)
        throw ParseError("num_partial_products inconsistent with routed wires");
    if (d.groups.size() > d.num_constants) throw ParseError("more selectors than constants");
    if (d.cap_height > d.degree_bits + d.rate_bits) throw ParseError("cap_height exceeds tree height");
    u64 asum = 0;
    for (u64 a : d.reduction_arity_bits) { if (a == 0 || a > 5) throw UnsupportedError("FRI arity out of range"); asum += a; }
    if (asum > d.degree_bits) throw ParseError("FRI reductions exceed degree");
    size_t maxc = 0, totc = 0;
    for (auto& g : d.gates) {
        size_t nc = g.num_constraints();
        maxc = nc > maxc ? nc : maxc;
        totc += nc;
        if (g.param > 4096 || g.p2 > 4096 || g.p3 > 4096) throw ParseError("gate parameter out of range");
        if (g.num_wires() > d.num_wires) throw ParseError("gate tag " + std::to_string(g.tag) + " needs more wires than the circuit has");
        if (g.num_constants() + d.groups.size() > d.num_constants)
            throw ParseError("gate tag " + std::to_string(g.tag) + " needs more constants than the circuit has");
        if (g.tag == GT_RANDOM_ACCESS && (g.param == 0 || g.param > 5)) throw UnsupportedError("RandomAccessGate: bits must be 1..5");
        if ((g.tag == GT_REDUCING || g.tag == GT_REDUCING_EXT || g.tag == GT_EXPONENTIATION) && g.param == 0)
            throw ParseError("gate with zero coefficients / power bits");
    }
    (void)totc;
    size_t ncoset = 0;
    for (auto& g : d.gates) ncoset += g.tag == GT_COSET_INTERP;
    if (ncoset > 1) throw UnsupportedError("more than one CosetInterpolationGate parameter set");
    if (maxc != d.num_gate_constraints) throw ParseError("num_gate_constraints inconsistent with the gate list");
    return d;
}

size_t CommonData::proof_size() const {
    size_t cap = (size_t(1) << cap_height) * 32;
    size_t N_bits = degree_bits + rate_bits;
    size_t s = 3 * cap;
    s += 16 * (num_constants + num_routed_wires + num_wires + 2 * num_challenges + num_challenges * num_partial_products +
               num_quotient_polys());
    s += reduction_arity_bits.size() * cap;
    size_t salt = salt_size();
    size_t widths[4] = {(size_t)(num_constants + num_routed_wires), (size_t)num_wires + salt, num_zs_pp() + salt,
                        num_quotient_polys() + salt};
    size_t per_round = 0;
    for (size_t w : widths) per_round += 8 * w + 1 + 32 * (N_bits - cap_height);
    size_t bits = N_bits;
    for (u64 a : reduction_arity_bits) {
        bits -= a;
        per_round += 16 * (size_t(1) << a) + 1 + 32 * (bits >= cap_height ? bits - cap_height : 0);
    }
    s += num_query_rounds * per_round;
    s += 16 * final_poly_len() + 8 + 8 + 8 * num_public_inputs;
    return s;
}

}  // namespace zkb
