// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// Circuit metadata and proof containers + their wire formats, restating
// qp-plonky2 1.1.1 `CommonCircuitData::{to,from}_bytes`, `VerifierOnlyCircuitData::to_bytes` and
// `ProofWithPublicInputs::{to,from}_bytes` as they are used by the reference at
//   /root/reference/wormhole/prover/src/lib.rs:114-121 (common.bin reader)
//   /root/reference/wormhole/verifier/src/lib.rs:102-106 (verifier.bin reader)
//   /root/reference/wormhole/circuit-builder/src/lib.rs:36-47 (writers)
//   /root/reference/wormhole/aggregator/src/util.rs:22 (ProofWithPublicInputs::from_bytes)
// Byte layouts: SURVEY.md Appendix B.1-B.3, pinned by wormhole/bench-data/{common,verifier,proof}.bin
// and wormhole/aggregator/data/dummy_proof*.bin (parsers must land exactly on EOF).
#pragma once
#include "poseidon.hpp"

namespace orc {

enum GateTag : u32 {  // default gate-serializer tag order (B.1)
    GATE_ARITHMETIC = 0,
    GATE_ARITHMETIC_EXT = 1,
    GATE_BASE_SUM_2 = 2,
    GATE_CONSTANT = 3,
    GATE_COSET_INTERP = 4,
    GATE_EXPONENTIATION = 5,
    GATE_LOOKUP = 6,
    GATE_LOOKUP_TABLE = 7,
    GATE_MUL_EXT = 8,
    GATE_NOOP = 9,
    GATE_POSEIDON_MDS = 10,
    GATE_POSEIDON = 11,
    GATE_PUBLIC_INPUT = 12,
    GATE_RANDOM_ACCESS = 13,
    GATE_REDUCING_EXT = 14,
    GATE_REDUCING = 15,
};

struct Gate {
    u32 tag = GATE_NOOP;
    // Constant: num_consts; BaseSum: num_limbs; Arithmetic / ArithmeticExtension / MulExtension: num_ops;
    // Reducing / ReducingExtension: num_coeffs; Exponentiation: num_power_bits; RandomAccess: bits;
    // CosetInterpolation: subgroup_bits
    u64 param = 0;
    u64 p2 = 0, p3 = 0;       // RandomAccess: num_copies, num_extra_constants; CosetInterpolation: degree, -
    std::vector<u64> weights; // CosetInterpolation: barycentric weights (2^subgroup_bits felts)
    Gate() = default;
    Gate(u32 t, u64 p, u64 q2 = 0, u64 q3 = 0) : tag(t), param(p), p2(q2), p3(q3) {}
    size_t num_constraints() const;
    unsigned degree() const;
    size_t num_constants() const;
};

size_t coset_interp_num_intermediates(const Gate& g);   // (2^subgroup_bits - 2) / (degree - 1)

struct FriConfig {
    u64 rate_bits = 3, cap_height = 4, num_query_rounds = 28;
    u32 proof_of_work_bits = 16;
    u8 strategy_tag = 1;            // 0 Fixed(vec) / 1 ConstantArityBits(a,b) / 2 MinSize(opt)
    std::vector<u64> strategy_args; // payload
};

struct CommonData {
    // CircuitConfig
    u64 num_wires = 135, num_routed_wires = 80, num_constants_cfg = 2, security_bits = 100;
    u64 num_challenges = 2, max_quotient_degree_factor = 8;
    bool use_base_arithmetic_gate = true, zero_knowledge = false;
    FriConfig fri_config;
    // FriParams
    std::vector<u64> reduction_arity_bits;
    u64 degree_bits = 0;
    bool hiding = false;
    // selectors
    std::vector<u64> selector_indices;
    std::vector<std::pair<u64, u64>> groups;
    u64 quotient_degree_factor = 8, num_gate_constraints = 0, num_constants = 0, num_public_inputs = 0;
    std::vector<u64> k_is;
    u64 num_partial_products = 0, num_lookup_polys = 0, num_lookup_selectors = 0;
    std::vector<Gate> gates;

    size_t degree() const { return size_t(1) << degree_bits; }
    size_t lde_size() const { return degree() << fri_config.rate_bits; }
    size_t num_selectors() const { return groups.size(); }
    size_t num_zs_pp() const { return num_challenges * (1 + num_partial_products); }
    size_t num_quotient_polys() const { return num_challenges * quotient_degree_factor; }
    size_t salt_size() const { return zero_knowledge ? 4 : 0; }
    size_t final_poly_len() const {
        u64 s = 0;
        for (u64 a : reduction_arity_bits) s += a;
        return size_t(1) << (degree_bits - s);
    }
};

struct VerifierOnly {
    std::vector<Digest> constants_sigmas_cap;
    Digest circuit_digest{};
};

struct OpeningSet {
    std::vector<E2> constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys;
};

struct InitialTreeProof {
    std::vector<u64> evals;
    std::vector<Digest> path;
};
struct FriQueryStep {
    std::vector<E2> evals;
    std::vector<Digest> path;
};
struct FriQueryRound {
    InitialTreeProof initial[4];
    std::vector<FriQueryStep> steps;
};

struct Proof {
    std::vector<Digest> wires_cap, zs_pp_cap, quotient_cap;
    OpeningSet openings;
    std::vector<std::vector<Digest>> commit_phase_caps;
    std::vector<FriQueryRound> query_rounds;
    std::vector<E2> final_poly;
    u64 pow_witness = 0;
    std::vector<u64> public_inputs;
};

// parsers throw std::runtime_error on malformed input (truncation, trailing bytes, non-canonical felts)
CommonData parse_common(const u8* p, size_t len, size_t* consumed = nullptr);
std::vector<u8> write_common(const CommonData& c);
VerifierOnly parse_verifier_only(const u8* p, size_t len, size_t* consumed);
Proof parse_proof(const u8* p, size_t len, const CommonData& c);
std::vector<u8> write_proof(const Proof& pr);

// FRI reduction schedule for ConstantArityBits(a, b) (A.6)
std::vector<u64> fri_reduction_arity_bits(const FriConfig& cfg, u64 degree_bits);

// circuit_digest = hash_no_pad(flatten(cap) ‖ hash_pad([]) ‖ [degree_bits])  (A.4)
Digest compute_circuit_digest(const std::vector<Digest>& constants_sigmas_cap, u64 degree_bits);

}  // namespace orc
