// Host-side mirror of the parts of qp-plonky2 1.1.1 `CommonCircuitData` the prover needs, read from
// `CommonCircuitData::to_bytes` (= wormhole/generated-bins/common.bin, written at
// /root/reference/wormhole/circuit-builder/src/lib.rs:36-39 and read at
// /root/reference/wormhole/prover/src/lib.rs:114). Byte layout: SURVEY.md Appendix B.1.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "field.cuh"

namespace zkb {

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };
struct UnsupportedError : std::runtime_error { using std::runtime_error::runtime_error; };

// DefaultGateSerializer tags. The first six are the wormhole / voting set; the rest is the recursion set
// `verify_proof` instantiates (wormhole/aggregator/src/circuits/tree.rs:119). Lookup gates: unsupported.
enum : u32 { GT_ARITHMETIC = 0, GT_ARITHMETIC_EXT = 1, GT_BASE_SUM = 2, GT_CONSTANT = 3, GT_COSET_INTERP = 4, GT_EXPONENTIATION = 5,
             GT_MUL_EXT = 8, GT_NOOP = 9, GT_POSEIDON_MDS = 10, GT_POSEIDON = 11, GT_PUBLIC_INPUT = 12, GT_RANDOM_ACCESS = 13,
             GT_REDUCING_EXT = 14, GT_REDUCING = 15 };

struct GateInfo {
    u32 tag = GT_NOOP;
    u64 param = 0;             // see kernels.h GateDesc
    u64 p2 = 0, p3 = 0;
    std::vector<u64> weights;  // CosetInterpolation barycentric weights
    size_t coset_intermediates() const { return ((size_t(1) << param) - 2) / (p2 - 1); }
    size_t num_constraints() const {
        switch (tag) {
            case GT_NOOP: return 0;
            case GT_CONSTANT: return param;
            case GT_PUBLIC_INPUT: return 4;
            case GT_BASE_SUM: return 1 + param;
            case GT_ARITHMETIC: return param;
            case GT_POSEIDON: return 123;
            case GT_ARITHMETIC_EXT: case GT_MUL_EXT: case GT_REDUCING: case GT_REDUCING_EXT: return 2 * param;
            case GT_RANDOM_ACCESS: return p2 * (param + 2) + p3;
            case GT_EXPONENTIATION: return param + 1;
            case GT_COSET_INTERP: return 4 + 4 * coset_intermediates();
            case GT_POSEIDON_MDS: return 24;
            default: return 0;
        }
    }
    size_t num_constants() const {
        switch (tag) {
            case GT_CONSTANT: return param;
            case GT_ARITHMETIC: case GT_ARITHMETIC_EXT: return 2;
            case GT_MUL_EXT: return 1;
            case GT_RANDOM_ACCESS: return p3;
            default: return 0;
        }
    }
    // highest wire index the gate's constraints read, + 1
    size_t num_wires() const {
        switch (tag) {
            case GT_CONSTANT: return param;
            case GT_PUBLIC_INPUT: return 4;
            case GT_BASE_SUM: return 1 + param;
            case GT_ARITHMETIC: return 4 * param;
            case GT_POSEIDON: return 135;
            case GT_ARITHMETIC_EXT: return 8 * param;
            case GT_MUL_EXT: return 6 * param;
            case GT_REDUCING: return 4 + 3 * param;
            case GT_REDUCING_EXT: return 4 + 4 * param;
            case GT_RANDOM_ACCESS: return (2 + (size_t(1) << param)) * p2 + p3 + param * p2;
            case GT_EXPONENTIATION: return 2 + 2 * param;
            case GT_COSET_INTERP: return 1 + 2 * (size_t(1) << param) + 4 + 4 * coset_intermediates() + 2;
            case GT_POSEIDON_MDS: return 48;
            default: return 0;
        }
    }
    bool is_recursion_gate() const {
        return tag == GT_ARITHMETIC_EXT || tag == GT_MUL_EXT || tag == GT_REDUCING || tag == GT_REDUCING_EXT ||
               tag == GT_RANDOM_ACCESS || tag == GT_EXPONENTIATION || tag == GT_COSET_INTERP || tag == GT_POSEIDON_MDS;
    }
};

struct CommonData {
    u64 num_wires = 0, num_routed_wires = 0, num_constants_cfg = 0, security_bits = 0, num_challenges = 0,
        max_quotient_degree_factor = 0;
    bool use_base_arithmetic_gate = false, zero_knowledge = false;
    u64 rate_bits = 0, cap_height = 0, num_query_rounds = 0;
    u32 proof_of_work_bits = 0;
    std::vector<u64> reduction_arity_bits;
    u64 degree_bits = 0;
    bool hiding = false;
    std::vector<u64> selector_indices;
    std::vector<std::pair<u64, u64>> groups;
    u64 quotient_degree_factor = 0, num_gate_constraints = 0, num_constants = 0, num_public_inputs = 0;
    std::vector<u64> k_is;
    u64 num_partial_products = 0;
    std::vector<GateInfo> gates;

    size_t degree() const { return size_t(1) << degree_bits; }
    size_t lde_size() const { return degree() << rate_bits; }
    size_t num_zs_pp() const { return num_challenges * (1 + num_partial_products); }
    size_t num_quotient_polys() const { return num_challenges * quotient_degree_factor; }
    size_t salt_size() const { return (zero_knowledge && hiding) ? 4 : 0; }
    size_t final_poly_len() const {
        u64 s = 0;
        for (u64 a : reduction_arity_bits) s += a;
        return size_t(1) << (degree_bits - s);
    }
    size_t proof_size() const;   // exact ProofWithPublicInputs::to_bytes length
};

// throws ParseError (malformed) or UnsupportedError (valid but outside the implemented gate set / config)
CommonData parse_common_data(const uint8_t* p, size_t len);

}  // namespace zkb
