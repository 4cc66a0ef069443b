// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called by the product (libzkb200.so).
//
// CPU restatement of the qp-plonky2 1.1.1 prover stages that run behind
// `ProverCircuitData::prove` (/root/reference/wormhole/prover/src/lib.rs:233-237) and
// `CircuitData::prove` (/root/reference/wormhole/aggregator/src/circuits/tree.rs:136):
// PolynomialBatch::{from_values,from_coeffs}, wires_permutation_partial_products_and_zs,
// compute_quotient_polys, OpeningSet::new, PolynomialBatch::prove_openings, fri_committed_trees,
// fri_proof_of_work, fri_prover_query_rounds.
//
// PARITY STATUS: the crate source is absent and no reference test pins prover intermediates, so
// these stages are pinned *through the verifier*: every proof this prover emits must be accepted by
// oracle/verifier.cpp, which itself is pinned on the reference's bench-data fixture. Every prover
// message is a unique function of (wires, salts, PoW witness), so acceptance plus
// exact integer arithmetic implies identical bytes. The two free knobs are fixed here as:
//   salts  = caller-supplied array, or the documented SplitMix64 generator below;
//   PoW    = MIN rule (smallest valid witness; the CPU prover's result with one rayon thread).
#pragma once
#include "circuit.hpp"
#include "ntt.hpp"
#include "verifier.hpp"

namespace orc {

// Documented salt generator (shared definition with the product): value stored at leaf `leaf`
// (bit-reversed storage index), salt column s in 0..3, batch b in {0 wires, 1 zs_pp, 2 quotient}.
inline u64 splitmix64_finalize(u64 z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
inline u64 salt_value(u64 seed, unsigned batch, unsigned s, u64 leaf) {
    u64 z = seed + 0x9e3779b97f4a7c15ULL * (u64)(batch * 4 + s + 1) + leaf;
    return from_u64(splitmix64_finalize(z + 0x9e3779b97f4a7c15ULL));
}

struct PolyBatch {
    size_t n = 0, ncols = 0;
    unsigned rate_bits = 3;
    bool blinding = false;
    std::vector<std::vector<u64>> coeffs;  // [ncols][n]
    MerkleTree tree;                       // 8n leaves of ncols (+4 if blinding), bit-reversed order
    const u64* lde_row(size_t natural_index) const {
        return tree.leaf(reverse_bits(natural_index, log2_strict(n) + rate_bits));
    }
};
// salts: nullptr or [4][8n] indexed by leaf (storage) position; if nullptr and blinding, salt_value(seed,...)
PolyBatch batch_from_coeffs(std::vector<std::vector<u64>> coeffs, unsigned rate_bits, bool blinding,
                            unsigned cap_height, const u64* salts, u64 seed, unsigned batch_id);
PolyBatch batch_from_values(std::vector<std::vector<u64>> values, unsigned rate_bits, bool blinding,
                            unsigned cap_height, const u64* salts, u64 seed, unsigned batch_id);

struct CircuitData {
    CommonData common;
    VerifierOnly vo;
    PolyBatch constants_sigmas;                   // built from constants‖sigmas values over H
    std::vector<std::vector<u64>> sigma_values;   // [num_routed][n] (values over H, natural order)
};
// const_sigma_values: [(num_constants + num_routed)][n]
CircuitData circuit_from_values(const CommonData& c, std::vector<std::vector<u64>> const_sigma_values);

// [num_challenges*(1+npp)][n] in the committed order Z_0..Z_{c-1}, pp(ch0), pp(ch1)
std::vector<std::vector<u64>> partial_products_and_zs(const CircuitData& cd, const std::vector<std::vector<u64>>& wires,
                                                      const std::vector<u64>& betas, const std::vector<u64>& gammas);
// num_challenges*qdf coefficient chunks of length n
std::vector<std::vector<u64>> compute_quotient_chunks(const CircuitData& cd, const PolyBatch& wires_b, const PolyBatch& zs_b,
                                                      const Digest& pi_hash, const std::vector<u64>& betas,
                                                      const std::vector<u64>& gammas, const std::vector<u64>& alphas);

struct ProveTrace {  // intermediates exported for stage-level parity tests
    std::vector<u64> betas, gammas, alphas;
    E2 zeta, fri_alpha;
    std::vector<E2> fri_betas;
    std::vector<std::vector<u64>> zs_pp_values;      // [20][n]
    std::vector<std::vector<u64>> quotient_chunks;   // [16][n]
    std::vector<E2> final_poly_pre_fri;              // n ext coeffs entering FRI
    std::vector<std::vector<E2>> fri_layer_values;   // values (bit-reversed order) committed per layer
    std::vector<size_t> query_indices;
};

// salts: nullptr or [3][4][8n]; pow rule = MIN. Throws on zeta in subgroup.
Proof prove(const CircuitData& cd, const std::vector<std::vector<u64>>& wires, const std::vector<u64>& public_inputs,
            const u64* salts, u64 salt_seed, ProveTrace* trace = nullptr);

// ---- synthetic wormhole-/voting-shaped circuits ----
struct SynthSpec {
    unsigned min_degree_bits = 0;  // pad with Noop rows up to at least this
    bool zk = false;
    size_t n_poseidon = 488, n_base_sum = 3800, n_arith = 2520, n_const = 100;
    size_t num_public_inputs = 16;
    u64 seed = 1;
    // rows of the recursion gate set; any non-zero count switches the circuit to the 14-gate set that
    // `verify_proof` (aggregator/src/circuits/tree.rs:119) instantiates, with 4 selector groups. The aggregator builds its
    // chunk circuits with the config of the proofs it verifies (tree.rs:111, aggregator.rs:21: standard_recursion_zk_config),
    // so they are zero-knowledge like the leaf circuit; the tests at tree.rs:165 use the non-zk config
    size_t n_arith_ext = 0, n_mul_ext = 0, n_reducing = 0, n_reducing_ext = 0, n_random_access = 0, n_exp = 0, n_coset = 0,
           n_mds = 0;
};
struct SynthCircuit {
    CommonData common;
    std::vector<std::vector<u64>> const_sigma_values;  // [num_constants + 80][n] (selectors, 2 gate constants, sigmas)
    std::vector<std::vector<u64>> wires;               // [135][n]
    std::vector<u64> public_inputs;
};
SynthCircuit make_synth_circuit(const SynthSpec& spec);
// returns "" if every gate constraint and copy constraint holds on the witness
std::string check_witness(const SynthCircuit& sc);

}  // namespace orc
