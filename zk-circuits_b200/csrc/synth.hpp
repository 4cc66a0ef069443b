// Synthetic workload generator (see synth.cpp). Host-only.
#pragma once
#include <cstdint>
#include <vector>
#include "field.cuh"

namespace zkb {

struct SynthSpec {
    unsigned min_degree_bits = 0;
    bool zk = false;
    size_t n_poseidon = 488, n_base_sum = 3800, n_arith = 2520, n_const = 100;
    size_t num_public_inputs = 16;
    u64 seed = 1;
    // rows of the recursion gate set: any non-zero count selects the 14-gate set a recursive-verifier
    // circuit has (aggregator/src/circuits/tree.rs:119) with its four selector groups; zk as configured (the aggregator
    // inherits the leaf circuit's standard_recursion_zk_config, aggregator.rs:21 / tree.rs:111)
    size_t n_arith_ext = 0, n_mul_ext = 0, n_reducing = 0, n_reducing_ext = 0, n_random_access = 0, n_exp = 0, n_coset = 0,
           n_mds = 0;
    bool recursion() const {
        return n_arith_ext + n_mul_ext + n_reducing + n_reducing_ext + n_random_access + n_exp + n_coset + n_mds > 0;
    }
};
struct SynthCircuit {
    std::vector<uint8_t> common;                       // CommonCircuitData::to_bytes
    u64 degree_bits = 0;
    std::vector<u64> reduction_arity_bits;
    std::vector<std::vector<u64>> const_sigma_values;  // [num_selectors + 2 + 80][n] values over H
    std::vector<std::vector<u64>> wires;               // [135][n]
    std::vector<u64> public_inputs;
};
SynthCircuit make_synth_circuit(const SynthSpec& spec);

}  // namespace zkb
