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
};
struct SynthCircuit {
    std::vector<uint8_t> common;                       // CommonCircuitData::to_bytes
    u64 degree_bits = 0;
    std::vector<u64> reduction_arity_bits;
    std::vector<std::vector<u64>> const_sigma_values;  // [4 + 80][n] values over H
    std::vector<std::vector<u64>> wires;               // [135][n]
    std::vector<u64> public_inputs;
};
SynthCircuit make_synth_circuit(const SynthSpec& spec);

}  // namespace zkb
