// ORACLE — TEST INFRASTRUCTURE ONLY (see verifier.cpp header).
#pragma once
#include "circuit.hpp"

namespace orc {

struct Challenges {
    Digest pi_hash{};
    std::vector<u64> betas, gammas, alphas;
    E2 zeta, fri_alpha;
    std::vector<E2> fri_betas;
    u64 pow_response = 0;
    std::vector<size_t> query_indices;
};

// Fiat-Shamir transcript replay (A.4).
Challenges derive_challenges(const CommonData& c, const VerifierOnly& vo, const Proof& pr);
// Returns "" if the proof verifies, otherwise the reason it was rejected.
std::string verify_proof(const CommonData& c, const VerifierOnly& vo, const Proof& pr);

}  // namespace orc
