/* zkb200 — C ABI of the B200-native Plonky2 proving backend (libzkb200.so).
 *
 * This is the drop-in boundary for the reference's prove path. The reference has no FFI of its own:
 * the seam is the Rust method call into the qp-plonky2 dependency,
 *     ProverCircuitData::<F,C,D>::prove(&self, PartialWitness<F>) -> Result<ProofWithPublicInputs<F,C,D>>
 * at /root/reference/wormhole/prover/src/lib.rs:233-237 and CircuitData::prove at
 * /root/reference/wormhole/aggregator/src/circuits/tree.rs:136. A Rust shim runs
 * witness generation on the CPU, flattens the wire matrix and calls zkb_prove(); the bytes it gets back
 * are ProofWithPublicInputs::to_bytes() (reference: wormhole/tests/src/prover/prover_tests.rs:66).
 *
 * Conventions
 *   - all field elements are canonical Goldilocks u64 (< 2^64 - 2^32 + 1), little endian in byte blobs;
 *   - matrices are column-major [col][row]: element (c, r) at base[c * rows + r];
 *   - every pointer is a HOST pointer unless the name ends in _dev; the caller owns every buffer for the
 *     duration of the call only (pinned host memory makes the copies asynchronous DMA);
 *   - functions return 0 (ZKB_OK) or a negative zkb_status; zkb_last_error() gives the thread-local text;
 *   - a zkb_circuit is bound to one device and is NOT re-entrant: one per worker thread / stream.
 *   - no CPU fallback exists: every entry point fails with ZKB_E_CUDA when no sm_100 device is usable.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum zkb_status {
    ZKB_OK = 0,
    ZKB_E_ARG = -1,               /* null pointer, bad size, non-canonical field element            */
    ZKB_E_PARSE = -2,             /* malformed CommonCircuitData bytes                              */
    ZKB_E_UNSUPPORTED_GATE = -3,  /* gate tag / config outside the implemented set                  */
    ZKB_E_UNSAT = -4,             /* reserved (optional vanishing-identity self-check at zeta); not returned by this
                                     version: like the CPU prover, a witness that violates a constraint yields a
                                     proof the verifier rejects (witness generation, which catches it, stays in Rust) */
    ZKB_E_ZETA_IN_SUBGROUP = -5,  /* reference: "Opening point is in the subgroup."                 */
    ZKB_E_CUDA = -6,
    ZKB_E_NCCL = -7,
    ZKB_E_BUFFER = -8,            /* output buffer too small; required size is reported             */
    ZKB_E_DIGEST = -9             /* supplied circuit digest differs from the recomputed one        */
} zkb_status;

typedef struct zkb_circuit zkb_circuit;

const char* zkb_version(void);
const char* zkb_last_error(void);
int zkb_device_count(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py reports the per-step delta) */
unsigned long long zkb_kernel_launch_count(void);

/* ---- circuit context: replaces the prover-side use of ProverOnlyCircuitData + CommonCircuitData
 * (wormhole/prover/src/lib.rs:114-130; built by circuit/src/circuit.rs:98-108).
 *   common_bin        CommonCircuitData::to_bytes(&DefaultGateSerializer) (= generated-bins/common.bin,
 *                     circuit-builder/src/lib.rs:36-39)
 *   const_sigma       [(num_constants + num_routed_wires)][n] column-major; coefficient form
 *                     (prover_only.constants_sigmas_commitment.polynomials) if is_values == 0,
 *                     evaluations over the subgroup H if is_values != 0
 *   circuit_digest    prover_only.circuit_digest, or NULL to accept the recomputed one
 * Builds the constants/sigmas LDE + Merkle tree on the device once; all work buffers are allocated here. */
int zkb_circuit_create(const uint8_t* common_bin, size_t common_len, const uint64_t* const_sigma, int is_values,
                       const uint64_t circuit_digest[4], int device, zkb_circuit** out);
int zkb_circuit_destroy(zkb_circuit* c);
/* VerifierOnlyCircuitData: cap_out receives 2^cap_height digests (4 u64 each), digest_out the circuit digest */
int zkb_circuit_verifier_only(const zkb_circuit* c, uint64_t* cap_out, size_t cap_words, uint64_t digest_out[4]);
size_t zkb_proof_size(const zkb_circuit* c);

/* pow_rule: how the FRI proof-of-work witness is chosen (the CPU prover's rayon find_any is not deterministic) */
#define ZKB_POW_MIN 0u /* smallest valid witness = CPU result with RAYON_NUM_THREADS=1 */

/* ---- prove: replaces circuit_data.prove(partial_witness) after witness generation.
 *   wires          [num_wires][n] column-major full witness (partition_witness.full_witness().wire_values)
 *   public_inputs  n_pi field elements
 *   salts          NULL, or [3][4][n << rate_bits]: blinding salt columns for the wires / Z-partial-product /
 *                  quotient batches, indexed by leaf position; ignored unless the circuit is zero-knowledge.
 *                  When NULL a zk circuit draws salts on the device from salt_seed (documented SplitMix64 stream)
 *   proof_out      receives ProofWithPublicInputs::to_bytes(); *proof_len = bytes written (or required, on
 *                  ZKB_E_BUFFER) */
int zkb_prove(zkb_circuit* c, const uint64_t* wires, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts,
              uint64_t salt_seed, uint32_t pow_rule, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
/* split form used by the benchmark: upload once, prove from HBM-resident wires */
int zkb_witness_upload(zkb_circuit* c, const uint64_t* wires);
int zkb_prove_resident(zkb_circuit* c, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts, uint64_t salt_seed,
                       uint32_t pow_rule, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
/* per-stage device times (ms, CUDA events on the circuit's stream) of the last prove; returns count written.
 * order: wires_intt, wires_lde, wires_merkle, partial_products, zs_commit, quotient, quotient_commit, openings,
 *        fri_combine, fri_commit, pow, queries, total, then two host-side figures: milliseconds spent in the Fiat-Shamir
 *        sponge (serial, between launches) and its permutation count */
int zkb_last_timings(const zkb_circuit* c, float* ms_out, int cap);
#define ZKB_NUM_TIMINGS 15

/* ---- stage-level entry points (parity tests + microbenchmarks); host pointers, run on `device` ---- */
/* width-12 Poseidon permutation of `count` states (12 u64 each), in place */
int zkb_poseidon_permute_batch(uint64_t* states, size_t count, int device);
/* PolynomialBatch::from_values / from_coeffs without hashing: coeffs_out [ncols][n] (may be NULL),
 * lde_out [ncols][n << rate_bits] in leaf (bit-reversed) order (may be NULL) */
int zkb_lde_batch(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, int from_coeffs, uint64_t* coeffs_out,
                  uint64_t* lde_out, int device);
/* MerkleTree::new over column-major leaves [width][num_leaves]; digests_out (may be NULL) receives all levels
 * bottom-up down to the cap level; cap_out receives 2^cap_height digests */
int zkb_merkle_commit(const uint64_t* leaves, size_t width, size_t num_leaves, unsigned cap_height, uint64_t* digests_out,
                      uint64_t* cap_out, int device);
/* fused from_values: iNTT + coset LDE + Merkle; only the cap comes back. times_ms (may be NULL) receives
 * {lde_ms, merkle_ms} measured with CUDA events on the launching stream, inputs already resident */
int zkb_commit_batch(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps,
                     uint64_t* cap_out, float* times_ms, int device);
/* One GPU's share of a COSET-SHARDED from_values commitment (multi-GPU mode for a single large batch): in leaf order LDE
 * coset j is the contiguous leaf block bitrev(j), so a rank that owns leaf blocks [blk_lo, blk_hi) of the 2^rate_bits
 * blocks computes the (cheap) iNTT of every column, the LDE of its blocks only, their leaf hashes and whole Merkle
 * subtrees, and ends with its (blk_hi - blk_lo) * 2^(cap_height - rate_bits) cap digests — the caller all-gathers
 * 16 x 32 bytes (NCCL) to obtain MerkleTree::cap. Needs cap_height >= rate_bits and a power-of-two aligned block range. */
int zkb_commit_cosets(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, unsigned blk_lo,
                      unsigned blk_hi, int reps, uint64_t* cap_part_out, float* times_ms, int device);
/* wires_permutation_partial_products_and_zs: out [num_challenges*(1+num_partial_products)][n] */
int zkb_partial_products(zkb_circuit* c, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas, uint64_t* out);
/* compute_quotient_polys from wire / Z-partial-product VALUES (unsalted): out [num_challenges*qdf][n] coefficients */
int zkb_quotient(zkb_circuit* c, const uint64_t* wires, const uint64_t* zs_pp, const uint64_t* public_inputs, size_t n_pi,
                 const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, uint64_t* out);

/* ---- synthetic workload generator (host code; stands in for the Rust side of the boundary, i.e.
 * CircuitBuilder::build_prover + witness generation, which need a Rust toolchain). Produces a circuit with the
 * reference wormhole circuit's configuration, gate set and row mix and a satisfying witness.
 * Used by bench.py / smoke() / tests to obtain workloads; not part of the proving path. ---- */
typedef struct zkb_synth zkb_synth;
int zkb_synth_create(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith, size_t n_const,
                     size_t num_public_inputs, uint64_t seed, zkb_synth** out);
/* recursion-shaped circuit (configs #4/#5: the gate set a recursive-verifier circuit instantiates at
 * wormhole/aggregator/src/circuits/tree.rs:119 — SURVEY.md App. C.2 — with four selector groups; zk as the aggregator's
 * chunk circuits are, which inherit the leaf circuit's standard_recursion_zk_config, aggregator.rs:21 / tree.rs:111):
 * recursion_rows[8] = rows of ArithmeticExtension, MulExtension, Reducing, ReducingExtension, RandomAccess,
 * Exponentiation, CosetInterpolation, PoseidonMds on top of the base counts; const_sigma_values is then [6 + 80][n] */
int zkb_synth_create_recursion(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith, size_t n_const,
                               size_t num_public_inputs, uint64_t seed, const size_t recursion_rows[8], zkb_synth** out);
int zkb_synth_destroy(zkb_synth* s);
/* number of constant columns (selectors + gate constants) of the synthetic circuit: const_sigma_values has this + 80 columns */
size_t zkb_synth_num_constants(const zkb_synth* s);
size_t zkb_synth_common_len(const zkb_synth* s);
size_t zkb_synth_degree(const zkb_synth* s);
/* any output pointer may be NULL: common [common_len] bytes, const_sigma_values [num_constants + 80][n], wires [135][n],
 * public_inputs [num_public_inputs] */
int zkb_synth_get(const zkb_synth* s, uint8_t* common, uint64_t* const_sigma_values, uint64_t* wires, uint64_t* public_inputs);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
