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
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the libraries are built with -fvisibility=hidden: only this C ABI is exported */
#endif

typedef enum zkb_status {
    ZKB_OK = 0,
    ZKB_E_ARG = -1,               /* null pointer, bad size, non-canonical field element            */
    ZKB_E_PARSE = -2,             /* malformed CommonCircuitData bytes                              */
    ZKB_E_UNSUPPORTED_GATE = -3,  /* gate tag / config outside the implemented set                  */
    ZKB_E_UNSAT = -4,             /* with ZKB_CHECK_WITNESS: the witness violates a gate or copy constraint (the
                                     reference surfaces such inputs as Err from witness generation, e.g.
                                     voting/src/lib.rs:399-403). Without the flag, like the CPU prover, the call
                                     returns a proof the verifier rejects                                           */
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

/* ---- `flags` of the prove calls: low byte = proof-of-work rule, then option bits ----
 * pow rule: how the FRI proof-of-work witness is chosen (the CPU prover's rayon find_any is not deterministic) */
#define ZKB_POW_MIN 0u /* smallest valid witness = CPU result with RAYON_NUM_THREADS=1 */
/* Salt source of a zero-knowledge circuit when `salts` is NULL. DEFAULT (flag clear): a CSPRNG — ChaCha20 on the device,
 * keyed per proof with 256 bits from the OS (getrandom); salt_seed is ignored. This is what production callers use: the
 * blinding columns hide the witness only if a verifier cannot predict them (the CPU prover draws F::rand_vec from an
 * OS-seeded RNG). With ZKB_SALTS_FROM_SEED the salts are the documented SplitMix64 stream of salt_seed — a public bijection
 * of (seed, position), so ONE opened leaf reveals every other salt: for byte-parity tests and benchmarks only. */
#define ZKB_SALTS_FROM_SEED 0x100u
/* Evaluate every constraint of the circuit (gates, Z(1) = 1, the partial-product chain = copy constraints) on the subgroup
 * from the witness values before committing Z, on the device (~2 % of a proof); a violation returns ZKB_E_UNSAT instead of
 * an unverifiable proof. */
#define ZKB_CHECK_WITNESS 0x200u
/* zkb_engine_submit only: prove the witness the context already holds (benchmark's device-resident arm) */
#define ZKB_WITNESS_RESIDENT 0x400u

/* ---- prove: replaces circuit_data.prove(partial_witness) after witness generation.
 *   wires          [num_wires][n] column-major full witness (partition_witness.full_witness().wire_values)
 *   public_inputs  n_pi field elements
 *   salts          NULL, or [3][4][n << rate_bits]: blinding salt columns for the wires / Z-partial-product /
 *                  quotient batches, indexed by leaf position (canonical; checked); ignored unless the circuit is
 *                  zero-knowledge. When NULL a zk circuit draws its salts on the device from a CSPRNG keyed by the OS
 *                  (see ZKB_SALTS_FROM_SEED below for the deterministic test mode that uses salt_seed)
 *   flags          ZKB_POW_MIN | option bits (below)
 *   proof_out      receives ProofWithPublicInputs::to_bytes(); *proof_len = bytes written (or required, on
 *                  ZKB_E_BUFFER) */
int zkb_prove(zkb_circuit* c, const uint64_t* wires, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts,
              uint64_t salt_seed, uint32_t flags, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
/* split form used by the benchmark: upload once, prove from HBM-resident wires */
int zkb_witness_upload(zkb_circuit* c, const uint64_t* wires);
int zkb_prove_resident(zkb_circuit* c, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts, uint64_t salt_seed,
                       uint32_t flags, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* ---- proof engine: asynchronous proving with pinned, double-buffered witness hand-off.
 * The reference fills a PartialWitness (wormhole/prover/src/lib.rs:209-225), runs the generator graph and proves, one
 * blocking call per proof; its aggregator fans such calls out over rayon (aggregator/src/circuits/tree.rs:93-103). An
 * engine owns n_contexts prover contexts of one circuit on one GPU, ONE driver thread that steps all of them (a proof is
 * ~10 GPU stages separated by serial Fiat-Shamir steps; the driver runs whichever context's stage has finished), and
 * n_slots >= n_contexts pinned witness buffers:
 *     slot = zkb_engine_acquire(e, &buf);     blocks while every slot is in use
 *     ... write the wire matrix [num_wires][n] (column-major, canonical) into buf — witness generation for proof k+1
 *         overlaps the GPU work of proofs k, k-1, ... ...
 *     zkb_engine_submit(e, slot, public_inputs, n_pi, salts, salt_seed, flags, proof_out, cap);     returns at once
 *     zkb_engine_wait(e, slot, &len);         blocks until THAT proof is done; returns its status; frees the slot
 * Any number of caller threads may use one engine. public_inputs are copied at submit; salts (if not NULL) and proof_out
 * must stay valid until the wait returns. Argument errors are reported by submit, proving errors by wait (a failed proof
 * fails only its own slot; the context is reset and reused). The driver thread spins while proofs are in flight and sleeps
 * otherwise. zkb_engine_destroy drains queued and running proofs first; call it only when no thread is inside acquire / wait. */
typedef struct zkb_engine zkb_engine;
int zkb_engine_create(const uint8_t* common_bin, size_t common_len, const uint64_t* const_sigma, int is_values,
                      const uint64_t circuit_digest[4], int device, int n_contexts, int n_slots, zkb_engine** out);
int zkb_engine_destroy(zkb_engine* e);
size_t zkb_engine_proof_size(const zkb_engine* e);
/* returns the slot id (>= 0) or a negative zkb_status */
int zkb_engine_acquire(zkb_engine* e, uint64_t** wires_buf);
int zkb_engine_release(zkb_engine* e, int slot);     /* give an acquired slot back without proving */
int zkb_engine_submit(zkb_engine* e, int slot, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts,
                      uint64_t salt_seed, uint32_t flags, uint8_t* proof_out, size_t proof_cap);
int zkb_engine_wait(zkb_engine* e, int slot, size_t* proof_len);
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
/* ---- the same sharding with the exchange INSIDE the library: NCCL over NVLink (the reference is single
 * process — its only parallelism is rayon over chunks, wormhole/aggregator/src/circuits/tree.rs:93-103). One process per GPU:
 * rank 0 calls zkb_comm_unique_id and hands the 128 bytes to the other ranks by any means (the Rust host: its own channel;
 * bench.py: torch.distributed); every rank then calls zkb_comm_create. Errors of the collective layer return ZKB_E_NCCL.
 * NCCL is loaded at run time (libnccl.so.2 or $ZKB_NCCL_LIB); the library has no link-time dependency on it. */
typedef struct zkb_comm zkb_comm;
#define ZKB_COMM_ID_BYTES 128
int zkb_comm_unique_id(uint8_t id_out[ZKB_COMM_ID_BYTES]);
int zkb_comm_create(const uint8_t id[ZKB_COMM_ID_BYTES], int nranks, int rank, int device, zkb_comm** out);
int zkb_comm_destroy(zkb_comm* c);
/* PolynomialBatch::from_values of ONE batch across the ranks of `c` (nranks | 2^rate_bits, cap_height >= rate_bits): every rank
 * passes the same values [ncols][n] but uploads and interpolates only ITS column slice, into a window of its memory that the
 * peers have mapped (CUDA IPC); after one barrier each rank pulls the other slices out of their owners' windows with its copy
 * engines over NVLink and starts the LDE of a slice as soon as its copy has landed; the rank extends and hashes its own
 * leaf blocks and builds their subtrees; one all-gather of the cap digests. Where the peer mapping is unavailable (or with
 * ZKB_SHARDED_P2P=0) the coefficients are all-gathered with NCCL in column chunks behind the LDE instead. cap_out (all
 * 2^cap_height digests) is identical on every rank and equals the single-GPU commitment's cap. times_ms (may be NULL):
 * {lde_ms (iNTT + exchange + LDE), merkle_ms, exchange ms alone (barrier + peer copies, or the NCCL gather; overlapped)} */
int zkb_commit_sharded(zkb_comm* c, const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps,
                       uint64_t* cap_out, float* times_ms);
/* 1 if this communicator's sharded commits pull from peer memory (windows mapped), 0 if they use the NCCL gather
 * (also before the first zkb_commit_sharded call, which is where the windows are set up) */
int zkb_comm_peer_windows(zkb_comm* c);
/* Chunks of the quotient polynomial from COSET-LOCAL evaluations (compute_quotient_polys sharded by coset): q_values
 * [num_challenges][B n] = t on this rank's B = 2^rate_bits / nranks leaf blocks (leaf order); one coset iNTT per block, ONE
 * all-to-all of coefficient slices, an R x R Vandermonde solve per coefficient index. chunks_out [num_challenges][2^rate_bits][n / nranks]:
 * coefficients [rank n / nranks, (rank + 1) n / nranks) of chunk m of challenge ch. times_ms: {interpolation, exchange + solve} */
int zkb_quotient_chunks_sharded(zkb_comm* c, const uint64_t* q_values, size_t num_challenges, size_t n, unsigned rate_bits,
                                uint64_t* chunks_out, float* times_ms);
/* wires_permutation_partial_products_and_zs: out [num_challenges*(1+num_partial_products)][n] */
int zkb_partial_products(zkb_circuit* c, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas, uint64_t* out);
/* compute_quotient_polys from wire / Z-partial-product VALUES (unsalted): out [num_challenges*qdf][n] coefficients */
int zkb_quotient(zkb_circuit* c, const uint64_t* wires, const uint64_t* zs_pp, const uint64_t* public_inputs, size_t n_pi,
                 const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, uint64_t* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
