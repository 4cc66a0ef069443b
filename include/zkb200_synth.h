/* zkb200_synth — synthetic workload generator (libzkb200_synth.so). TEST / BENCH TOOLING, not part of the proving library:
 * it stands in for the Rust side of the boundary — CircuitBuilder::build_prover
 * (/root/reference/wormhole/circuit/src/circuit.rs:98-108) and witness generation
 * (/root/reference/wormhole/prover/src/lib.rs:209-225) — which need a Rust toolchain. It produces a circuit with the reference
 * wormhole circuit's configuration, gate set and row mix and a satisfying witness, i.e. the inputs
 * zkb_circuit_create() / zkb_prove() consume. Host code only; libzkb200.so neither links nor exports any of it. */
#ifndef ZKB200_SYNTH_H
#define ZKB200_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the libraries are built with -fvisibility=hidden: only this C ABI is exported */
#endif

/* 0 on success, -1 on a bad argument (text via zkb_synth_last_error) */
const char* zkb_synth_last_error(void);
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

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ZKB200_SYNTH_H */
