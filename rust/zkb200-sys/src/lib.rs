//! Raw bindings to `include/zkb200.h` (hand-written, no bindgen). Every item cites the header declaration it mirrors;
//! semantics, ownership and error codes are documented there.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_uint, c_ulonglong};

#[repr(C)]
pub struct zkb_circuit {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkb_engine {
    _private: [u8; 0],
}
#[repr(C)]
pub struct zkb_comm {
    _private: [u8; 0],
}
pub const ZKB_COMM_ID_BYTES: usize = 128;

// zkb_status
pub const ZKB_OK: c_int = 0;
pub const ZKB_E_ARG: c_int = -1;
pub const ZKB_E_PARSE: c_int = -2;
pub const ZKB_E_UNSUPPORTED_GATE: c_int = -3;
pub const ZKB_E_UNSAT: c_int = -4;
pub const ZKB_E_ZETA_IN_SUBGROUP: c_int = -5;
pub const ZKB_E_CUDA: c_int = -6;
pub const ZKB_E_NCCL: c_int = -7;
pub const ZKB_E_BUFFER: c_int = -8;
pub const ZKB_E_DIGEST: c_int = -9;
pub const ZKB_POW_MIN: u32 = 0;
pub const ZKB_SALTS_FROM_SEED: u32 = 0x100;
pub const ZKB_CHECK_WITNESS: u32 = 0x200;
pub const ZKB_WITNESS_RESIDENT: u32 = 0x400;
pub const ZKB_NUM_TIMINGS: usize = 15;

extern "C" {
    pub fn zkb_version() -> *const c_char;
    pub fn zkb_last_error() -> *const c_char;
    pub fn zkb_device_count() -> c_int;
    pub fn zkb_kernel_launch_count() -> c_ulonglong;

    pub fn zkb_circuit_create(common_bin: *const u8, common_len: usize, const_sigma: *const u64, is_values: c_int,
                              circuit_digest: *const u64, device: c_int, out: *mut *mut zkb_circuit) -> c_int;
    pub fn zkb_circuit_destroy(c: *mut zkb_circuit) -> c_int;
    pub fn zkb_circuit_verifier_only(c: *const zkb_circuit, cap_out: *mut u64, cap_words: usize, digest_out: *mut u64) -> c_int;
    pub fn zkb_proof_size(c: *const zkb_circuit) -> usize;

    pub fn zkb_prove(c: *mut zkb_circuit, wires: *const u64, public_inputs: *const u64, n_pi: usize, salts: *const u64,
                     salt_seed: u64, flags: u32, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn zkb_witness_upload(c: *mut zkb_circuit, wires: *const u64) -> c_int;
    pub fn zkb_prove_resident(c: *mut zkb_circuit, public_inputs: *const u64, n_pi: usize, salts: *const u64, salt_seed: u64,
                              flags: u32, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn zkb_engine_create(common_bin: *const u8, common_len: usize, const_sigma: *const u64, is_values: c_int,
                             circuit_digest: *const u64, device: c_int, n_contexts: c_int, n_slots: c_int,
                             out: *mut *mut zkb_engine) -> c_int;
    pub fn zkb_engine_destroy(e: *mut zkb_engine) -> c_int;
    pub fn zkb_engine_proof_size(e: *const zkb_engine) -> usize;
    pub fn zkb_engine_acquire(e: *mut zkb_engine, wires_buf: *mut *mut u64) -> c_int;
    pub fn zkb_engine_release(e: *mut zkb_engine, slot: c_int) -> c_int;
    pub fn zkb_engine_submit(e: *mut zkb_engine, slot: c_int, public_inputs: *const u64, n_pi: usize, salts: *const u64,
                             salt_seed: u64, flags: u32, proof_out: *mut u8, proof_cap: usize) -> c_int;
    pub fn zkb_engine_wait(e: *mut zkb_engine, slot: c_int, proof_len: *mut usize) -> c_int;
    pub fn zkb_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn zkb_comm_create(id: *const u8, nranks: c_int, rank: c_int, device: c_int, out: *mut *mut zkb_comm) -> c_int;
    pub fn zkb_comm_destroy(c: *mut zkb_comm) -> c_int;
    pub fn zkb_commit_sharded(c: *mut zkb_comm, values: *const u64, ncols: usize, n: usize, rate_bits: c_uint, cap_height: c_uint,
                              reps: c_int, cap_out: *mut u64, times_ms: *mut c_float) -> c_int;
    pub fn zkb_comm_peer_windows(c: *mut zkb_comm) -> c_int;
    pub fn zkb_quotient_chunks_sharded(c: *mut zkb_comm, q_values: *const u64, num_challenges: usize, n: usize, rate_bits: c_uint,
                                       chunks_out: *mut u64, times_ms: *mut c_float) -> c_int;
    pub fn zkb_last_timings(c: *const zkb_circuit, ms_out: *mut c_float, cap: c_int) -> c_int;

    pub fn zkb_poseidon_permute_batch(states: *mut u64, count: usize, device: c_int) -> c_int;
    pub fn zkb_lde_batch(values: *const u64, ncols: usize, n: usize, rate_bits: c_uint, from_coeffs: c_int,
                         coeffs_out: *mut u64, lde_out: *mut u64, device: c_int) -> c_int;
    pub fn zkb_merkle_commit(leaves: *const u64, width: usize, num_leaves: usize, cap_height: c_uint, digests_out: *mut u64,
                             cap_out: *mut u64, device: c_int) -> c_int;
    pub fn zkb_commit_batch(values: *const u64, ncols: usize, n: usize, rate_bits: c_uint, cap_height: c_uint, reps: c_int,
                            cap_out: *mut u64, times_ms: *mut c_float, device: c_int) -> c_int;
    pub fn zkb_commit_cosets(values: *const u64, ncols: usize, n: usize, rate_bits: c_uint, cap_height: c_uint, blk_lo: c_uint,
                             blk_hi: c_uint, reps: c_int, cap_part_out: *mut u64, times_ms: *mut c_float, device: c_int) -> c_int;
    pub fn zkb_partial_products(c: *mut zkb_circuit, wires: *const u64, betas: *const u64, gammas: *const u64, out: *mut u64) -> c_int;
    pub fn zkb_quotient(c: *mut zkb_circuit, wires: *const u64, zs_pp: *const u64, public_inputs: *const u64, n_pi: usize,
                        betas: *const u64, gammas: *const u64, alphas: *const u64, out: *mut u64) -> c_int;
}
