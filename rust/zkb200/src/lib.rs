//! Drop-in for `circuit_data.prove(pw)` of the reference (wormhole/prover/src/lib.rs:233-237,
//! wormhole/aggregator/src/circuits/tree.rs:136). Untested source: no Rust toolchain in the build image (see ../README.md).
use anyhow::{anyhow, Result};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::iop::{generator::generate_partial_witness, witness::PartialWitness};
use plonky2::plonk::{circuit_data::{CommonCircuitData, ProverOnlyCircuitData}, proof::ProofWithPublicInputs};
use plonky2::util::serialization::DefaultGateSerializer;
use zk_circuits_common::circuit::{C, D, F};

pub struct B200Circuit { raw: *mut zkb200_sys::zkb_circuit }
unsafe impl Send for B200Circuit {}

impl B200Circuit {
    /// Built once per circuit (per GPU): uploads constants/sigmas, builds their LDE + Merkle tree on the device.
    pub fn new(po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>, device: i32) -> Result<Self> {
        let common = cd.to_bytes(&DefaultGateSerializer).map_err(|e| anyhow!("{e:?}"))?;
        let n = cd.degree();
        let mut cs = Vec::with_capacity(po.constants_sigmas_commitment.polynomials.len() * n);
        for p in &po.constants_sigmas_commitment.polynomials {           // coefficient form, column-major
            cs.extend(p.coeffs.iter().map(|x| x.to_canonical_u64()));
        }
        let digest: Vec<u64> = po.circuit_digest.elements.iter().map(|x| x.to_canonical_u64()).collect();
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { zkb200_sys::zkb_circuit_create(common.as_ptr(), common.len(), cs.as_ptr(), 0,
                                                         digest.as_ptr(), device, &mut raw) };
        if rc != 0 { return Err(anyhow!("zkb_circuit_create: {}", last_error())); }
        Ok(Self { raw })
    }

    /// Drop-in for `ProverCircuitData::prove` (wormhole/prover/src/lib.rs:234-235, aggregator tree.rs:136).
    pub fn prove(&mut self, po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>,
                 pw: PartialWitness<F>, salt_seed: u64) -> Result<ProofWithPublicInputs<F, C, D>> {
        let pwit = generate_partial_witness(pw, po, cd)?;                 // CPU: generator graph stays in Rust
        let public_inputs: Vec<u64> = pwit.get_targets(&po.public_inputs).iter().map(|x| x.to_canonical_u64()).collect();
        let witness = pwit.full_witness();
        let mut wires = Vec::with_capacity(cd.config.num_wires * cd.degree());
        for col in &witness.wire_values { wires.extend(col.iter().map(|x| x.to_canonical_u64())); }
        let cap = unsafe { zkb200_sys::zkb_proof_size(self.raw) };
        let mut bytes = vec![0u8; cap];
        let mut len = 0usize;
        let rc = unsafe { zkb200_sys::zkb_prove(self.raw, wires.as_ptr(), public_inputs.as_ptr(), public_inputs.len(),
                                                std::ptr::null(), salt_seed, 0 /* ZKB_POW_MIN */,
                                                bytes.as_mut_ptr(), cap, &mut len) };
        if rc != 0 { return Err(anyhow!("Failed to prove: {}", last_error())); }
        bytes.truncate(len);
        ProofWithPublicInputs::from_bytes(bytes, cd).map_err(|e| anyhow!("{e:?}"))
    }
}
impl Drop for B200Circuit { fn drop(&mut self) { unsafe { zkb200_sys::zkb_circuit_destroy(self.raw); } } }
fn last_error() -> String { unsafe { std::ffi::CStr::from_ptr(zkb200_sys::zkb_last_error()) }.to_string_lossy().into_owned() }

/// L4 byte-parity fixture: everything `zkb_prove` needs to reproduce a CPU proof byte for byte.
/// Run the reference prover with RAYON_NUM_THREADS=1 so that its proof-of-work witness is the minimum (ZKB_POW_MIN).
pub struct Fixture {
    pub common_bin: Vec<u8>,
    pub const_sigma_coeffs: Vec<u64>,
    pub circuit_digest: [u64; 4],
    pub wires: Vec<u64>,
    pub public_inputs: Vec<u64>,
    pub proof_bytes: Vec<u8>,
}
pub fn export_fixture(po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>, pw: PartialWitness<F>,
                      cpu_proof: &ProofWithPublicInputs<F, C, D>) -> Result<Fixture> {
    let common_bin = cd.to_bytes(&DefaultGateSerializer).map_err(|e| anyhow!("{e:?}"))?;
    let mut const_sigma_coeffs = Vec::new();
    for p in &po.constants_sigmas_commitment.polynomials {
        const_sigma_coeffs.extend(p.coeffs.iter().map(|x| x.to_canonical_u64()));
    }
    let mut circuit_digest = [0u64; 4];
    for (d, x) in circuit_digest.iter_mut().zip(po.circuit_digest.elements.iter()) { *d = x.to_canonical_u64(); }
    let pwit = generate_partial_witness(pw, po, cd)?;
    let public_inputs = pwit.get_targets(&po.public_inputs).iter().map(|x| x.to_canonical_u64()).collect();
    let witness = pwit.full_witness();
    let mut wires = Vec::new();
    for col in &witness.wire_values { wires.extend(col.iter().map(|x| x.to_canonical_u64())); }
    Ok(Fixture { common_bin, const_sigma_coeffs, circuit_digest, wires, public_inputs, proof_bytes: cpu_proof.to_bytes() })
}
