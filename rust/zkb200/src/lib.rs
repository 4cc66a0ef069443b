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
    /// Salts of a zero-knowledge circuit come from the library's CSPRNG (ChaCha20 on the device keyed from the OS RNG per
    /// proof) — never from a caller-chosen seed. ZKB_CHECK_WITNESS turns an unsatisfied witness into `Err` like the
    /// reference's generators do (voting/src/lib.rs:399-403).
    pub fn prove(&mut self, po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>,
                 pw: PartialWitness<F>) -> Result<ProofWithPublicInputs<F, C, D>> {
        self.prove_with(po, cd, pw, None)
    }

    /// `salts`: explicit blinding columns `[3][4][8n]` (wires, Z/partial products, quotient; leaf order) — only for
    /// replaying a CPU proof byte for byte (see `Fixture`).
    pub fn prove_with(&mut self, po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>,
                      pw: PartialWitness<F>, salts: Option<&[u64]>) -> Result<ProofWithPublicInputs<F, C, D>> {
        let pwit = generate_partial_witness(pw, po, cd)?;                 // CPU: generator graph stays in Rust
        let public_inputs: Vec<u64> = pwit.get_targets(&po.public_inputs).iter().map(|x| x.to_canonical_u64()).collect();
        let witness = pwit.full_witness();
        let mut wires = Vec::with_capacity(cd.config.num_wires * cd.degree());
        for col in &witness.wire_values { wires.extend(col.iter().map(|x| x.to_canonical_u64())); }
        let cap = unsafe { zkb200_sys::zkb_proof_size(self.raw) };
        let mut bytes = vec![0u8; cap];
        let mut len = 0usize;
        let rc = unsafe { zkb200_sys::zkb_prove(self.raw, wires.as_ptr(), public_inputs.as_ptr(), public_inputs.len(),
                                                salts.map_or(std::ptr::null(), |s| s.as_ptr()), 0,
                                                zkb200_sys::ZKB_POW_MIN | zkb200_sys::ZKB_CHECK_WITNESS,
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
/// `salts` is REQUIRED for a zero-knowledge circuit (every production config: prover/benches/prover.rs:12,
/// aggregator.rs:21): the `[3][4][8n]` blinding columns the CPU prover drew for the wires, Z/partial-product and quotient
/// batches, in leaf order. They are internal to `PolynomialBatch::from_values/from_coeffs` (the proof only reveals the 28
/// opened leaves), so the exporting host runs qp-plonky2 with a one-line hook there that records `salt_vecs` — or proves with
/// the non-zk config, where `salts` is `None`.
pub struct Fixture {
    pub common_bin: Vec<u8>,
    pub const_sigma_coeffs: Vec<u64>,
    pub circuit_digest: [u64; 4],
    pub wires: Vec<u64>,
    pub public_inputs: Vec<u64>,
    pub salts: Option<Vec<u64>>,
    pub proof_bytes: Vec<u8>,
}
pub fn export_fixture(po: &ProverOnlyCircuitData<F, C, D>, cd: &CommonCircuitData<F, D>, pw: PartialWitness<F>,
                      salts: Option<Vec<u64>>, cpu_proof: &ProofWithPublicInputs<F, C, D>) -> Result<Fixture> {
    if cd.config.zero_knowledge && salts.is_none() {
        return Err(anyhow!("zero-knowledge circuit: the CPU prover's salt columns are needed to replay its proof"));
    }
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
    Ok(Fixture { common_bin, const_sigma_coeffs, circuit_digest, wires, public_inputs, salts, proof_bytes: cpu_proof.to_bytes() })
}
impl Fixture {
    /// The directory layout `zkb200.fixture.load` (Python) and `tests/test_gpu_fixture.py` read: raw little-endian u64
    /// arrays, `salts.u64` only for zero-knowledge circuits.
    pub fn write_dir(&self, dir: &std::path::Path) -> std::io::Result<()> {
        fn u64s(v: &[u64]) -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes()).collect() }
        std::fs::create_dir_all(dir)?;
        std::fs::write(dir.join("common.bin"), &self.common_bin)?;
        std::fs::write(dir.join("const_sigma_coeffs.u64"), u64s(&self.const_sigma_coeffs))?;
        std::fs::write(dir.join("circuit_digest.u64"), u64s(&self.circuit_digest))?;
        std::fs::write(dir.join("wires.u64"), u64s(&self.wires))?;
        std::fs::write(dir.join("public_inputs.u64"), u64s(&self.public_inputs))?;
        if let Some(s) = &self.salts { std::fs::write(dir.join("salts.u64"), u64s(s))?; }
        std::fs::write(dir.join("proof.bin"), &self.proof_bytes)
    }
}
